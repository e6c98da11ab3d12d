"""GPU: one convolution layer (forward, weight gradient, data gradient) through the layer primitives, on the TMA-fed
im2col tcgen05 kernels (operands with lo planes) and on the SIMT-producer kernels, against float64 torch conv2d.
Geometries: AtariCnn c2 / c3 (cnn/base.rs:29-32) at ragged and whole-tile batch sizes, plus a 96-channel layer."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from border_b200 import _lib as L


def _tma_count():
    n, r = C.c_uint64(), C.c_uint64()
    L.check(L.lib().bb_tma_stats(C.byref(n), C.byref(r), 0))
    return n.value, r.value


def _conv(mode, use_tma, x, w, bias, dy, s):
    B, H, W, Cc = x.shape
    OC, k = w.shape[0], w.shape[1]
    OH, OW = (H - k) // s + 1, (W - k) // s + 1
    shape = {0: (B, OH, OW, OC), 1: w.shape, 2: x.shape}[mode]
    out = np.empty(shape, np.float32)
    L.check(L.lib().bb_test_conv(0, mode, use_tma, B, Cc, H, W, OC, k, s, x.ctypes.data, w.ctypes.data,
                                 bias.ctypes.data if bias is not None else None,
                                 dy.ctypes.data if dy is not None else None, out.ctypes.data))
    return out


GEOMS = [(32, 32, 20, 20, 64, 4, 2), (256, 32, 20, 20, 64, 4, 2), (32, 64, 9, 9, 64, 3, 1), (256, 64, 9, 9, 64, 3, 1),
         (5, 96, 12, 10, 32, 2, 2)]


@pytest.mark.parametrize("use_tma", [1, 0])
@pytest.mark.parametrize("geom", GEOMS)
def test_conv_layer_matches_float64(geom, use_tma, monkeypatch):
    monkeypatch.setenv("BB_TMA", str(use_tma))   # 0: the SIMT-producer tcgen05 kernels (tc_gemm.cuh)
    B, Cc, H, W, OC, k, s = geom
    rng = np.random.default_rng(B + Cc + H)
    x = rng.standard_normal((B, H, W, Cc)).astype(np.float32)
    w = (rng.standard_normal((OC, k, k, Cc)) / np.sqrt(k * k * Cc)).astype(np.float32)
    bias = rng.standard_normal(OC).astype(np.float32)
    OH, OW = (H - k) // s + 1, (W - k) // s + 1
    dy = rng.standard_normal((B, OH, OW, OC)).astype(np.float32)
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.from_numpy(w).double().permute(0, 3, 1, 2).requires_grad_(True)
    yt = F.conv2d(xt, wt, torch.from_numpy(bias).double(), stride=s)
    yt.backward(torch.from_numpy(dy).double().permute(0, 3, 1, 2))
    refs = {0: yt.detach().permute(0, 2, 3, 1).numpy(), 1: wt.grad.permute(0, 2, 3, 1).numpy(), 2: xt.grad.permute(0, 2, 3, 1).numpy()}
    for mode in (0, 1, 2):
        n0, _ = _tma_count()
        got = _conv(mode, use_tma, x, w, bias if mode == 0 else None, dy if mode else None, s)
        n1, rej = _tma_count()
        M_, N_, K_ = {0: (B * OH * OW, OC, k * k * Cc), 1: (k * k * Cc, OC, B * OH * OW), 2: (B * (H // s) * (W // s), s * s * Cc, (k // s) ** 2 * OC)}[mode]
        if use_tma and Cc % 32 == 0 and OC % 32 == 0 and M_ * N_ * K_ >= 1 << 22 and M_ >= 64:  # (smaller problems stay on CUDA cores)
            assert n1 > n0 and rej == 0, "the TMA path was expected to take this layer (mode %d)" % mode
        ref = refs[mode]
        kdim = {0: k * k * Cc, 1: B * OH * OW, 2: k * k * OC}[mode]
        tol = 1.2e-5 * (np.abs(ref).max() + 1.0) * max(1.0, (kdim / 4096.0) ** 0.5)
        err = np.abs(got - ref).max()
        assert err <= tol, (mode, use_tma, err, tol)
