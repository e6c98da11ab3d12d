"""GPU: the TMA-fed tcgen05 GEMM (tma_gemm.cuh: cp.async.bulk.tensor producers, K-major and MN-major UMMA operands,
3xTF32 from the fp32 plane + its lo plane, split-K finished in the kernel) against float64 and against the
SIMT-producer tcgen05 kernel, for the dense operand patterns (forward / data gradient / weight gradient) incl.
ragged shapes; the im2col patterns are covered by the DQN-on-CNN parity tests run with and without the path."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from border_b200 import _lib as L
from tests.test_tc_gemm_gpu import _ref, _run


def _tma_count():
    n, r = C.c_uint64(), C.c_uint64()
    L.check(L.lib().bb_tma_stats(C.byref(n), C.byref(r), 0))
    return n.value, r.value


def _operands(mode, M, N, K, rng):
    if mode == 0:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    elif mode == 2:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    else:
        A, B = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    return A.astype(np.float32), B.astype(np.float32)


# (mode, M, N, K): whole tiles, ragged M / N / K (TMA zero-fills out-of-bounds boxes), split-K shapes
SHAPES = [(0, 128, 64, 32), (0, 128, 64, 256), (0, 256, 512, 3136), (0, 300, 72, 100), (0, 20736, 64, 512),
          (0, 12544, 64, 576), (0, 1000, 36, 40), (0, 128, 128, 8), (0, 512, 32, 4096),
          (2, 256, 3136, 512), (2, 128, 64, 32), (2, 300, 96, 100), (2, 20736, 512, 64), (2, 1000, 32, 40),
          (3, 512, 3136, 256), (3, 64, 512, 20736), (3, 128, 64, 32), (3, 96, 160, 1000), (3, 32, 256, 40000)]


@pytest.mark.parametrize("case", SHAPES)
def test_tma_gemm_matches_float64(case):
    mode, M, N, K = case
    rng = np.random.default_rng(M * 31 + N * 7 + K + mode)
    A, B = _operands(mode, M, N, K, rng)
    bias = rng.standard_normal(N).astype(np.float32) if mode == 0 else None
    relu = 1 if mode == 0 and K % 2 == 0 else 0
    ref = _ref(mode, A, B, bias, relu)
    n0, _ = _tma_count()
    got = _run(mode, 3, A, B, M, N, K, bias, relu)  # raises if the TMA path declines or C's lo plane is wrong
    n1, rej = _tma_count()
    assert n1 == n0 + 1 and rej == 0
    mag = np.abs(ref).max() + 1.0
    assert np.abs(got - ref).max() <= 1.2e-5 * mag * max(1.0, (K / 4096.0) ** 0.5), (np.abs(got - ref).max(), mag)


def test_fp32_plane_is_the_tf32_hi_operand():
    """The tensor core ignores the low 13 mantissa bits of a TF32 operand, so feeding the fp32 plane itself equals
    feeding the explicitly truncated hi part: the TMA kernel (raw plane + lo plane) and the SIMT-producer kernel (hi =
    x & 0xffffe000, lo = x - hi, the same MMA order) agree BIT FOR BIT on an unsplit problem."""
    rng = np.random.default_rng(5)
    M, N, K = 20736, 64, 512
    A, B = _operands(0, M, N, K, rng)
    a = _run(0, 3, A, B, M, N, K)
    b = _run(0, 1, A, B, M, N, K)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), np.abs(a - b).max()


def test_tma_split_k_is_deterministic():
    rng = np.random.default_rng(6)
    M, N, K = 256, 512, 3136
    A, B = _operands(0, M, N, K, rng)
    outs = [_run(0, 3, A, B, M, N, K) for _ in range(3)]
    assert all(np.array_equal(outs[0].view(np.uint32), o.view(np.uint32)) for o in outs[1:])
