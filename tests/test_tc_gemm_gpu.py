"""GPU: the tcgen05 (3xTF32) GEMM against a float64 reference and against the fp32 CUDA-core path,
for the three operand patterns (forward / dgrad / wgrad) incl. ragged shapes and split-K."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from border_b200 import _lib as L


def _run(mode, use_tc, A, B, M, N, K, bias=None, relu=0):
    C_ = np.empty((M, N), np.float32)
    L.check(L.lib().bb_test_gemm(0, mode, use_tc, M, N, K, A.ctypes.data, B.ctypes.data,
                                 bias.ctypes.data if bias is not None else None, relu, C_.ctypes.data))
    return C_


def _ref(mode, A, B, bias, relu):
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    C_ = A64 @ B64.T if mode == 0 else (A64 @ B64 if mode == 2 else A64.T @ B64)
    if bias is not None:
        C_ = C_ + bias.astype(np.float64)
    if relu:
        C_ = np.maximum(C_, 0)
    return C_


SHAPES = [(128, 64, 32), (128, 64, 256), (256, 512, 3136), (300, 70, 100), (20736, 64, 512), (64, 6, 512),
          (512, 3136, 256), (12544, 64, 576), (1000, 33, 37), (128, 128, 8)]


@pytest.mark.parametrize("mode", [0, 2, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_tc_gemm_matches_float64(mode, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 31 + N * 7 + K + mode)
    if mode == 0:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    elif mode == 2:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    else:
        A, B = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    A, B = A.astype(np.float32), B.astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) if mode == 0 else None
    relu = 1 if mode == 0 and K % 2 == 0 else 0
    ref = _ref(mode, A, B, bias, relu)
    tcr = _run(mode, 1, A, B, M, N, K, bias, relu)
    simt = _run(mode, 0, A, B, M, N, K, bias, relu)
    # 3xTF32 keeps every product to ~2^-21, but the tensor core accumulates fp32 with truncation:
    # measured error ~1e-6 * |C| * sqrt(k-slices); fp32 FFMA (round-to-nearest) is ~10x tighter.
    mag = np.abs(ref).max() + 1.0
    assert np.abs(tcr - ref).max() <= 1.2e-5 * mag, (np.abs(tcr - ref).max(), mag)
    assert np.abs(simt - ref).max() <= 2e-6 * mag, (np.abs(simt - ref).max(), mag)


def test_dqn_parity_holds_on_the_tensor_core_path(monkeypatch):
    """The whole DQN update (implicit-GEMM convs with gather/u8 operands, wgrad, dgrad) through
    tcgen05 3xTF32: same oracle parity bars as the fp32 CUDA-core path."""
    monkeypatch.setenv("BB_TC", "1")
    from tests.test_dqn_gpu import _run
    _run("cnn", 32, "Mse", False, per=False, clip=False, steps=3, lr=1e-4)
    _run("cnn", 256, "SmoothL1", True, per=False, clip=False, steps=1, lr=1e-4)
    _run("mlp", 64, "Mse", False, per=False, clip=False, steps=3)


def test_dqn_parity_holds_on_the_cuda_core_path(monkeypatch):
    """BB_TC=0: the fp32 CUDA-core implicit GEMM (the tensor-core path's on-device reference)."""
    monkeypatch.setenv("BB_TC", "0")
    from tests.test_dqn_gpu import _run
    _run("cnn", 32, "Mse", False, per=False, clip=False, steps=2, lr=1e-4)
    _run("mlp", 64, "SmoothL1", True, per=False, clip=False, steps=2)


@pytest.mark.parametrize("gather", ["1", "0"])
def test_dqn_parity_holds_with_both_conv_data_gradient_forms(monkeypatch, gather):
    """BB_DGRAD_GATHER=1 (default): conv data gradients as ONE tcgen05 GEMM over the zero-padded dY (transposed
    convolution with separable gather operands and a separable output map); 0: dY*W + col2im."""
    monkeypatch.setenv("BB_TC", "1")
    monkeypatch.setenv("BB_DGRAD_GATHER", gather)
    from tests.test_dqn_gpu import _run
    _run("cnn", 32, "Mse", False, per=False, clip=False, steps=3, lr=1e-4)
    _run("cnn", 256, "SmoothL1", True, per=False, clip=False, steps=1, lr=1e-4)


# (mode, M, N, K): skinny contractions (one dimension <= 16: warp / thread per output element), ragged column counts,
# k not a multiple of the unroll, and split-K-heavy shapes that end in the coalesced reduce kernels
DISPATCH_SHAPES = [(0, 256, 6, 512), (0, 512, 1, 256), (0, 300, 16, 100), (0, 1, 6, 512), (0, 37, 3, 33),
                   (2, 256, 512, 6), (2, 512, 256, 1), (2, 77, 50, 16),
                   (3, 6, 512, 256), (3, 1, 256, 512), (3, 16, 70, 300), (3, 5, 33, 1000),
                   (3, 64, 100, 20000), (3, 32, 256, 40000), (0, 40, 48, 9000)]


@pytest.mark.parametrize("case", DISPATCH_SHAPES)
def test_layer_gemm_dispatcher_matches_float64(case):
    mode, M, N, K = case
    rng = np.random.default_rng(M * 131 + N * 17 + K + mode)
    if mode == 0:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    elif mode == 2:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    else:
        A, B = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    A, B = A.astype(np.float32), B.astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) if mode == 0 else None
    relu = 1 if mode == 0 and K % 2 == 0 else 0
    ref = _ref(mode, A, B, bias, relu)
    got = _run(mode, 2, A, B, M, N, K, bias, relu)
    mag = np.abs(ref).max() + 1.0
    assert np.abs(got - ref).max() <= 1.2e-5 * mag * max(1.0, (K / 4096.0) ** 0.5), (np.abs(got - ref).max(), mag)
