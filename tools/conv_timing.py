"""Scratch: time the AtariCnn c2 / c3 layers (forward, weight gradient, data gradient) on both tcgen05 paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["BB_CONV_ITERS"] = "30"
from tests.test_conv_gpu import _conv
rng = np.random.default_rng(0)
for (B, Cc, H, W, OC, k, s) in [(256, 32, 20, 20, 64, 4, 2), (256, 64, 9, 9, 64, 3, 1)]:
    x = rng.standard_normal((B, H, W, Cc)).astype(np.float32)
    w = rng.standard_normal((OC, k, k, Cc)).astype(np.float32)
    OH, OW = (H - k) // s + 1, (W - k) // s + 1
    dy = rng.standard_normal((B, OH, OW, OC)).astype(np.float32)
    for mode in (0, 1, 2):
        for tma in (1, 0):
            _conv(mode, tma, x, w, None, dy if mode else None, s)
