"""Scratch: per-k-slice clock64 timeline of CTA 0 of a TMA conv GEMM (forward = im2col A operand): c2 / c3 of the AtariCnn."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BB_TMA_TRACE"] = "1"
os.environ["BB_CONV_ITERS"] = "1"
import numpy as np
from border_b200 import _lib as L
from tests.test_conv_gpu import _conv
rng = np.random.default_rng(0)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for (B, Cc, H, W, OC, k, s) in [(256, 32, 20, 20, 64, 4, 2), (256, 64, 9, 9, 64, 3, 1)]:
    x = rng.standard_normal((B, H, W, Cc)).astype(np.float32)
    w = rng.standard_normal((OC, k, k, Cc)).astype(np.float32)
    OH, OW = (H - k) // s + 1, (W - k) // s + 1
    dy = rng.standard_normal((B, OH, OW, OC)).astype(np.float32)
    _conv(mode, 1, x, w, None, dy if mode else None, s)
    t = np.zeros((3, 64, 4), np.int64)
    L.check(L.lib().bb_tma_trace(t.ctypes.data))
    t0 = t[0, 0, 0]
    print("conv C=%d k=%d s=%d mode %d" % (Cc, k, s, mode))
    print("slice | producer: wait_empty issue done | mma: wait ready issued | split: wait full loaded stored arrived")
    for i in range(18):
        r = lambda a: (a - t0) if a else -1
        print("%3d | %6d %6d %6d | %6d %6d %6d | %6d %6d %6d %6d" % (i, r(t[0, i, 0]), r(t[0, i, 1]), r(t[0, i, 2]), r(t[1, i, 0]), r(t[1, i, 1]), r(t[1, i, 2]),
                                                                 r(t[2, i, 0]), r(t[2, i, 1]), r(t[2, i, 2]), r(t[2, i, 3])))
    c = np.zeros((1024, 4), np.int64)
    L.check(L.lib().bb_tma_trace_ctas(c.ctypes.data))
    c = c[c[:, 2] > 0]
    g0 = c[:, 0].min()
    per_sm = {}
    for row in c:
        per_sm.setdefault(int(row[3]), []).append(((row[0] - g0) / 1e3, (row[1] - g0) / 1e3, (row[2] - g0) / 1e3))
    n2 = sum(1 for v in per_sm.values() if len(v) > 1)
    print("CTAs %d on %d SMs (%d SMs with >1); entry us min/max %.2f/%.2f; exit us min/median/max %.2f/%.2f/%.2f; life us median %.2f max %.2f" % (
        len(c), len(per_sm), n2, (c[:, 0] - g0).min() / 1e3, (c[:, 0] - g0).max() / 1e3, (c[:, 2] - g0).min() / 1e3,
        np.median(c[:, 2] - g0) / 1e3, (c[:, 2] - g0).max() / 1e3, np.median(c[:, 2] - c[:, 1]) / 1e3, (c[:, 2] - c[:, 1]).max() / 1e3))
    life = (c[:, 2] - c[:, 1]) / 1e3
    print("life histogram (us):", np.histogram(life, bins=8)[0].tolist(), np.round(np.histogram(life, bins=8)[1], 1).tolist())
    print("epilogue phase A done %d, after bar %d" % (t[2, 62, 0] - t0, t[2, 62, 1] - t0))
    print("body start %d | epilogue: wait accum %d, accum ready %d, stores done %d, after final sync %d" % (t[1, 63, 0] - t0, t[2, 63, 0] - t0, t[2, 63, 1] - t0, t[2, 63, 2] - t0, t[2, 63, 3] - t0))
