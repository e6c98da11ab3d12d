"""Scratch: per-k-slice clock64 timeline of CTA 0 of one TMA GEMM (BB_TMA_TRACE=1): where does a slice's time go?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BB_TMA_TRACE"] = "1"
import numpy as np
from border_b200 import _lib as L
lib = L.lib()
mode, M, N, K = [int(x) for x in sys.argv[1:5]]
ms = C.c_float()
L.check(lib.bb_bench_gemm(0, mode, 3, M, N, K, 2, C.byref(ms)))
t = np.zeros((3, 64, 4), np.int64)
L.check(lib.bb_tma_trace(t.ctypes.data))
t0 = t[0, 0, 0]
n = min(64, (K + 31) // 32)
print("slice | producer: wait_empty issue done | mma: wait ready issued | split: wait full loaded stored arrived   (cycles since start)")
for i in range(min(n, 24)):
    r = lambda a: (a - t0) if a else -1
    print("%3d | %6d %6d %6d | %6d %6d %6d | %6d %6d %6d %6d" % (i, r(t[0, i, 0]), r(t[0, i, 1]), r(t[0, i, 2]), r(t[1, i, 0]), r(t[1, i, 1]), r(t[1, i, 2]),
                                                             r(t[2, i, 0]), r(t[2, i, 1]), r(t[2, i, 2]), r(t[2, i, 3])))
print("kernel body start %d | epilogue: wait accum %d, accum ready %d, stores done %d, after final sync %d" % (t[1, 63, 0] - t0, t[2, 63, 0] - t0, t[2, 63, 1] - t0, t[2, 63, 2] - t0, t[2, 63, 3] - t0))
print("ms per GEMM (2 iterations, traced): %.4f" % ms.value)
