"""Scratch: fixed-overhead probe of the tcgen05 kernel (tiny shapes)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
shapes = [("l2.dgrad", 2, 256, 512, 6), ("k32", 0, 256, 512, 32), ("1tile-k32", 0, 128, 64, 32), ("1tile-k1024", 0, 128, 64, 1024),
          ("1tile-k4096", 0, 128, 64, 4096), ("148tiles-k32", 0, 18944, 64, 32), ("148tiles-k512", 0, 18944, 64, 512),
          ("296tiles-k512", 0, 37888, 64, 512), ("l2.fwd", 0, 256, 6, 512), ("l2.wgrad", 3, 6, 512, 256)]
for name, mode, M, N, K in shapes:
    out = []
    for tc in (0, 1):
        ms = C.c_float()
        L.check(lib.bb_bench_gemm(0, mode, tc, M, N, K, 50, C.byref(ms)))
        out.append(ms.value)
    print("%-14s M=%6d N=%5d K=%6d  simt %8.1f us   tc %8.1f us" % (name, M, N, K, out[0] * 1e3, out[1] * 1e3))
