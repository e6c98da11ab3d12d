"""Scratch: clock64 trace of CTA (0,0,0) of the default tcgen05 kernel (BB_TC_DEBUG=16 [+ floor bits])."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
mode = int(sys.argv[1]); M, N, K = [int(x) for x in sys.argv[2:5]]
ms = C.c_float()
L.check(lib.bb_bench_gemm(0, mode, 1, M, N, K, 3, C.byref(ms)))
t = np.zeros((2, 64, 8), np.int64)
L.check(lib.bb_debug_tc_trace(t.ctypes.data))
t0 = t[1, 56, 0]
print("shape", mode, M, N, K, "DEBUG", os.environ.get("BB_TC_DEBUG"), "ms %.4f" % ms.value)
print("cta (cycles from start): after_setup %d, producer_loop_done %d, accum_ready %d, epilogue_done %d, exit %d"
      % tuple(int(t[1, i, 0] - t0) for i in (57, 58, 59, 60, 61)))
for ks in range(min(20, (K + 31) // 32)):
    a = t[0, ks] - t0; m = t[1, ks] - t0
    print("ks %2d prod: store_begin %6d empty_ok %6d arrived %6d | mma: wait_begin %6d full_ok %6d committed %6d"
          % (ks, a[0], a[2], a[3], m[0], m[1], m[3]))
