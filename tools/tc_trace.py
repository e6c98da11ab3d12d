"""Scratch: per-stage clock64 trace of the cp.async tcgen05 kernel (run with BB_TC_CFG=5 BB_TC_DEBUG=16[+bits])."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
M, N, K = [int(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else (8192, 8192, 1024)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ms = C.c_float()
L.check(lib.bb_bench_gemm(0, mode, 1, M, N, K, 3, C.byref(ms)))
t = np.zeros((2, 64, 8), np.int64)
L.check(lib.bb_debug_tc_trace_variants(t.ctypes.data))
t0 = t[t > 0].min()
print("env", {k: v for k, v in os.environ.items() if k.startswith("BB_")}, "ms", ms.value)
print("prod order: pre_wait after_cpwait after_empty after_A after_B after_issue after_waitst after_arrive\nks | prod: pre_wait after_cpwait after_empty after_arrive | mma: pre_full after_full after_fence after_commit")
print("cta: start, after_setup, mainloop_done, accum_ready, epilogue_done, exit:", (t[1, 56:62, 0] - t0).tolist())
for ks in range(min(32, (K + 31) // 32)):
    a = t[0, ks] - t0
    print(ks, "prod", a[[0, 1, 2, 4, 5, 6, 7, 3]].tolist(), "| d:", np.diff(a[[0, 1, 2, 4, 5, 6, 7, 3]]).tolist(), "| mma", (t[1, ks, :4] - t0).tolist())
