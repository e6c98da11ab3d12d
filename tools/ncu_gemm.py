"""Scratch: one launch of a dense GEMM shape through the TMA path (for ncu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
mode, M, N, K = [int(x) for x in sys.argv[1:5]]
ms = C.c_float()
L.check(lib.bb_bench_gemm(0, mode, 3, M, N, K, 1, C.byref(ms)))
print(ms.value)
