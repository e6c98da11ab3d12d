"""Scratch: component floors of the cp.async tcgen05 kernel via BB_TC_DEBUG bits (1 no loads, 2 no MMA, 4 no split stores, 8 one MMA pass)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
lib.bb_bench_gemm.restype = C.c_int32
lib.bb_bench_gemm.argtypes = [C.c_int32] * 7 + [C.POINTER(C.c_float)]
shapes = [("c1.fwd", 0, 102400, 32, 256), ("c2.fwd", 0, 20736, 64, 512), ("l1.fwd", 0, 256, 512, 3136),
          ("c2.dgrad", 2, 20736, 512, 64), ("c2.wgrad", 3, 64, 512, 20736), ("big", 0, 8192, 8192, 1024)]
out = []
for name, mode, M, N, K in shapes:
    ms = C.c_float()
    L.check(lib.bb_bench_gemm(0, mode, 1, M, N, K, 30, C.byref(ms)))
    out.append("%s %.1f" % (name, ms.value * 1e3))
print("DEBUG=%s ASYNC=%s CFG=%s: " % (os.environ.get("BB_TC_DEBUG"), os.environ.get("BB_TC_ASYNC"), os.environ.get("BB_TC_CFG")) + "  ".join(out))
