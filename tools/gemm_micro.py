"""Scratch: time single GEMM shapes through the SIMT / tcgen05 paths (device-resident operands)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
lib.bb_bench_gemm.restype = C.c_int32
lib.bb_bench_gemm.argtypes = [C.c_int32] * 7 + [C.POINTER(C.c_float)]
shapes = [("c1.fwd", 0, 102400, 32, 256), ("c2.fwd", 0, 20736, 64, 512), ("c3.fwd", 0, 12544, 64, 576), ("l1.fwd", 0, 256, 512, 3136),
          ("l1.dgrad", 2, 256, 3136, 512), ("c2.dgrad", 2, 20736, 512, 64), ("c2.wgrad", 3, 64, 512, 20736), ("l1.wgrad", 3, 512, 3136, 256), ("big", 0, 8192, 8192, 1024)]
only = os.environ.get("ONLY")
for name, mode, M, N, K in shapes:
    if only and name not in only.split(","): continue
    out = []
    for tc in (0, 1, 3):
        ms = C.c_float()
        L.check(lib.bb_bench_gemm(0, mode, tc, M, N, K, 50, C.byref(ms)))
        out.append(ms.value)
    fl = 2.0 * M * N * K
    print("%-9s M=%6d N=%5d K=%6d  simt %8.1f us (%6.1f TF)   tc %8.1f us (%6.1f TF)   tma %8.1f us (%6.1f TF)" % (name, M, N, K, out[0] * 1e3, fl / out[0] / 1e9, out[1] * 1e3, fl / out[1] / 1e9, out[2] * 1e3, fl / out[2] / 1e9))
