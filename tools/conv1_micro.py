"""Scratch: time the dedicated conv1 forward kernel alone (BB_CONV1_DEBUG bits: 1 no MMA, 2 no loads, 4 no TMEM stores)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from border_b200 import _lib as L
lib = L.lib()
ms = C.c_float()
L.check(lib.bb_bench_conv1(0, 256, 4, 200, C.byref(ms)))
print("conv1 fwd B=256: %.2f us  (debug=%s)" % (ms.value * 1e3, os.environ.get("BB_CONV1_DEBUG")))
