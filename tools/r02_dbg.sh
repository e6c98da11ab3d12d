#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02i.log 2>&1
echo "== micro c1.fwd cfg1"
BB_TMA_CFG=1 ONLY=c1.fwd timeout 120 python tools/gemm_micro.py 2>&1 | tail -2
echo "== micro c1.fwd passes1"
BB_TMA_PASSES=1 ONLY=c1.fwd timeout 120 python tools/gemm_micro.py 2>&1 | tail -2
echo "== micro c1.fwd default again"
ONLY=c1.fwd timeout 120 python tools/gemm_micro.py 2>&1 | tail -2
echo "== micro shapes N=32 varying M"
python - <<'PY' 2>&1 | tail -12
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from border_b200 import _lib as L
lib = L.lib()
for M in (128, 1024, 8192, 20000, 40000, 102400):
    ms = C.c_float()
    try:
        L.check(lib.bb_bench_gemm(0, 0, 3, M, 32, 256, 20, C.byref(ms)))
        e = C.c_int32(); lib.bb_device_error(C.byref(e), 1)
        print("M", M, "ok %.1f us" % (ms.value * 1e3), "flag", e.value)
    except Exception as ex:
        print("M", M, "FAIL", str(ex)[-80:]); break
PY
echo "== determinism of DQN loss"
for i in 1 2 3; do timeout 300 python tools/quick_bench.py 65536 2>&1 | grep -E "loss|error|opt step"; done
