#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02d.log 2>&1
echo "== conv layer tests"
timeout 600 python -m pytest tests/test_conv_gpu.py -q --timeout 300 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED" | head -40
echo "== DQN with lo check, mask 2"
BB_DEBUG_CHECK_LO=1 BB_TMA_MASK=2 timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x -s -k "tensor_core_path" --timeout 300 2>&1 | grep -E "check_lo|^E  .*Assert|passed|failed" | head -40
echo "== conv timing"
timeout 300 python tools/conv_timing.py 2>&1 | grep timing
echo "== gemm micro"
timeout 300 python tools/gemm_micro.py
echo "== ncu conv wgrad"
BB_CONV_ITERS=0 timeout 600 ncu --set full --clock-control none -k regex:tma_gemm -c 2 -o gpurun_out/r02d_wgrad python - <<'PY' 2>&1 | tail -5
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from tests.test_conv_gpu import _conv
rng = np.random.default_rng(0)
B, Cc, H, W, OC, k, s = 256, 32, 20, 20, 64, 4, 2
x = rng.standard_normal((B, H, W, Cc)).astype(np.float32)
w = rng.standard_normal((OC, k, k, Cc)).astype(np.float32)
dy = rng.standard_normal((B, 9, 9, OC)).astype(np.float32)
_conv(1, 1, x, w, None, dy, s)
_conv(0, 1, x, w, None, None, s)
PY
