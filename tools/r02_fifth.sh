#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02e.log 2>&1
for mask in 0 2; do
echo "== DQN with lo check, mask $mask"
BB_DEBUG_CHECK_LO=1 BB_TMA_MASK=$mask timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x -s -k "tensor_core_path" --timeout 300 2>&1 | grep -E "check_lo|^E  .*Assert|passed|failed" | head -12
done
echo "== serial streams, mask 2"
BB_SERIAL=1 BB_TMA_MASK=2 timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x -s -k "tensor_core_path" --timeout 300 2>&1 | grep -E "^E  .*Assert|passed|failed" | head -12
