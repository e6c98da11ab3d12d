#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02b.log 2>&1
echo "== probe"
timeout 120 ./build/tma_probe
echo "== dense TMA gemm tests"
timeout 600 python -m pytest tests/test_tma_gemm_gpu.py -q --timeout 300 2>&1 | grep -E "passed|failed|FAILED|Error" | head -30
for mask in 2 32 4 8 16 63; do
  echo "== DQN cnn parity with BB_TMA_MASK=$mask"
  BB_TMA_MASK=$mask timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x -k "tensor_core_path" --timeout 300 2>&1 | grep -E "^E  |passed|failed" | head -12
done
echo "== gemm micro"
timeout 300 python tools/gemm_micro.py
echo "== quick bench TMA"
timeout 300 python tools/quick_bench.py 65536
echo "== breakdown"
timeout 300 python tools/prof_breakdown.py
