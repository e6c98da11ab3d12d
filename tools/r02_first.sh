#!/bin/bash
# round 2, first GPU call: is the TMA-fed GEMM correct, which operand combinations work, how fast is it
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02a.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== dense TMA gemm tests"
timeout 600 python -m pytest tests/test_tma_gemm_gpu.py -q -x --timeout 300 2>&1 | tail -15
echo "== dense TMA gemm tests (continue past failures)"
timeout 600 python -m pytest tests/test_tma_gemm_gpu.py -q --timeout 300 2>&1 | tail -30
for mask in 0 1 2 4 8 16 3 31; do
  echo "== DQN cnn parity with BB_TMA_MASK=$mask"
  BB_TMA_MASK=$mask timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x -k "tensor_core_path" --timeout 300 2>&1 | tail -6
done
echo "== gemm micro"
timeout 300 python tools/gemm_micro.py
echo "== quick bench TMA"
timeout 300 python tools/quick_bench.py 65536
echo "== quick bench TMA cfg1"
BB_TMA_CFG=1 timeout 300 python tools/quick_bench.py 65536
echo "== quick bench no TMA"
BB_TMA=0 timeout 300 python tools/quick_bench.py 65536
echo "== breakdown"
timeout 300 python tools/prof_breakdown.py
