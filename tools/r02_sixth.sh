#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02f.log 2>&1
for mask in 2 32 63; do
echo "== DQN mask $mask"
BB_TMA_MASK=$mask timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x -k "tensor_core_path" --timeout 300 2>&1 | grep -E "^E  .*Assert|passed|failed" | head -12
done
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED|rror" | head -30
echo "== quick bench TMA"
timeout 300 python tools/quick_bench.py 65536
echo "== breakdown"
timeout 300 python tools/prof_breakdown.py
