#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02g.log 2>&1
echo "== gemm + conv + dqn tests"
timeout 1500 python -m pytest tests/test_tma_gemm_gpu.py tests/test_conv_gpu.py tests/test_tc_gemm_gpu.py tests/test_dqn_gpu.py -q --timeout 600 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED|rror" | head -30
echo "== conv timing"
timeout 300 python tools/conv_timing.py 2>&1 | grep timing
echo "== gemm micro"
timeout 300 python tools/gemm_micro.py
echo "== quick bench TMA"
timeout 300 python tools/quick_bench.py 65536
echo "== breakdown"
timeout 300 python tools/prof_breakdown.py
