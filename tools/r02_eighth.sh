#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02j.log 2>&1
echo "== gemm micro BN=128 where possible"
timeout 300 python tools/gemm_micro.py
echo "== gemm micro BN=64"
BB_TMA_BN=64 timeout 300 python tools/gemm_micro.py
echo "== quick bench (BN128)"
timeout 300 python tools/quick_bench.py 65536 | grep -E "opt step|loss|flag"
echo "== quick bench (BN64)"
BB_TMA_BN=64 timeout 300 python tools/quick_bench.py 65536 | grep -E "opt step|loss|flag"
echo "== tests"
timeout 1500 python -m pytest tests/test_tma_gemm_gpu.py tests/test_conv_gpu.py tests/test_tc_gemm_gpu.py tests/test_dqn_gpu.py -q --timeout 600 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED|rror" | head -30
