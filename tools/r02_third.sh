#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02c.log 2>&1
echo "== conv layer tests"
timeout 600 python -m pytest tests/test_conv_gpu.py -q --timeout 300 2>&1 | grep -E "^E  .*Assert|passed|failed|FAILED" | head -40
echo "== gemm micro (default)"
timeout 300 python tools/gemm_micro.py
echo "== gemm micro PASSES=1"
BB_TMA_PASSES=1 timeout 300 python tools/gemm_micro.py
echo "== gemm micro BN=128"
BB_TMA_BN=128 timeout 300 python tools/gemm_micro.py
echo "== gemm micro BN=128 PASSES=1"
BB_TMA_BN=128 BB_TMA_PASSES=1 timeout 300 python tools/gemm_micro.py
echo "== gemm micro CFG=1 (2 CTAs/SM, 2 stages)"
BB_TMA_CFG=1 timeout 300 python tools/gemm_micro.py
