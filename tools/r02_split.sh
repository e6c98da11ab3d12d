#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02_split.log 2>&1
for cfg in "50 16" "100 16" "100 32" "200 32" "200 64"; do
  set -- $cfg
  echo "== FILL=$1 CAP=$2"
  BB_TMA_FILL=$1 BB_TMA_SPLIT_CAP=$2 timeout 300 python tools/conv_timing.py 2>&1 | grep "timing" | grep "tma 1" | grep "mode 1"
  BB_TMA_FILL=$1 BB_TMA_SPLIT_CAP=$2 ONLY=l1.fwd,c2.wgrad timeout 300 python tools/gemm_micro.py 2>&1 | tail -2
  BB_TMA_FILL=$1 BB_TMA_SPLIT_CAP=$2 timeout 300 python tools/quick_bench.py 65536 | grep -E "opt step"
done
