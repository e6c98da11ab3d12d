#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
for envs in "BB_TMA_CFG=0" "BB_TMA_CFG=-1" "BB_TMA_CFG=-1 BB_TMA_FILL=75" "BB_TMA_CFG=-1 BB_TMA_FILL=100" "BB_TMA_CFG=-1 BB_TMA_FILL=100 BB_TMA_SPLIT_CAP=32" "BB_TMA_CFG=0"; do
  echo "== $envs"; env $envs python tools/quick_bench.py 65536 2>&1 | grep "dqn opt"
done
BB_TMA_CFG=-1 python tools/gemm_micro.py 2>&1 | tail -9
