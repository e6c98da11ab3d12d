#!/bin/bash
# A/B of the TMEM-A tcgen05 kernel (tc_gemm5.cuh) variants
for v in 0 1; do
  echo "== BB_TC_CFG=6 BB_TC_ASYNC=$v"
  BB_TC_CFG=6 BB_TC_ASYNC=$v timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -x -q 2>&1 | tail -3
  BB_TC_CFG=6 BB_TC_ASYNC=$v timeout 200 python tools/gemm_micro.py 2>&1 | tail -9
  BB_TC_CFG=6 BB_TC_ASYNC=$v timeout 200 python tools/quick_bench.py 65536 2>&1 | grep -E "dqn opt|loss|rror"
done
