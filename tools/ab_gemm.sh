#!/bin/bash
# A/B of the GEMM register-tile variants and split-K fill targets (scratch tuning aid)
for v in 0 1 2 3; do for f in 2 3 4; do
  echo "== variant $v fill $f"
  BB_GEMM_VARIANT=$v BB_GEMM_FILL=$f python tools/quick_bench.py 131072 2>&1 | grep -E "dqn opt|loss"
done; done
