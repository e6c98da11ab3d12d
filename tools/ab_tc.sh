#!/bin/bash
# A/B of the tcgen05 kernel configurations (scratch tuning aid)
for fm in 1 0; do for c in 2 3; do
  echo "== BB_TC_FENCE=$fm BB_TC_CFG=$c"
  BB_TC=1 BB_TC_FENCE=$fm BB_TC_CFG=$c timeout 200 python tools/quick_bench.py 131072 2>&1 | grep -E "dqn opt|loss|rror"
done; done
