#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
for rep in 1 2; do
for envs in "BB_DGRAD_DIRECT=0 BB_DGRAD_WT_SIDE=0" "BB_DGRAD_DIRECT=0 BB_DGRAD_WT_SIDE=1" "BB_DGRAD_DIRECT=1 BB_DGRAD_WT_SIDE=1" "BB_DGRAD_DIRECT=1 BB_DGRAD_WT_SIDE=0"; do
  echo -n "$envs: "; env $envs python tools/quick_bench.py 65536 2>&1 | grep "dqn opt"
done
done
