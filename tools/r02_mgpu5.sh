#!/bin/bash
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
for mode in overlapped late legacy; do
  BB_GRAD_SYNC=$mode timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_check_w${W}_$mode.log 2>&1
  echo "mgpu_check $mode rc=$?"; grep MGPU_OK gpurun_out/mgpu_check_w${W}_$mode.log | tail -1
done
