#!/bin/bash
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_check_w${W}_overlapped.log 2>&1
echo "mgpu_check rc=$?"; grep MGPU_OK gpurun_out/mgpu_check_w${W}_overlapped.log | tail -1
tools/r02_mgpu4.sh $W "BB_XCHG_EARLY_BLOCKS=32" "BB_XCHG_EARLY_BLOCKS=32 BB_XCHG_MID=0" "BB_XCHG_EARLY_BLOCKS=16" "BB_XCHG_DBG=2" | grep -v step_ns
tools/r02_mgpu6.sh $W
