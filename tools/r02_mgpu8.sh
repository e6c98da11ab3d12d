#!/bin/bash
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
for mode in overlapped legacy; do
  BB_GRAD_SYNC=$mode timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_check_w${W}_$mode.log 2>&1
  echo "mgpu_check $mode rc=$?"; grep MGPU_OK gpurun_out/mgpu_check_w${W}_$mode.log | tail -1
done
tools/r02_mgpu4.sh $W "BB_GRAD_SYNC=overlapped BB_XCHG_EARLY_BLOCKS=64" "BB_GRAD_SYNC=overlapped BB_XCHG_EARLY_BLOCKS=32" "BB_GRAD_SYNC=legacy"
python bench.py --steps 100 --warmup 5 --repeats 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'])"
