#!/bin/bash
# bench at N=WORLD for the given sync modes; usage: r02_mgpu2.sh WORLD mode...
W=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
for mode in "$@"; do
  BB_GRAD_SYNC=$mode timeout 900 $TR --master-port 29542 bench.py --gpus $W --steps 200 --warmup 5 > gpurun_out/bench_w${W}_$mode.json 2> gpurun_out/bench_w${W}_$mode.err
  echo "bench $mode rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_w${W}_$mode.json'));print('N=$W $mode',d['ms_per_step'],d['value'],d.get('ranks_bit_identical'),d['e2e']['value'])
for r in d.get('exchange_trace_rank0') or []: print(r)"
done
