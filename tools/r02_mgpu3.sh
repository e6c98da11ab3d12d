#!/bin/bash
# bench at N=WORLD over env settings; usage: r02_mgpu3.sh WORLD "ENV=.. ENV=.." ...
W=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 900 $TR --master-port 29542 bench.py --gpus $W --steps 200 --warmup 5 > gpurun_out/bench_w${W}_v$i.json 2> gpurun_out/bench_w${W}_v$i.err
  echo "bench [$envs] rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_w${W}_v$i.json'));print('N=$W',d['ms_per_step'],d['value'],d.get('ranks_bit_identical'),d['e2e']['value'])
for r in (d.get('exchange_trace_rank0') or [])[:3]: print(r)"
done
