#!/bin/bash
W=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 900 $TR --master-port 29542 bench.py --gpus $W --steps 100 --warmup 5 --repeats 3 --no-extra --no-cpu-baseline > gpurun_out/bench_w${W}_d$i.json 2> gpurun_out/bench_w${W}_d$i.err
  echo "bench [$envs] rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_w${W}_d$i.json'));print('N=$W',d['ms_per_step'],d['value'],d.get('ranks_bit_identical'))
for r in (d.get('exchange_trace_rank0') or [])[:2]: print(r)"
done
