#!/bin/bash
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29542 bench.py --gpus $W --steps 100 --warmup 5 --repeats 3 --no-extra --no-cpu-baseline --sync replicas > gpurun_out/bench_w${W}_nosync.json 2> gpurun_out/bench_w${W}_nosync.err
python -c "
import json;d=json.load(open('gpurun_out/bench_w${W}_nosync.json'));print('N=$W nosync',d['ms_per_step'],d['value'])"
