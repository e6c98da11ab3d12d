#!/bin/bash
# End-of-round multi-GPU numbers: bench at N GPUs in the default exchange mode (+ replicas when asked).
n=${1:-8}; tag=${2:-r01d}; shift 2
mkdir -p gpurun_out
for mode in default "$@"; do
  extra=""; if [ "$mode" = replicas ]; then extra="--sync replicas"; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $n --steps 200 --warmup 20 $extra > gpurun_out/${tag}_bench_n${n}_${mode}.json 2> gpurun_out/${tag}_bench_n${n}_${mode}.err
  echo "n$n $mode rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_n${n}_${mode}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['env_steps_per_sec'], d['config']['grad_sync'][:30], d['clocks'])"
done
