#!/bin/bash
# Lean N-GPU pass: parity of the default gradient exchange, bench default + fused for comparison.
n=${1:-8}; tag=${2:-r01c}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py 2>&1 | grep -E "MGPU_OK|Error|error|assert" | head -5
for mode in default fused; do
  if [ $mode = fused ]; then export BB_GRAD_SYNC=fused; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/${tag}_bench_n${n}_${mode}.json 2> gpurun_out/${tag}_bench_n${n}_${mode}.err
  echo "n$n $mode rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_n${n}_${mode}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['config']['grad_sync'][:40], d['clocks'])"
  tail -2 gpurun_out/${tag}_bench_n${n}_${mode}.err
done
