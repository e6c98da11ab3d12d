#!/bin/bash
# N-GPU pass: multi-GPU parity in both gradient-exchange modes, then the bench in both.
n=${1:-2}; tag=${2:-r01c}
mkdir -p gpurun_out
for mode in fused sharded; do
  echo "== parity BB_GRAD_SYNC=$mode"
  BB_GRAD_SYNC=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_check.py 2>&1 | grep -E "MGPU_OK|Error|error|assert" | head -5
done
for mode in fused sharded; do
  BB_GRAD_SYNC=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/${tag}_bench_n${n}_${mode}.json 2> gpurun_out/${tag}_bench_n${n}_${mode}.err
  echo "n$n $mode rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_n${n}_${mode}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['config']['grad_sync'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus $n --steps 200 --warmup 20 --sync replicas > gpurun_out/${tag}_bench_n${n}_replicas.json 2> gpurun_out/${tag}_bench_n${n}_replicas.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_n${n}_replicas.json').read().strip().splitlines()[-1])
print('replicas', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])"
head -c 300 gpurun_out/${tag}_bench_n${n}_fused.json; echo
