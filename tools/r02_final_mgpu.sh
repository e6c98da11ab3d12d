#!/bin/bash
# End-of-round multi-GPU evidence; usage: r02_final_mgpu.sh WORLD [check]
cd "$(dirname "$0")/.." || exit 1
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
if [ "$2" = "check" ]; then
  timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_check_w${W}_overlapped.log 2>&1
  echo "mgpu_check rc=$?"; grep MGPU_OK gpurun_out/mgpu_check_w${W}_overlapped.log | tail -1
fi
timeout 900 $TR --master-port 29542 bench.py --gpus $W --steps 200 --warmup 5 > gpurun_out/r02_bench_n$W.json 2> gpurun_out/r02_bench_n$W.err
echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_n$W.json'));print('N=$W',d['ms_per_step'],d['value'],d.get('ranks_bit_identical'),d['e2e']['value'],d['env_steps_per_sec'])"
