#!/bin/bash
# multi-GPU check + bench; usage: r02_mgpu.sh WORLD
W=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
for mode in overlapped late legacy; do
  BB_GRAD_SYNC=$mode timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_check_w${W}_$mode.log 2>&1
  echo "mgpu_check $mode rc=$?"; grep MGPU_OK gpurun_out/mgpu_check_w${W}_$mode.log | tail -3
done
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_w1.json 2> gpurun_out/bench_w1.err
python -c "import json;d=json.load(open('gpurun_out/bench_w1.json'));print('N=1',d['ms_per_step'],d['value'])"
for mode in overlapped late legacy; do
  BB_GRAD_SYNC=$mode timeout 900 $TR --master-port 29542 bench.py --gpus $W --steps 200 --warmup 5 > gpurun_out/bench_w${W}_$mode.json 2> gpurun_out/bench_w${W}_$mode.err
  echo "bench $mode rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_w${W}_$mode.json'));print('N=$W $mode',d['ms_per_step'],d['value'],d.get('ranks_bit_identical'),d['e2e']['value'])"
done
