#!/bin/bash
# 2-GPU pass: multi-GPU parity test, the bench under torchrun (both arms), then the 1-GPU bench line.
tag=${1:-r01b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mgpu_gpu.py -x -q 2>&1 | tail -3
for sync in allreduce replicas; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 20 --sync $sync > gpurun_out/${tag}_bench_n2_${sync}.json 2> gpurun_out/${tag}_bench_n2_${sync}.err
echo "n2 $sync rc=$?"; cat gpurun_out/${tag}_bench_n2_${sync}.json; tail -3 gpurun_out/${tag}_bench_n2_${sync}.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_ref_n2.json 2> gpurun_out/${tag}_bench_ref_n2.err
echo "ref n2 rc=$?"; cat gpurun_out/${tag}_bench_ref_n2.json
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
