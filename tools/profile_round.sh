#!/bin/bash
# One GPU-box pass that refreshes the round's evidence: GPU tests, both bench arms, the ncu launch list of
# the bench step and full captures of the top kernels.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_ref.json
timeout 300 python tools/prof_breakdown.py > gpurun_out/${tag}_breakdown.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/prof_step.py 3 > gpurun_out/${tag}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 26 -c 13 -f \
    -o gpurun_out/${tag}_tc python tools/prof_step.py 3 > gpurun_out/${tag}_ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'replay_sample_gather|conv1_fwd|conv1_wgrad|adam_kernel|col2im' -s 18 -c 8 -f \
    -o gpurun_out/${tag}_misc python tools/prof_step.py 3 > gpurun_out/${tag}_ncu_misc.log 2>&1
ls -la gpurun_out
