#!/bin/bash
# Round-2 evidence: step breakdown (CUDA events), ncu launch list of the bench command, ncu --set full of one step's GEMMs
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_breakdown.py > gpurun_out/r02_step_breakdown_cuda_events.txt 2>&1; echo "breakdown rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launch_list.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:tma_gemm -s 60 -c 12 -f -o /tmp/r02_tma_gemm_step \
    python tools/prof_breakdown.py > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r02_tma_gemm_step.ncu-rep --page raw --csv > gpurun_out/r02_tma_gemm_step_raw.csv 2>/dev/null
ncu -i /tmp/r02_tma_gemm_step.ncu-rep --page details > gpurun_out/r02_tma_gemm_step_details.txt 2>/dev/null
# one launch (the first of the step: c2 forward) with source correlation, kept as a report
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_gemm -s 60 -c 1 -f -o gpurun_out/r02_tma_gemm_c2fwd \
    python tools/prof_breakdown.py >> gpurun_out/r02_ncu_full.log 2>&1; echo "ncu one rc=$?"
ls -la gpurun_out | head -30
du -sh gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
