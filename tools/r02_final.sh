#!/bin/bash
# End-of-round evidence at N=1: GPU tests, smoke, both bench arms, step breakdown, ncu launch list + full capture of one step's GEMMs
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_n1.err; echo "ref rc=$?"
timeout 300 python tools/prof_breakdown.py > gpurun_out/r02_step_breakdown_cuda_events.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launch_list.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:tma_gemm -s 60 -c 12 -f -o /tmp/r02_tma_gemm_step \
    python tools/prof_breakdown.py > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r02_tma_gemm_step.ncu-rep --page raw --csv > gpurun_out/r02_tma_gemm_step_raw.csv 2>/dev/null
ncu -i /tmp/r02_tma_gemm_step.ncu-rep --page details > gpurun_out/r02_tma_gemm_step_details.txt 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["env_steps"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("traffic"), d["cpu_baseline"]["value"], d["clocks"])
print(json.dumps(d["other_workloads"])[:2500])
print(open("gpurun_out/r02_bench_reference_arm.json").read()[:400])
PY
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
