#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r02_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["env_steps"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
print({k:(v.get("us_per_step") or v.get("ms_per_step")) for k,v in d["other_workloads"].items() if isinstance(v,dict)})
PY
