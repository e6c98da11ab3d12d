#!/bin/bash
# End-of-round evidence: GPU tests, both bench arms, then the ncu passes (CSV exports only).
tag=${1:-r01d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
bash tools/profile_ncu.sh ${tag} > /dev/null 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","env_steps_per_sec")}, d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline_replay"]["us_per_launch"], d["cpu_baseline"]["value"])
print(json.dumps(d["other_workloads"])[:1800])
print(open("gpurun_out/${tag}_bench_ref.json").read()[:300])
PY
du -sh gpurun_out
