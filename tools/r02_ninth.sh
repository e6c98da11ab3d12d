#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02k.log 2>&1
echo "== tests"
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^E  .*|passed|failed|FAILED|rror" | head -30
echo "== quick bench"
timeout 300 python tools/quick_bench.py 65536 | grep -E "opt step|loss|flag|gather"
echo "== quick bench, materialised batch"
BB_GATHER_DIRECT=0 timeout 300 python tools/quick_bench.py 65536 | grep -E "opt step|loss|flag"
echo "== breakdown"
timeout 300 python tools/prof_breakdown.py
