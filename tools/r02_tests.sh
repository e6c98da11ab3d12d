#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02_tests.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -E "^E  .*|passed|failed|FAILED|rror|relative error" | head -60
