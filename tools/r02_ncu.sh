#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
exec > gpurun_out/r02h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_gemm -s 3 -c 1 -o gpurun_out/r02h_big python tools/ncu_gemm.py 0 8192 8192 1024 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_gemm -s 3 -c 1 -o gpurun_out/r02h_c2 python tools/ncu_gemm.py 0 20736 64 512 2>&1 | tail -3
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
