#!/bin/bash
# ncu evidence only (launch list + full captures exported to CSV on the box; the .ncu-rep files are too big
# to travel back through gpurun_out's 64 MiB cap, so only the small ones are kept).
tag=${1:-r01b}
mkdir -p gpurun_out
timeout 300 python tools/prof_breakdown.py > gpurun_out/${tag}_breakdown.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/prof_step.py 3 > gpurun_out/${tag}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 26 -c 13 -f \
    -o /tmp/${tag}_tc python tools/prof_step.py 3 > gpurun_out/${tag}_ncu_tc.log 2>&1
ncu -i /tmp/${tag}_tc.ncu-rep --page raw --csv > gpurun_out/${tag}_tc_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_tc.ncu-rep --page source --csv -c 1 > gpurun_out/${tag}_tc_src_c2fwd.csv 2>/dev/null
ncu -i /tmp/${tag}_tc.ncu-rep --page details --csv -c 1 > gpurun_out/${tag}_tc_details_c2fwd.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'replay_sample_gather|conv1_fwd|conv1_wgrad|adam_kernel|col2im' -s 18 -c 8 -f \
    -o /tmp/${tag}_misc python tools/prof_step.py 3 > gpurun_out/${tag}_ncu_misc.log 2>&1
ncu -i /tmp/${tag}_misc.ncu-rep --page raw --csv > gpurun_out/${tag}_misc_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
