#!/bin/bash
# round 2, visit N: what is the library attention kernel? ncu --set full of torch SDPA at N = 5120 next to ours
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/prof_sdpa_r2n -f python tools/sdpa_probe.py 5120 > gpurun_out/ncu_sdpa_r2n.log 2>&1; echo "ncu sdpa exit $?"
tail -n 3 gpurun_out/ncu_sdpa_r2n.log
ncu -i gpurun_out/prof_sdpa_r2n.ncu-rep --page details --csv > gpurun_out/sdpa_details_r2n.csv 2>/dev/null; wc -l gpurun_out/sdpa_details_r2n.csv
