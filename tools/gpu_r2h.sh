#!/bin/bash
# round 2, visit H: banded tile order of the wide-K GEMMs inside the real step (A/B at cfg3 and cfg5) + its DRAM traffic
mkdir -p gpurun_out
bash tools/gpu_ab.sh band3 "--m-band 0" "--m-band 2" "--m-band 5" "--m-band 8"
bash tools/gpu_ab.sh band5 "--workload cfg5 --steps 10 --m-band 0" "--workload cfg5 --steps 10 --m-band 3" "--workload cfg5 --steps 10 --m-band 5"
for B in 0 5; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:tcgen05|ln_modulate" --csv --log-file gpurun_out/launches_r2h_cfg3_band$B.csv \
   python tools/one_step.py --workload cfg3 --m-band $B > gpurun_out/ncu_launches_r2h_$B.log 2>&1; echo "ncu band $B exit $?"
python tools/traffic_from_ncu.py gpurun_out/launches_r2h_cfg3_band$B.csv gpurun_out/traffic_r2h_cfg3_band$B.json
done
