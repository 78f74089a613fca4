#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lora.py -m gpu -q -p no:cacheprovider -k "banded" > gpurun_out/pytest_nband.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/pytest_nband.log
bash tools/gpu_ab.sh nb3 "--m-band -1" "--m-band -106" "--m-band -104" "--m-band -106"
bash tools/gpu_ab.sh nb5 "--workload cfg5 --steps 10 --m-band -1" "--workload cfg5 --steps 10 --m-band -106"
for B in -106; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:tcgen05|ln_modulate" --csv --log-file gpurun_out/launches_nband_cfg3.csv \
   python tools/one_step.py --workload cfg3 --m-band $B > /dev/null 2>&1; echo "ncu exit $?"
python tools/traffic_from_ncu.py gpurun_out/launches_nband_cfg3.csv gpurun_out/traffic_nband_cfg3.json
done
