#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lora.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_snake.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_snake.log
bash tools/gpu_ab.sh sn3 "--k-snake 0" "--k-snake 1" "--k-snake 0" "--k-snake 1"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:tcgen05|ln_modulate" --csv --log-file gpurun_out/launches_snake_cfg3.csv \
   python tools/one_step.py --workload cfg3 --k-snake 1 > /dev/null 2>&1; echo "ncu exit $?"
python tools/traffic_from_ncu.py gpurun_out/launches_snake_cfg3.csv gpurun_out/traffic_snake_cfg3.json
