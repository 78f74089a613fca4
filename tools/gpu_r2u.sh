#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2u}
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_dropin.py -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread -k "vae" > gpurun_out/pytest_vae_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_vae_$R.log | head
timeout 900 python tools/bench_vae.py --json gpurun_out/bench_vae_$R.json --sizes 1024x1152,1024x2048 > gpurun_out/bench_vae_$R.log 2>&1; echo "bench exit $?"
tail -n 3 gpurun_out/bench_vae_$R.log | cut -c1-420
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_${R}_decode.csv python tools/one_vae.py --what decode > /dev/null 2>&1; echo "ncu exit $?"
