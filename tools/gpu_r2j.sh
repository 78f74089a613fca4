#!/bin/bash
# round 2, visit J: VAE tests + drop-in with the VAE attached + VAE benchmark against CUDA eager
mkdir -p gpurun_out
R=${1:-r2j}
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_dropin.py -m gpu -q -p no:cacheprovider -s --timeout=300 --timeout-method=thread > gpurun_out/pytest_vae_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error|floor|attached:|^FAILED|^ERROR" gpurun_out/pytest_vae_$R.log | head -n 40
timeout 900 python tools/bench_vae.py --json gpurun_out/bench_vae_$R.json > gpurun_out/bench_vae_$R.log 2>&1; echo "bench exit $?"
tail -n 8 gpurun_out/bench_vae_$R.log | cut -c1-600
