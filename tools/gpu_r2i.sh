#!/bin/bash
# round 2, visit I: AutoencoderKL on the GPU
mkdir -p gpurun_out
R=${1:-r2i}
timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -q -p no:cacheprovider -s --timeout=300 --timeout-method=thread > gpurun_out/pytest_vae_$R.log 2>&1; echo "pytest vae exit $?"
grep -E "passed|failed|error|floor|Error|timeout|assert" gpurun_out/pytest_vae_$R.log | head -n 40
tail -n 30 gpurun_out/pytest_vae_$R.log | cut -c1-300
