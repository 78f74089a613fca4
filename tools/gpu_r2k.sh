#!/bin/bash
# round 2, visit K: prompt encoders on the GPU
mkdir -p gpurun_out
R=${1:-r2k}
timeout 900 python -m pytest tests/test_gpu_textenc.py -m gpu -q -p no:cacheprovider -s --timeout=300 --timeout-method=thread > gpurun_out/pytest_textenc_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error|floor|^FAILED|^ERROR|Error" gpurun_out/pytest_textenc_$R.log | head -n 40
