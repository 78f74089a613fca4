#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2s}
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention and dbuf" --timeout=120 --timeout-method=thread > gpurun_out/pytest_attn_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed" gpurun_out/pytest_attn_$R.log | tail -n 2; grep -E "^FAILED" gpurun_out/pytest_attn_$R.log | head -5
python tools/attn_trace7.py --S 4608 2>&1 | tail -12
python tools/attn_trace.py --S 4608 2>&1 | grep -A3 "split-P"
