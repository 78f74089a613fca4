#!/bin/bash
mkdir -p gpurun_out
R=${1:-r2d}
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider --timeout=120 --timeout-method=thread -k "attention" > gpurun_out/pytest_attn_$R.log 2>&1; echo "pytest attention exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_attn_$R.log | tail -n 3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_attn_$R.log | head -20
timeout 600 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$R.json > gpurun_out/kernels_attn_$R.log 2>&1; echo "kernels exit $?"
grep attention gpurun_out/kernels_attn_$R.log | python -c "
import sys, ast
for line in sys.stdin:
    if line.startswith('{'):
        r = ast.literal_eval(line); print(r['N'], {k: round(v) for k, v in r.items() if k.endswith('tflops')})
"
bash tools/gpu_ab.sh $R "--attn-variant 5" "--attn-variant 9" "--attn-variant 9 --workload cfg2" "--attn-variant 5 --workload cfg2" "--attn-variant 9 --workload cfg5" "--attn-variant 5 --workload cfg5"
