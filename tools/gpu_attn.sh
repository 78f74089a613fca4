#!/bin/bash
# attention: parity of every schedule + micro-benchmark against torch SDPA (+ the clock trace of schedule 3)
mkdir -p gpurun_out
R=${1:-attn}
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention" --timeout=120 --timeout-method=thread > gpurun_out/pytest_attn_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_attn_$R.log | head -n 8
timeout 600 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$R.json > gpurun_out/kernels_attn_$R.log 2>&1; echo "kernels exit $?"
python - <<PY
import ast
for line in open("gpurun_out/kernels_attn_$R.log"):
    if line.startswith("{") and "'attention'" in line:
        r = ast.literal_eval(line); print(r["N"], {k.replace("_tflops", ""): round(v) for k, v in r.items() if k.endswith("tflops")})
PY
python tools/attn_trace.py --S 4608 2>&1 | grep -A3 "split-P"
