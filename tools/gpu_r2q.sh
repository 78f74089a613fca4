#!/bin/bash
# round 2, visit Q: schedule 6 (row-split softmax over 20 warps): parity, micro-benchmark vs SDPA
mkdir -p gpurun_out
R=${1:-r2q}
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention and (dbuf or deterministic)" --timeout=120 --timeout-method=thread > gpurun_out/pytest_attn_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error|^FAILED|^ERROR|Timeout|timeout" gpurun_out/pytest_attn_$R.log | head -n 20
timeout 600 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$R.json > gpurun_out/kernels_attn_$R.log 2>&1; echo "kernels exit $?"
python - <<PY
import ast
for line in open("gpurun_out/kernels_attn_$R.log"):
    if line.startswith("{") and "'attention'" in line:
        r = ast.literal_eval(line); print(r["N"], {k.replace("_tflops", ""): round(v) for k, v in r.items() if k.endswith("tflops")})
PY
tail -n 3 gpurun_out/kernels_attn_$R.log | cut -c1-300
