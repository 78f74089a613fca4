#!/bin/bash
# round 2, visit M: lag-1 lazy maximum in the attention softmax: parity, micro-benchmark vs SDPA, A/B inside the step
mkdir -p gpurun_out
R=${1:-r2m}
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -s -k "attention" --timeout=300 --timeout-method=thread > gpurun_out/pytest_attn_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error|^FAILED|^ERROR|jump|ramp|first_tile" gpurun_out/pytest_attn_$R.log | tail -n 30
timeout 600 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$R.json > gpurun_out/kernels_attn_$R.log 2>&1; echo "kernels exit $?"
python - <<'PY'
import ast
for line in open("gpurun_out/kernels_attn_r2m.log"):
    if line.startswith("{") and "'attention'" in line:
        r = ast.literal_eval(line); print(r["N"], {k.replace("_tflops", ""): round(v) for k, v in r.items() if k.endswith("tflops")})
PY
bash tools/gpu_ab.sh lag3 "--attn-lag 0" "--attn-lag 1"
bash tools/gpu_ab.sh lag5 "--workload cfg5 --steps 10 --attn-lag 0" "--workload cfg5 --steps 10 --attn-lag 1"
bash tools/gpu_ab.sh lag2 "--workload cfg2 --attn-lag 0" "--workload cfg2 --attn-lag 1"
