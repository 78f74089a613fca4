#!/bin/bash
# round 2, visit G: unfused-LoRA side path + banded GEMM tile order: tests, micro-benchmark
mkdir -p gpurun_out
R=${1:-r2g}
timeout 900 python -m pytest tests/test_gpu_lora.py tests/test_gpu_loader.py -m gpu -q -p no:cacheprovider -s --timeout=300 --timeout-method=thread > gpurun_out/pytest_lora_$R.log 2>&1; echo "pytest lora exit $?"
grep -E "passed|failed|error|config-4|fold noise|ulp-sized" gpurun_out/pytest_lora_$R.log | tail -n 12; grep -E "^FAILED|^ERROR|Error" gpurun_out/pytest_lora_$R.log | head -20
timeout 600 python tools/bench_kernels.py --only band --json gpurun_out/kernels_band_$R.json > gpurun_out/kernels_band_$R.log 2>&1; echo "kernels exit $?"
python - <<'PY'
import ast
for line in open("gpurun_out/kernels_band_r2g.log"):
    if line.startswith("{"):
        r = ast.literal_eval(line); print(r["name"], r["M"], r["N"], r["K"], {k.replace("_tflops", ""): round(v) for k, v in r.items() if k.endswith("tflops")})
PY
tail -n 5 gpurun_out/kernels_band_$R.log | cut -c1-300
