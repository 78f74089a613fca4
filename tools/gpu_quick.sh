#!/bin/bash
# quick visit: full gpu suite, GEMM micro-benchmark (first 4 shapes), headline bench
mkdir -p gpurun_out
R=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=120 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_kernels.py --only gemm --json gpurun_out/kernels_gemm_$R.json 2>&1 | python -c "
import sys, ast
for line in sys.stdin:
    if line.startswith('{'):
        r = ast.literal_eval(line)
        print(r['name'], {k: round(v) for k, v in r.items() if k.endswith('tflops')})
"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$R.json')); print(round(d['value'],3), round(d['ms_per_step'],3), d['clocks'], d['finite'], d['roofline']['kernel_families_us'], d['e2e'])"
tail -n 3 gpurun_out/bench_$R.err
