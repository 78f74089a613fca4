#!/bin/bash
# round 2, visit B: whole gpu suite (no -x), kernel micro-benchmarks with the stream attention schedule, headline bench
mkdir -p gpurun_out
R=${1:-r2b}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread -s > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$R.log | tail -n 3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_$R.log | head -20
timeout 600 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$R.json > gpurun_out/kernels_attn_$R.log 2>&1; echo "kernels exit $?"
grep attention gpurun_out/kernels_attn_$R.log | python -c "
import sys, ast
for line in sys.stdin:
    if line.startswith('{'):
        r = ast.literal_eval(line); print(r['N'], {k: round(v) for k, v in r.items() if k.endswith('tflops')})
"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$R.json')); print(round(d['value'],3), round(d['ms_per_step'],3), d['clocks'], d['finite'], d['roofline']['kernel_families_us'], d['e2e']['value'], d['dropin']['value'], {k:(round(v['value'],2), v.get('parity',{}).get('rel_l2')) for k,v in d['configs'].items()}, d['parity'])"
tail -n 3 gpurun_out/bench_$R.err
