#!/bin/bash
# round 2, visit L: whole GPU suite on the current tree, text-encoder benchmark, headline bench (did the conv / side-path branches cost the GEMM anything?)
mkdir -p gpurun_out
R=${1:-r2l}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$R.log | tail -n 3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_$R.log | head -20
timeout 900 python tools/bench_textenc.py --json gpurun_out/bench_textenc_$R.json > gpurun_out/bench_textenc_$R.log 2>&1; echo "bench textenc exit $?"
tail -n 3 gpurun_out/bench_textenc_$R.log | cut -c1-500
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$R.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['dropin']['value'], {k:(v['value'], v.get('parity',{}).get('rel_l2')) for k,v in d['configs'].items()}, d.get('gpu_eager_baseline'), d.get('parity'))"
