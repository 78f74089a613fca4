#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu_r2aa.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_r2aa.log | tail -n 2; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2aa.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2aa.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke_r2aa.log | cut -c1-250
S0=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err; echo "bench exit $? in $(( $(date +%s) - S0 )) s"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2aa.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['dropin']['value'], {k:v['value'] for k,v in d['configs'].items()}, d['once_per_image'])"
