#!/bin/bash
# schedule-3 attention bring-up: parity of every variant, clock trace, micro-benchmark, then the step A/B
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attention" --timeout=60 > gpurun_out/attn_all.log 2>&1
echo "pytest attention exit $?"; tail -n 4 gpurun_out/attn_all.log
timeout 120 python tools/attn_trace.py --json gpurun_out/attn_trace.json 2>&1 | tail -n 16
timeout 300 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn3.json 2>&1 | grep attention
for flag in "--attn-variant 5" "--attn-variant 6"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $flag > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('flag[$flag]', round(d['value'],3), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['finite'], d['roofline']['kernel_families_us']['gemm']['us'], d['roofline']['kernel_families_us']['attn']['us'])"
tail -n 2 gpurun_out/ab.err
done
