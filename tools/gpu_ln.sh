#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider --timeout=120 > gpurun_out/ln_tests.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/ln_tests.log
timeout 300 python tools/bench_kernels.py --only ln 2>&1 | grep -E "ln_modulate|gemv"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ln.json 2> gpurun_out/bench_ln.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_ln.json')); print(round(d['value'],3), round(d['ms_per_step'],3), d['clocks'], d['finite'], d['roofline']['kernel_families_us'], d['e2e'])"
