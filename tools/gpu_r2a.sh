#!/bin/bash
# round 2, visit A: whole gpu suite, headline bench (all legs), reference arm as the driver runs it, kernel micro-benchmarks
mkdir -p gpurun_out
R=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$R.txt 2>&1
nproc > gpurun_out/nproc_$R.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc_$R.txt; free -g | head -2 >> gpurun_out/nproc_$R.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread -x -s > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$R.log | tail -n 3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_$R.json; tail -n 5 gpurun_out/bench_$R.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref exit $?"
cat gpurun_out/bench_ref_$R.json | cut -c1-1500
timeout 600 python tools/bench_kernels.py --json gpurun_out/kernels_$R.json > gpurun_out/kernels_$R.log 2>&1; echo "kernels exit $?"
tail -n 8 gpurun_out/kernels_$R.log | cut -c1-900
