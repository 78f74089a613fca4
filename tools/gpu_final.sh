#!/bin/bash
# final visit of a session: whole gpu suite, smoke, headline bench (with the CPU baseline), reference arm, other configs
mkdir -p gpurun_out
R=${1:-r1m}
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=120 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
cat gpurun_out/bench_$R.json; tail -n 3 gpurun_out/bench_$R.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref exit $?"
cat gpurun_out/bench_ref_$R.json
bash tools/gpu_cfgs.sh
for w in cfg3 cfg4 cfg5; do cp gpurun_out/bench_$w.json gpurun_out/bench_${R}_$w.json; done
