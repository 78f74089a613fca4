#!/bin/bash
# one GPU visit: full gpu test suite, smoke, kernel micro-benchmarks, headline bench, ncu launch list + full captures
mkdir -p gpurun_out
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=120 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python tools/bench_kernels.py --json gpurun_out/kernels_$R.json > gpurun_out/kernels_$R.log 2>&1; echo "kernels exit $?"
cat gpurun_out/kernels_$R.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
cat gpurun_out/bench_$R.json; tail -n 5 gpurun_out/bench_$R.err
if [ "$2" != "noncu" ]; then
KREG='regex:tcgen05|ln_modulate|gemv_kernel|rope_table|timestep_embed|set_float'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 900 -c 300 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 420 -c 3 -o gpurun_out/prof_gemm_$R -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 60 -c 2 -o gpurun_out/prof_attn_$R -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out/
fi
