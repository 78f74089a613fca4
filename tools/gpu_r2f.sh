#!/bin/bash
# round 2, visit F: whole gpu suite, GEMM micro-benchmark, compute-sanitizer on the tiny config, ncu evidence for one cfg3 step
mkdir -p gpurun_out
R=${1:-r2f}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$R.log | tail -n 3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_$R.log | head -20
timeout 600 python tools/bench_kernels.py --only gemm --json gpurun_out/kernels_gemm_$R.json > gpurun_out/kernels_gemm_$R.log 2>&1; echo "kernels exit $?"
grep "'gemm'" gpurun_out/kernels_gemm_$R.log | python -c "
import sys, ast
for line in sys.stdin:
    if line.startswith('{'):
        r = ast.literal_eval(line); print(r['name'], r['M'], r['N'], r['K'], {k: round(v) for k, v in r.items() if k.endswith('tflops')})
"
# ---- compute-sanitizer: memcheck and racecheck on the tiny config (smoke) and on small op tests
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_${TOOL}_smoke_$R.log 2>&1; echo "$TOOL smoke exit $?"
  tail -n 4 gpurun_out/sanitizer_${TOOL}_smoke_$R.log
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -x \
     -k "(test_linear_store and (128-256-64 or 80-256-128 or 300-320-192)) or (test_attention and (1-2-128-128-128 or 2-4-16-64-64 or 1-3-40-217-128)) or test_linear_gelu_and_gate_res or test_linear_qkv_epilogue or test_linear_euler_epilogue or test_ln_modulate or test_gemv" \
     > gpurun_out/sanitizer_${TOOL}_ops_$R.log 2>&1; echo "$TOOL ops exit $?"
  tail -n 4 gpurun_out/sanitizer_${TOOL}_ops_$R.log
done
# ---- ncu: launch list + DRAM traffic of exactly one cfg3 step, full captures of the dominant kernels
KREG='regex:tcgen05|ln_modulate|gemv_kernel|rope_table|timestep_embed|set_float|mod_cache'
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/launches_${R}_cfg3.csv \
   python tools/one_step.py --workload cfg3 > gpurun_out/ncu_launches_$R.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 6 -o gpurun_out/prof_gemm_${R}_cfg3 -f \
   python tools/one_step.py --workload cfg3 > gpurun_out/ncu_gemm_$R.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention -s 3 -c 2 -o gpurun_out/prof_attn_${R}_cfg3 -f \
   python tools/one_step.py --workload cfg3 > gpurun_out/ncu_attn_$R.log 2>&1; echo "ncu attn cfg3 exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention -s 3 -c 2 -o gpurun_out/prof_attn_${R}_cfg5 -f \
   python tools/one_step.py --workload cfg5 > gpurun_out/ncu_attn5_$R.log 2>&1; echo "ncu attn cfg5 exit $?"
ls -la gpurun_out/ | grep $R | head -40
