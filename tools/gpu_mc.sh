#!/bin/bash
mkdir -p gpurun_out
R=${1:-mc}
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "linear and (mc2 or mc4)" > gpurun_out/pytest_mc.log 2>&1; echo "pytest mc exit $?"; tail -n 15 gpurun_out/pytest_mc.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_model.log 2>&1; echo "pytest model exit $?"; tail -n 12 gpurun_out/pytest_model.log
timeout 600 python tools/bench_kernels.py --json gpurun_out/kernels_$R.json > gpurun_out/kernels_$R.log 2>&1; echo "kernels exit $?"
python - <<PY
import json
for r in json.load(open("gpurun_out/kernels_$R.json")):
    if r["kernel"]=="gemm": print(r["name"], {k: round(v) for k,v in r.items() if k.endswith("tflops")})
    else: print(r)
PY
for mc in 0 2 4; do
timeout 600 python bench.py --steps 10 --warmup 3 --mcast $mc --no-cpu-baseline > gpurun_out/bench_${R}_mc$mc.json 2> gpurun_out/bench_${R}_mc$mc.err; echo "bench mc$mc exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${R}_mc$mc.json')); print('mcast $mc', d['value'], d['ms_per_step'], d['roofline']['kernel_families_us'], d['clocks'])"
tail -n 3 gpurun_out/bench_${R}_mc$mc.err
done
