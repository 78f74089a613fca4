#!/bin/bash
# same-box A/B of two library builds: tools/experiments/_build/libtextflux_b200_old.so against the in-tree one (attention micro-benchmark, then the step)
mkdir -p gpurun_out
OLD=$PWD/tools/experiments/_build/libtextflux_b200_old.so
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention" 2>&1 | tail -n 1
for L in old new old new; do
  if [ $L = old ]; then export TEXTFLUX_B200_LIB=$OLD; else unset TEXTFLUX_B200_LIB; fi
  timeout 300 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$L.json 2>&1 | tail -n 5 | python -c "
import sys,ast
for l in sys.stdin:
    if l.startswith('{'):
        d=ast.literal_eval(l); print('$L', d['N'], {k[:-7]:round(v) for k,v in d.items() if k.endswith('tflops')})"
done
if [ "$1" = step ]; then
for rep in 1 2; do
for L in old new; do
  if [ $L = old ]; then export TEXTFLUX_B200_LIB=$OLD; else unset TEXTFLUX_B200_LIB; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-configs --no-image-stages > gpurun_out/ab_on_${L}_$rep.json 2> gpurun_out/ab_on_${L}_$rep.err || tail -n 3 gpurun_out/ab_on_${L}_$rep.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_on_${L}_$rep.json')); f=d['roofline']['kernel_families_us']
print('[$L] rep $rep:', round(d['ms_per_step'],3), 'ms/step | gemm', f['gemm']['us'], 'attn', f['attn']['us'], 'ln', f['ln']['us'], '| clocks', d['clocks']['sm_mhz'])" 2>&1 | cut -c1-300
done; done
fi
