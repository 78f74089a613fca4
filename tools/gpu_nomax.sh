#!/bin/bash
# -DTFX_ATTN_NOMAX=1 build (tools/experiments/_build/libtextflux_b200_nomax.so) against the default build: parity, kernel micro-benchmark, in-step A/B
mkdir -p gpurun_out
NM=$PWD/tools/experiments/_build/libtextflux_b200_nomax.so
TEXTFLUX_B200_LIB=$NM timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention" > gpurun_out/pytest_nomax.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_nomax.log
for L in default nomax; do
  if [ $L = nomax ]; then export TEXTFLUX_B200_LIB=$NM; else unset TEXTFLUX_B200_LIB; fi
  timeout 300 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn_$L.json 2>&1 | tail -n 5 | python -c "
import sys,ast
for l in sys.stdin:
    if l.startswith('{'):
        d=ast.literal_eval(l); print('$L', d['N'], {k[:-7]:round(v) for k,v in d.items() if k.endswith('tflops')})"
done
for rep in 1 2; do
for L in default nomax; do
  if [ $L = nomax ]; then export TEXTFLUX_B200_LIB=$NM; else unset TEXTFLUX_B200_LIB; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-configs --no-image-stages > gpurun_out/ab_nomax_${L}_$rep.json 2> gpurun_out/ab_nomax_${L}_$rep.err || tail -n 3 gpurun_out/ab_nomax_${L}_$rep.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_nomax_${L}_$rep.json')); f=d['roofline']['kernel_families_us']
print('[$L] rep $rep:', round(d['ms_per_step'],3), 'ms/step | gemm', f['gemm']['us'], 'attn', f['attn']['us'], '| parity', d.get('parity'))" 2>&1 | cut -c1-400
done; done
