#!/bin/bash
# same-box A/B of engine options inside the real step: tools/gpu_ab.sh <tag> "<bench flags A>" "<bench flags B>" ...
mkdir -p gpurun_out
TAG=$1; shift
i=0
for FLAGS in "$@"; do
  for rep in 1 2; do
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-configs --no-image-stages $FLAGS > gpurun_out/ab_${TAG}_${i}_$rep.json 2> gpurun_out/ab_${TAG}_${i}_$rep.err || tail -n 3 gpurun_out/ab_${TAG}_${i}_$rep.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_${TAG}_${i}_$rep.json')); f=d['roofline']['kernel_families_us']
print('[$FLAGS] rep $rep:', round(d['ms_per_step'],3), 'ms/step', round(d['value'],3), 'steps/s | clocks', d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), '| gemm', f['gemm']['us'], 'attn', f['attn']['us'], 'ln', f['ln']['us'], '| dropin', round(d['dropin']['ms_per_step'],3))"
  done
  i=$((i+1))
done
