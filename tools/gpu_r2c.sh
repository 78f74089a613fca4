#!/bin/bash
mkdir -p gpurun_out
for S in 2048 4608 8192; do
  python tools/attn_timeline.py --S $S --json gpurun_out/timeline_warm_$S.json
  python tools/attn_timeline.py --S $S --cold --json gpurun_out/timeline_cold_$S.json
done
python tools/attn_trace.py --S 4608 | tail -8
timeout 600 python bench.py --steps 60 --warmup 30 --no-cpu-baseline --no-eager --no-configs > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err
python -c "
import json; d=json.load(open('gpurun_out/bench_long.json')); print('long run:', round(d['ms_per_step'],3), d['clocks'], 'dropin', round(d['dropin']['ms_per_step'],3), round(d['dropin']['first_image_ms_per_step'],3))"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python -c "
import json; d=json.load(open('gpurun_out/bench_short.json')); print('short run:', round(d['ms_per_step'],3), d['clocks'], 'dropin', round(d['dropin']['ms_per_step'],3), round(d['dropin']['first_image_ms_per_step'],3))"
