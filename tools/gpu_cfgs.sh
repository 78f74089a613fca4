mkdir -p gpurun_out
for w in cfg3 cfg4 cfg5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']; print('$w', round(d['value'],3), 'steps/s', round(d['ms_per_step'],2), 'ms', 'TF/s', round(r['achieved'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['kernel_families_us'], d['finite'])"
tail -n 2 gpurun_out/bench_$w.err
done
