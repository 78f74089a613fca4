mkdir -p gpurun_out
for flag in "--mcast 0" "--mcast 2" "--mcast 4" "--mcast 0 --cta-group 1" "--mcast 0"; do
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline $flag > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('flag[$flag]', round(d['value'],3), round(d['ms_per_step'],3), d['clocks'], d['roofline']['kernel_families_us']['gemm']['us'], d['roofline']['kernel_families_us']['attn']['us'])"
tail -n 2 gpurun_out/ab.err
done
