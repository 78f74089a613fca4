mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for flag in "--pdl 0" "--pdl 1" "--pdl 0" "--pdl 1" "--pdl 1 --attn-variant 3"; do
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline $flag > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('flag[$flag]', round(d['value'],3), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['finite'], d['roofline']['kernel_families_us']['gemm']['us'], d['roofline']['kernel_families_us']['attn']['us'])"
tail -n 2 gpurun_out/ab.err
done
