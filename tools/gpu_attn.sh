mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attention" > gpurun_out/attn_all.log 2>&1
tail -n 6 gpurun_out/attn_all.log
timeout 600 python tools/bench_kernels.py --only attention 2>&1 | grep attention
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err
python -c "
import json; d=json.load(open('gpurun_out/bench_attn.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_families_us'], d['clocks'])"
tail -n 3 gpurun_out/bench_attn.err
