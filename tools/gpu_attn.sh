mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attention" --timeout=60 > gpurun_out/attn_all.log 2>&1
tail -n 3 gpurun_out/attn_all.log
timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep attention
