#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attention and pair" --timeout=60 > gpurun_out/attn_pair.log 2>&1
echo "pytest pair exit $?"; tail -n 15 gpurun_out/attn_pair.log
timeout 300 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_pair.json 2>&1 | grep attention
