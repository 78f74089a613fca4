#!/bin/bash
# experiment build (-DTFX_ATTN8): parity of schedule 8 on every attention test shape, then the kernel micro-benchmark next to the others
mkdir -p gpurun_out
export TFX_EXTRA_QT=${TFX_EXTRA_QT:-28,8,38}
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention and extra" > gpurun_out/pytest_attn8.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_attn8.log
timeout 300 python tools/bench_kernels.py --only attention --json gpurun_out/kernels_attn8.json 2>&1 | tail -n 6
