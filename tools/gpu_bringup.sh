#!/bin/bash
# first-contact bring-up: each group in its own process so a trapped kernel cannot poison the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest "$@" -m gpu -q -rA -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run t_probe tests/test_gpu_ops.py -k "umma"
run t_pointwise tests/test_gpu_ops.py -k "not umma and not attention and not linear"
run t_linear1 tests/test_gpu_ops.py -k "linear and not cg2"
run t_linear2 tests/test_gpu_ops.py -k "linear and cg2"
run t_attn1 tests/test_gpu_ops.py -k "attention and qt1"
run t_attn2 tests/test_gpu_ops.py -k "attention and qt2"
run t_model tests/test_gpu_model.py -s
cat gpurun_out/summary.txt
for f in gpurun_out/t_*.log; do echo "=== $f"; tail -n 25 $f; done
