#!/bin/bash
mkdir -p gpurun_out
for CODE in 24 26; do
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:attention -o gpurun_out/prof_attn9_$CODE -f python tools/experiments/attn_probe.py $CODE 5120 > gpurun_out/ncu_attn9_$CODE.log 2>&1; echo "ncu $CODE exit $?"
done
ls -la gpurun_out/prof_attn9_*.ncu-rep
