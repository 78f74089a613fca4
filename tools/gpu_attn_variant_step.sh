#!/bin/bash
# in-step A/B of the attention schedule (5 = schedule 3 split-P, 9 = schedule 5 persistent, default = per shape) at the three image sizes
mkdir -p gpurun_out
for W in cfg2 cfg3 cfg5; do
  bash tools/gpu_ab.sh av_$W "--workload $W --steps 12 --attn-variant 5" "--workload $W --steps 12 --attn-variant 9" "--workload $W --steps 12"
done
