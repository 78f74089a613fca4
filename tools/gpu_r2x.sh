#!/bin/bash
# round 2, visit X: ncu captures of the convolution-mode GEMM launches of one VAE decode (tensor pipe activity per shape); the
# report stays on the box, only the condensed table comes back
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section InstructionStats \
   --clock-control none -k regex:gemm_tcgen05 -o /tmp/prof_vae_conv -f python tools/one_vae.py --what decode > gpurun_out/ncu_vae_conv_r2x.log 2>&1; echo "ncu exit $?"
python tools/ncu_brief.py /tmp/prof_vae_conv.ncu-rep > gpurun_out/vae_conv_brief_r2x.md; wc -l gpurun_out/vae_conv_brief_r2x.md
