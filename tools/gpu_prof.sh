#!/bin/bash
# ncu evidence for one denoising step of `python bench.py` (cfg2): launch list (time), DRAM traffic list, full captures of
# the dominant GEMMs and of the attention kernel.  Per-launch values are cold-cache and serialised (ncu replays kernels
# one at a time at boost clocks), so shares and pipe percentages are comparable with the live numbers, absolutes are not.
mkdir -p gpurun_out
R=${1:-r1m}
KREG='regex:tcgen05|ln_modulate|gemv_kernel|rope_table|timestep_embed|set_float'
# warm-up 3 steps (3*292 + set-up launches) then one step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 951 -c 293 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$KREG" -s 951 -c 293 --csv --log-file gpurun_out/dram_$R.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dram.log 2>&1; echo "ncu dram exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 467 -c 6 -o gpurun_out/prof_gemm_$R -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention3_tcgen05 -s 60 -c 2 -o gpurun_out/prof_attn_$R -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out/ | grep $R
