#!/bin/bash
# round 2, visit W: compute-sanitizer racecheck on the round's new kernels (convolution mode, GroupNorm, short-sequence attention,
# LoRA extension k-block, banded tile order), small cases
mkdir -p gpurun_out
R=${1:-r2w}
timeout 1500 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest -m gpu -q -p no:cacheprovider -x \
   tests/test_gpu_vae.py tests/test_gpu_textenc.py tests/test_gpu_lora.py \
   -k "vae_small or rejects or (short_and_ragged and (5 or 33)) or transformers_golden or (side_adapter and (300 or 128)) or (banded and 1300 and (1 or 5))" \
   > gpurun_out/sanitizer_racecheck_new_$R.log 2>&1; echo "racecheck exit $?"
tail -n 6 gpurun_out/sanitizer_racecheck_new_$R.log | cut -c1-250
grep -c "Race reported" gpurun_out/sanitizer_racecheck_new_$R.log
grep "Race reported" -A 2 gpurun_out/sanitizer_racecheck_new_$R.log | grep -v "^--" | cut -c1-200 | sort | uniq -c | sort -rn | head -12
