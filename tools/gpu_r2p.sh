#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r2p.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_r2p.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_textenc.py -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread > gpurun_out/pytest_textenc_r2p.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_textenc_r2p.log | head
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_textenc.py tests/test_gpu_vae.py -m gpu -q -p no:cacheprovider -k "short_and_ragged or golden or rejects" > gpurun_out/sanitizer_memcheck_stages_r2p.log 2>&1; echo "memcheck exit $?"; tail -n 3 gpurun_out/sanitizer_memcheck_stages_r2p.log
