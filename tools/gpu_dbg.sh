mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -s -p no:cacheprovider --timeout=50 --timeout-method=thread > gpurun_out/dbg_model.log 2>&1
echo "exit $?"; tail -n 60 gpurun_out/dbg_model.log
