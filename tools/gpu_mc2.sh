export TFX_DEBUG=1
python tools/gpu_mc2.py 2
python tools/gpu_mc2.py 21
python tools/gpu_mc2.py 22
python tools/gpu_mc2.py 24
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "linear and (mc2 or mc4)" 2>&1 | tail -3
