mkdir -p gpurun_out
cat > /tmp/attn_one.py <<PY
import os, sys, torch
sys.path.insert(0, os.getcwd())
from textflux_b200 import _lib
lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
T, S, H, dh = 512, 2048, 24, 128
N = T + S
q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
for qt in (2, 2, 2, 32):
    _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, qt, st))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 2 -c 2 -o gpurun_out/prof_attn_r1d -f python /tmp/attn_one.py > gpurun_out/ncu_attn_r1d.log 2>&1; echo "ncu exit $?"
tail -n 3 gpurun_out/ncu_attn_r1d.log
