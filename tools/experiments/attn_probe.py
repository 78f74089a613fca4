"""One launch of an attention schedule code (tfx_op_attention's q_tiles) on the joint-attention shape, for ncu:
python tools/experiments/attn_probe.py <code> [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from textflux_b200 import _lib  # noqa: E402

code = int(sys.argv[1])
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5120
lib = _lib.load()
H, dh, T = 24, 128, 512
q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream


def run():
    _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, N - T, dh, code, st))


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(out.float().abs().mean().item())
