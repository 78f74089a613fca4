"""torch SDPA (the library kernel the reference dispatches, attention_processor.py:2039) on the joint-attention shape, for ncu."""
import sys
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
q = torch.randn(1, 24, N, 128, device="cuda").to(torch.bfloat16)
k = torch.randn(1, 24, N, 128, device="cuda").to(torch.bfloat16)
v = torch.randn(1, 24, N, 128, device="cuda").to(torch.bfloat16)
for _ in range(3):
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
torch.cuda.synchronize()
torch.cuda.profiler.start()
o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(o.float().abs().mean().item())
