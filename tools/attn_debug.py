"""Debug aid: error map of tfx_op_attention (code from argv) per 128-row block and 64-column block of one head."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textflux_b200 import _lib

code = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for (T, S) in [(0, 128), (0, 256), (128, 256), (0, 512), (0, 1024)]:
    H, dh, N = 1, 128, T + S
    g = torch.Generator(device="cuda").manual_seed(N)
    q = torch.randn(1, H, N, dh, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(1, H, N, dh, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, H, N, dh, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.zeros(N, H * dh, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, code, st))
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())[0, 0]
    # variants of the reference that drop key blocks, to recognise the error pattern
    print(f"N={N}")
    for rb in range(N // 128):
        row = []
        for cb in range(2):
            o = out[rb * 128:(rb + 1) * 128, cb * 64:(cb + 1) * 64].float()
            r = ref[rb * 128:(rb + 1) * 128, cb * 64:(cb + 1) * 64]
            row.append(f"{((o - r).norm() / r.norm()).item():.3f} (|o|/|r| {(o.norm() / r.norm()).item():.2f})")
        print(f"  rows {rb * 128:5d}.. : cols 0-63 {row[0]}   cols 64-127 {row[1]}")
