"""Wait accounting of one CTA of the schedule-9 attention experiment (split-issue, kTrace build): cycles each role spends in its barrier waits."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from textflux_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
lib = _lib.load()
H, dh, T = 24, 128, 512
q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.check(lib.tfx_debug_set_attention_trace(tr.data_ptr()))
for _ in range(3):
    _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, N - T, dh, 114, st))
torch.cuda.synchronize()
_lib.check(lib.tfx_debug_set_attention_trace(None))
t = tr.cpu().tolist()
n = max(t[13], 1)
names = {0: "PV issuer: whole loop", 1: "PV issuer: wait v_full", 2: "PV issuer: wait P half 0", 3: "PV issuer: wait P half 1", 5: "QK issuer: wait pv_done",
         6: "QK issuer: wait k_full", 12: "QK issuer: whole loop", 7: "softmax wg0: wait s_full", 8: "softmax wg0: wait partner's reference", 9: "softmax wg0: whole loop",
         14: "softmax wg1: wait s_full", 15: "softmax wg1: wait partner's reference", 16: "softmax wg1: whole loop", 10: "K producer: wait k_empty", 11: "V producer: wait v_empty"}
print(f"N = {N}, {n} KV tiles; cycles per KV tile")
for i in sorted(names):
    print(f"  {names[i]:44s} {t[i] / n:8.0f}")
