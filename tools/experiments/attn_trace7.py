"""Clock trace of one CTA of the schedule-7 attention kernel (attention7.cuh, trace build) through the C ABI.
Per 64-key step s and query tile q: 0 s_full seen by the softmax warp, 1 scores in registers, 2 maximum known, 3 P handed over,
5 issuer saw P, 6 PV issued, 7 next QK issued.  Usage: python tools/attn_trace7.py [--S 4608]"""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textflux_b200 import _lib  # noqa: E402

NAMES = ["s_seen", "ld_done", "max_done", "p_handed", "-", "issuer_saw_p", "pv_issued", "qk_issued"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=4608)
    args = ap.parse_args()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    T, S, H, dh = 512, args.S, 24, 128
    N = T + S
    n_steps = (N + 63) // 64
    q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
    tr = torch.zeros(n_steps * 2 * 8, dtype=torch.int64, device="cuda")
    _lib.check(lib.tfx_debug_set_attention_trace(tr.data_ptr()))
    for _ in range(3):
        _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, 128, st))
    torch.cuda.synchronize()
    _lib.check(lib.tfx_debug_set_attention_trace(None))
    t = tr.view(n_steps, 2, 8).cpu()
    lo, hi = 6, n_steps - 4
    print(f"schedule 7, N={N}, {n_steps} steps of 64 keys; cycles relative to s_seen of the same (step, tile), medians over steps {lo}..{hi}")
    for qq in range(2):
        row = []
        for e in (1, 2, 3, 5, 6, 7):
            d = [int(t[s, qq, e] - t[s, qq, 0]) for s in range(lo, hi) if t[s, qq, e] > 0]
            row.append(f"{NAMES[e]} {statistics.median(d) if d else None}")
        per = [int(t[s + 1, qq, 0] - t[s, qq, 0]) for s in range(lo, hi)]
        gap = [int(t[s + 1, qq, 0] - t[s, qq, 3]) for s in range(lo, hi)]
        print(f"  q{qq}: " + "  ".join(row) + f"  | step period {statistics.median(per)}  (P handed -> next s_seen {statistics.median(gap)})")
    ph = [int(t[s, 1, 0] - t[s, 0, 0]) for s in range(lo, hi)]
    print(f"  q1 lags q0 by {statistics.median(ph)} cycles; total {int(t[n_steps - 1, 1, 3] - t[0, 0, 0])} cycles for {n_steps} steps")
    for s in range(10, 14):
        print("  step", s, [[int(t[s, qq, e] - t[10, 0, 0]) for e in (0, 1, 2, 3, 5, 6, 7)] for qq in range(2)])


if __name__ == "__main__":
    main()
