"""Clock trace of one CTA of the schedule-3 attention kernel (attention3.cuh, kTrace build) through the C ABI.

Per KV tile j and query tile q the kernel stamps clock64 at: 0 s_full seen by the softmax warp, 1 score row in
registers, 2 row max known, 3 first half of P handed over, 4 second half handed over, 5 issuer saw P, 6 PV issued,
7 next QK issued.  Prints the steady-state medians (cycles) and writes the raw stamps as JSON.
Usage: python tools/attn_trace.py [--json out.json] [--S 2048]"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textflux_b200 import _lib  # noqa: E402

NAMES = ["s_seen", "ld_done", "max_done", "p_half0", "p_half1", "issuer_saw_p", "pv_issued", "qk_issued"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--S", type=int, default=2048)
    args = ap.parse_args()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    T, S, H, dh = 512, args.S, 24, 128
    N = T + S
    n_kv = (N + 127) // 128
    q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
    res = {}
    for code, name in ((125, "whole-P"), (126, "split-P")):  # 100 + tfx_op_attention q_tiles code
        tr = torch.zeros(n_kv * 2 * 8, dtype=torch.int64, device="cuda")
        _lib.check(lib.tfx_debug_set_attention_trace(tr.data_ptr()))
        for _ in range(3):
            _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, code, st))
        torch.cuda.synchronize()
        _lib.check(lib.tfx_debug_set_attention_trace(None))
        t = tr.view(n_kv, 2, 8).cpu()
        res[name] = t.tolist()
        print(f"== {name} (N={N}, {n_kv} KV tiles); cycles relative to s_seen of the same (j, q), median over j = 3..{n_kv - 2}")
        for qq in range(2):
            row = []
            for e in range(1, 8):
                d = [int(t[j, qq, e] - t[j, qq, 0]) for j in range(3, n_kv - 1) if t[j, qq, e] > 0]
                row.append(f"{NAMES[e]} {statistics.median(d) if d else None}")
            per = [int(t[j + 1, qq, 0] - t[j, qq, 0]) for j in range(3, n_kv - 2)]
            print(f"  q{qq}: " + "  ".join(row) + f"  | period {statistics.median(per)}")
        ph = [int(t[j, 1, 0] - t[j, 0, 0]) for j in range(3, n_kv - 1)]
        print(f"  q1 lags q0 by {statistics.median(ph)} cycles; total {int(max(t[n_kv - 1, 1, 4], t[n_kv - 1, 1, 3]) - t[0, 0, 0])} cycles for {n_kv} tiles")
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f)


if __name__ == "__main__":
    main()
