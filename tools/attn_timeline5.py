"""Per-item timeline of the persistent attention kernel (attention5.cuh trace build).
Usage: python tools/attn_timeline5.py [--S 2048] [--json out.json]"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textflux_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=2048)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    T, S, H, dh = 512, args.S, 24, 128
    N = T + S
    q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
    G = torch.cuda.get_device_properties(0).multi_processor_count
    tr = torch.zeros(G * 64, dtype=torch.int64, device="cuda")
    for i in range(3):
        if i == 2:
            _lib.check(lib.tfx_debug_set_attention_cta_trace(tr.data_ptr()))
        _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, 127, st))
    torch.cuda.synchronize()
    _lib.check(lib.tfx_debug_set_attention_cta_trace(None))
    t = tr.view(G, 8, 8).cpu()
    t0 = int(t[:, 0, 6][t[:, 0, 6] > 0].min())
    print(f"N={N}: persistent kernel, {G} CTAs; per item (median over CTAs), us relative to the first Q request of the kernel")
    names = ["first S seen", "last P handed", "last PV retired", "O out of TMEM", "item done", "first PV issued", "Q requested"]
    span = 0
    for item in range(8):
        rows = [r for r in t[:, item] if int(r[0]) > 0]
        if not rows:
            break
        its = statistics.median(int(r[7]) for r in rows)
        med = {n: statistics.median((int(r[i]) - t0) / 1e3 for r in rows) for i, n in enumerate(names)}
        loop = statistics.median((int(r[1]) - int(r[0])) / 1e3 for r in rows)
        tail = statistics.median((int(r[4]) - int(r[1])) / 1e3 for r in rows)
        span = max(span, max((int(r[4]) - t0) / 1e3 for r in rows))
        print(f"  item {item} ({len(rows)} CTAs, {its} iterations): Q req {med['Q requested']:.1f}  first S {med['first S seen']:.1f}  first PV {med['first PV issued']:.1f}  "
              f"last P {med['last P handed']:.1f}  PV retired +{med['last PV retired'] - med['last P handed']:.2f}  O out +{med['O out of TMEM'] - med['last PV retired']:.2f}  "
              f"done +{med['item done'] - med['O out of TMEM']:.2f} | loop {loop:.1f} us = {1e3 * loop / max(its, 1):.0f} ns/iter, tail {tail:.2f} us")
    # gap between an item's end and the next item's first S for the same warpgroup
    gaps = []
    for c in range(G):
        for item in range(7):
            if int(t[c, item + 1, 0]) > 0 and int(t[c, item, 4]) > 0:
                gaps.append((int(t[c, item + 1, 0]) - int(t[c, item, 1])) / 1e3)
    if gaps:
        print(f"  last P of an item -> first S of the next (q0 warpgroup): median {statistics.median(gaps):.2f} us, max {max(gaps):.2f} us")
    print(f"  span {span:.1f} us")
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"N": N, "trace": t.tolist()}, f)


if __name__ == "__main__":
    main()
