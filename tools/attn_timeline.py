"""Per-CTA timeline of the schedule-3 attention kernel (attention3.cuh trace build): where a launch's time goes besides the
steady-state KV loop -- CTA set-up, first loads, pipeline fill, drain, output store, and the gaps between the CTAs an SM runs.
Usage: python tools/attn_timeline.py [--S 2048] [--cold] [--json out.json]"""
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
    ap.add_argument("--cold", action="store_true", help="flush L2 before the traced launch")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    T, S, H, dh = 512, args.S, 24, 128
    N = T + S
    n_kv, n_ctas = (N + 127) // 128, ((N + 255) // 256) * H
    q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
    out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    tr = torch.zeros(n_ctas * 8, dtype=torch.int64, device="cuda")
    for i in range(3):
        if i == 2:
            _lib.check(lib.tfx_debug_set_attention_cta_trace(tr.data_ptr()))
            if args.cold:
                flush.zero_()
        _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, 126, st))
    torch.cuda.synchronize()
    _lib.check(lib.tfx_debug_set_attention_cta_trace(None))
    t = tr.view(n_ctas, 8).cpu()
    t0 = int(t[:, 0].min())
    rows = [[int(x) - t0 for x in r[:7]] + [int(r[7])] for r in t]
    span = max(r[6] for r in rows)
    names = ["set-up", "Q+K0 load", "first QK", "KV loop", "last PV", "store"]
    print(f"N={N} ({n_kv} KV tiles), {n_ctas} CTAs, kernel span {span / 1e3:.1f} us, {'L2-cold' if args.cold else 'L2-warm'}")
    for i, nm in enumerate(names):
        d = [r[i + 1] - r[i] for r in rows]
        print(f"  {nm:10s} median {statistics.median(d) / 1e3:7.2f} us   p90 {sorted(d)[int(0.9 * len(d))] / 1e3:7.2f}   max {max(d) / 1e3:7.2f}")
    tot = [r[6] - r[0] for r in rows]
    loop = [r[4] - r[3] for r in rows]
    print(f"  CTA total  median {statistics.median(tot) / 1e3:7.2f} us; KV loop share {statistics.median(loop) / statistics.median(tot):.3f}; "
          f"per KV iteration {statistics.median(loop) / n_kv:.0f} ns")
    by_sm = {}
    for r in rows:
        by_sm.setdefault(r[7], []).append(r)
    gaps, per_sm_busy = [], []
    for sm, rs in by_sm.items():
        rs.sort(key=lambda r: r[0])
        gaps += [b[0] - a[6] for a, b in zip(rs, rs[1:])]
        per_sm_busy.append(sum(r[6] - r[0] for r in rs))
    if gaps:
        print(f"  gap between consecutive CTAs on an SM: median {statistics.median(gaps) / 1e3:.2f} us, max {max(gaps) / 1e3:.2f} us")
    cnt = sorted(len(v) for v in by_sm.values())
    print(f"  SMs used {len(by_sm)}, CTAs per SM min {cnt[0]} max {cnt[-1]}; mean SM busy {statistics.mean(per_sm_busy) / 1e3:.1f} us of {span / 1e3:.1f} us span "
          f"({statistics.mean(per_sm_busy) / span:.3f}); first CTA starts at {min(r[0] for r in rows) / 1e3:.2f} us, last ends {span / 1e3:.1f} us")
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"N": N, "rows": rows}, f)


if __name__ == "__main__":
    main()
