"""Per-kernel and per-GEMM-role summary of an ncu `--metrics gpu__time_duration.sum` launch list (one denoising step)."""
import csv
import sys
from collections import defaultdict


def main(path, M=2560, D=3072, L=19, Ls=38):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    seq = []
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        seq.append((name, v))
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in seq:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    out.append(f"| total | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |")
    g = [v for n, v in seq if n.startswith("gemm")]
    if len(g) == 2 + 4 * L + 2 * Ls + 1:
        fl = {"qkv (N=3D)": 2 * M * 3 * D * D, "attn out (K=D)": 2 * M * D * D, "ff up (N=4D, GELU)": 2 * M * 4 * D * D,
              "ff down (K=4D)": 2 * M * 4 * D * D, "single qkv+mlp (N=7D)": 2 * M * 7 * D * D, "single out (K=5D)": 2 * M * 5 * D * D}
        keys = list(fl)
        acc = {k: [] for k in fl}
        i = 2
        for _ in range(L):
            for k in keys[:4]:
                acc[k].append(g[i]); i += 1
        for _ in range(Ls):
            for k in keys[4:]:
                acc[k].append(g[i]); i += 1
        out += ["", "| GEMM role | launches | mean us | TFLOP/s | total ms |", "|---|---|---|---|---|"]
        for k, v in acc.items():
            m = sum(v) / len(v)
            out.append(f"| {k} | {len(v)} | {m:.1f} | {fl[k] / m / 1e6:.0f} | {sum(v) / 1000:.2f} |")
    a = [v for n, v in seq if n.startswith("attention")]
    if a:
        N = M
        out += ["", f"attention: mean {sum(a) / len(a):.1f} us, {4.0 * N * N * D / (sum(a) / len(a)) / 1e6:.0f} TFLOP/s"]
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
