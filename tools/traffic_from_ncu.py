"""DRAM traffic of one denoising step from an ncu `--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
launch list (tools/gpu_prof.sh): per kernel family and for the whole step -> JSON that bench.py reports as roofline.traffic.
Usage: python tools/traffic_from_ncu.py gpurun_out/dram_r1m.csv profiles/r1m_traffic.json"""
import csv
import json
import sys
from collections import defaultdict


def main(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(dict)  # launch id -> metric -> value (bytes / ns)
    name = {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1, "msecond": 1e6}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        per[int(r[0])][r[ix["Metric Name"]]] = v * scale.get(r[ix["Metric Unit"]], 1)
        name[int(r[0])] = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    fam = defaultdict(lambda: dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, time_us=0.0))
    for i, m in per.items():
        n = name[i]
        f = "gemm" if n.startswith("gemm") else "attention" if n.startswith("attention") else "ln_modulate" if n.startswith("ln_") else "other"
        fam[f]["launches"] += 1
        fam[f]["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
        fam[f]["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
        fam[f]["time_us"] += m.get("gpu__time_duration.sum", 0.0) / 1e3
    tot_r = sum(f["dram_read_bytes"] for f in fam.values())
    tot_w = sum(f["dram_write_bytes"] for f in fam.values())
    res = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {len(per)} kernels of one cfg2 step ({path})",
           "launches": len(per), "dram_read_bytes_per_step": tot_r, "dram_write_bytes_per_step": tot_w,
           "dram_bytes_per_step": tot_r + tot_w, "families": fam}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "families"}))
    for k, v in fam.items():
        print(k, {a: (round(b / 1e9, 3) if "bytes" in a else round(b, 1)) for a, b in v.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
