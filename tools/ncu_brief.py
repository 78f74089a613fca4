"""Condensed table of an .ncu-rep (read here, no GPU): one row per captured launch with the counters the roofline argument uses.
Usage: python tools/ncu_brief.py a.ncu-rep [b.ncu-rep ...] > profiles/xyz.md"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % active", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % elapsed", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", 1),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %", 1),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU %", 1),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("dram__bytes_read.sum", "DRAM rd MB", 1e-6), ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1), ("launch__registers_per_thread", "regs", 1),
        ("sm__cycles_active.avg", "SM active cyc", 1), ("sm__cycles_elapsed.avg", "SM elapsed cyc", 1)]
SCALE = {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(paths):
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        u = dict(zip(hdr, units))
        print(f"### `{path.split('/')[-1]}`\n")
        print("| kernel | grid | " + " | ".join(c[1] for c in COLS) + " |")
        print("|---|---|" + "---|" * len(COLS))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            cells = []
            for key, _, mul in COLS:
                v = d.get(key, "")
                if v in ("", "n/a"):
                    cells.append("-")
                    continue
                x = float(v.replace(",", "")) * SCALE.get(u[key], 1) * mul
                cells.append(f"{x:.1f}" if x < 1e5 else f"{x:.0f}")
            print(f"| {name} | {d['Grid Size']} | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
