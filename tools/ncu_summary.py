"""Summarise an .ncu-rep (read here, no GPU): per-kernel duration, DRAM bytes, tensor-pipe %, etc. -> markdown."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_uniform", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum ", "lts__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg", "smsp__inst_executed.sum ",
        "sm__pipe_tensor", "tensor", "smsp__average_warp", "launch__occupancy_limit", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_xu_cycles_active", "sm__inst_executed_pipe_xu", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct",
        "smsp__pcsamp_warps_issue_stalled", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg.pct"]


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of `{path}`", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"## {d['Kernel Name']}  grid {d['Grid Size']} block {d['Block Size']}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for h, u in zip(hdr, units):
            if any(k.strip() in h for k in KEYS) and "TriageCompute" not in h and d[h] not in ("", "n/a"):
                lines.append(f"| {h} | {d[h]} | {u} |")
        lines.append("")
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
