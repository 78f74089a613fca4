"""One denoising step of the 12B engine between cudaProfilerStart/Stop, for `ncu --profile-from-start off`:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/one_step.py --workload cfg3
Same engine, inputs and call (tfx_step_scheduled, step 3 of a 30-step schedule) as bench.py's timed loop."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.HEADLINE)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--attn-variant", type=int, default=-1)
    ap.add_argument("--m-band", type=int, default=-1)
    ap.add_argument("--k-snake", type=int, default=-1)
    args = ap.parse_args()
    from textflux_b200 import B200FluxTransformer, synthetic_getter
    from textflux_b200.engine import FrozenConfig
    dev = torch.device("cuda", 0)
    cfg = FrozenConfig(bench.CFG12B)
    eng = B200FluxTransformer(cfg, synthetic_getter(cfg, 1234, dev), device=dev)
    if args.attn_variant >= 0:
        eng.set_option("attn_variant", args.attn_variant)
    if args.k_snake >= 0:
        eng.set_option("gemm_k_snake", args.k_snake)
    if args.m_band >= 0 or args.m_band <= -100:
        eng.set_option("gemm_m_band", args.m_band)
    inp = bench.Inputs(args.workload, dev, 0, 1)
    eng.set_schedule(inp.ts, inp.guidance, inp.pooled_b, inp.S, inp.T)
    lat = inp.latents0
    for k in range(args.warmup):
        lat = eng.step_scheduled(k, lat, inp.cond, inp.prompt_b, inp.img_ids, inp.txt_ids, inp.sig_cpu[k], inp.sig_cpu[k + 1])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    k = args.warmup
    lat = eng.step_scheduled(k, lat, inp.cond, inp.prompt_b, inp.img_ids, inp.txt_ids, inp.sig_cpu[k], inp.sig_cpu[k + 1])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("step done, finite:", bool(torch.isfinite(lat.float()).all()), "launches:", eng.counter("launches"))


if __name__ == "__main__":
    main()
