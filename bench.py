#!/usr/bin/env python
"""Headline benchmark: denoising-steps/sec (device-timed) of the FLUX-Fill 12B hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: FluxTransformer2DModel.forward on
cat(latents, cond) + FlowMatchEulerDiscreteScheduler.step, i.e. the loop body of pipeline_flux_fill.py:2082-2098.
Headline workload = the configuration BASELINE.json's metric is quoted on, "1024^2 glyph+scene": configs[2], the
TextFlux-beta strip (1024x1024 scene + 128-px glyph strip -> 1152x1024 canvas, S = 4608 image tokens, T = 512 text
tokens, one sample per GPU, bf16, full 19 + 38 blocks).  The other 12B configurations of BASELINE.json (cfg2 S=2048, cfg4
S=4096, cfg5 S=8192) are measured in the same run and reported under "configs".  Multi-GPU: weight replicas, one sample
per rank, one NCCL broadcast of the prompt embeddings + sigma schedule before the loop, no per-step collective ("weak").
Prints ONE JSON line on rank 0.

Legs of the default (ours) run, in order:
  value / roofline   engine's fused step (tfx_step_scheduled: one CUDA-graph launch per step), inputs resident in HBM
  e2e                the same call with pinned HOST buffers: H2D of the step's inputs + D2H of its result inside the timing
  dropin             the call sequence of the UNMODIFIED pipeline: torch.cat -> transformer.forward -> scheduler.step
  configs            cfg2 / cfg4 / cfg5, 10 fused steps each (+ parity at N = 1)
  gpu_eager_baseline the reference's own CUDA-eager bf16 path (baseline/_ref modules on CUDA tensors; oracle port if the
                     reference is not installed), same weights, same inputs               [rank 0, N = 1 only]
  parity             engine noise_pred vs that eager result at FULL depth, per config      [rank 0, N = 1 only]
  cpu_baseline       one whole reference step on the host cores                           [rank 0, N = 1 only]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {  # name -> (canvas h2 x w2 packed tokens, T, description, schedule length)
    "cfg2": (64, 32, 512, "FLUX.1-Fill-dev 12B, 512x512 scene + full-mask glyph concat (1024x512 canvas, S=2048, T=512)", 30),
    "cfg3": (72, 64, 512, "TextFlux-beta strip: 1024x1024 scene + 128-px glyph strip (1152x1024 canvas, S=4608, T=512)", 30),
    "cfg4": (64, 64, 512, "LoRA r16 folded, 1024x1024 (S=4096, T=512)", 30),
    "cfg5": (128, 64, 512, "multi-line full-mask 1024x2048 concat (S=8192, T=512)", 50),
}
HEADLINE = "cfg3"
METRIC = "denoising-steps/sec (device-timed) FLUX-Fill 12B 1024^2 glyph+scene"
UNIT = "steps/s"
CFG12B = dict(patch_size=1, in_channels=384, out_channels=64, num_layers=19, num_single_layers=38, attention_head_dim=128,
              num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True,
              axes_dims_rope=(16, 56, 56))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (measured)")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback")


def measured_traffic(workload):
    """DRAM bytes of one step's kernels from the newest committed ncu pass for this workload (profiles/*_traffic_<cfg>.json,
    tools/gpu_prof.sh + tools/traffic_from_ncu.py); None if no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_traffic_{workload}.json")))
    if not files and workload == "cfg2":
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1*_traffic.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        return {"bytes": d["dram_bytes_per_step"], "read": d["dram_read_bytes_per_step"], "write": d["dram_write_bytes_per_step"],
                "file": os.path.relpath(files[-1], ROOT)}
    except Exception:
        return None


def flops_per_step(S, T, L=19, Ls=38, D=3072):
    """Algorithmic FLOPs per sample-step (equals FlopCounterMode over the reference model)."""
    N = S + T
    return ((L + Ls) * (24 * N * D * D + 4 * N * N * D) + 2 * S * 384 * D + 2 * T * 4096 * D + 2 * S * D * 64
            + L * 2 * (2 * D * 6 * D) + Ls * (2 * D * 3 * D) + 2 * D * 2 * D
            + 3 * (2 * 256 * D + 2 * D * D) - 2 * 256 * D + 2 * 768 * D)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_steps(h2, w2, T, n_sched, warmup: int, steps: int, budget_s: float):
    """WHOLE denoising steps (19 + 38 blocks, D = 3072, + scheduler.step) of the reference on this box's host cores in
    bf16, with every host thread torch can use.  Runs `warmup` untimed and up to `steps` timed steps, stopping early once
    `budget_s` seconds of timed work are spent (at least one timed step).  Model: the unmodified reference modules from
    baseline/_ref when installed (kind "reference"), else the oracle port (kind "port"); in both cases the 19 double
    blocks alias ONE set of block weights and the 38 single blocks another (skips a minute of 12B random init; a step
    executes exactly the full-depth op sequence)."""
    torch.set_num_threads(os.cpu_count() or 1)
    S, B = h2 * w2, 1
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(B, S, 64, generator=g).to(torch.bfloat16)
    cond = torch.randn(B, S, 320, generator=g).to(torch.bfloat16)
    prompt = torch.randn(B, T, 4096, generator=g).to(torch.bfloat16)
    pooled = torch.randn(B, 768, generator=g).to(torch.bfloat16)
    guidance = torch.full([B], 30.0)
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] += torch.arange(h2)[:, None]
    ids[..., 2] += torch.arange(w2)[None, :]
    img_ids, txt_ids = ids.reshape(S, 3).to(torch.bfloat16), torch.zeros(T, 3, dtype=torch.bfloat16)
    from baseline import reference_arm as ra
    if ra.available():
        kind = "reference"
        model = ra.build_transformer_aliased(CFG12B, "cpu", torch.bfloat16)
        sch = ra.build_scheduler()
        sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n_sched, n_sched), device="cpu", mu=ra.calculate_shift(S))
        ref_step = ra.reference_step_fn(model, sch)

        def step(i, x):
            return ref_step(i, x, cond, prompt, pooled, guidance, txt_ids, img_ids)[0]
    else:
        kind = "port"
        from oracle import flux_oracle as fo
        small = fo.FluxConfig(num_layers=1, num_single_layers=1)
        sd1 = fo.init_state_dict(small, seed=7, dtype=torch.bfloat16)
        full = fo.FLUX_FILL_12B
        sd = dict(sd1)
        for k, v in sd1.items():
            if k.startswith("transformer_blocks.0."):
                for i in range(1, full.num_layers):
                    sd[k.replace("transformer_blocks.0.", f"transformer_blocks.{i}.", 1)] = v
            if k.startswith("single_transformer_blocks.0."):
                for i in range(1, full.num_single_layers):
                    sd[k.replace("single_transformer_blocks.0.", f"single_transformer_blocks.{i}.", 1)] = v
        sig, ts = fo.euler_set_timesteps(n_sched, S)

        @torch.no_grad()
        def step(i, x):
            t = ts[i].expand(B).to(torch.bfloat16) / 1000
            v = fo.flux_forward(sd, full, torch.cat((x, cond), dim=2), prompt, pooled, t, img_ids, txt_ids, guidance)
            return fo.euler_step(v, sig[i], sig[i + 1], x)

    x = lat
    for i in range(warmup):
        x = step(i % n_sched, x)
    times = []
    for i in range(max(1, steps)):
        t0 = time.perf_counter()
        x = step((warmup + i) % n_sched, x)
        times.append(time.perf_counter() - t0)
        if sum(times) >= budget_s:
            break
    step_s = statistics.mean(times)
    info = dict(value=1.0 / step_s, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                sample=(f"{len(times)} whole step(s) timed after {warmup} warm-up step(s): full-depth 19+38-block forward + scheduler.step at "
                        f"S={S}, T={T}, B=1, bf16, {step_s:.2f} s/step (per-step times {[round(t, 2) for t in times]}); the 19 double / "
                        f"38 single blocks alias one set of block weights each (random init of 12B parameters skipped)"))
    return info, step_s, len(times)


def run_reference(args, name):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    h2, w2, T, desc, n_sched = WORKLOADS[name]
    info, step_s, n_timed = cpu_reference_steps(h2, w2, T, n_sched, warmup=min(args.warmup, 1), steps=min(args.steps, 3),
                                                budget_s=args.cpu_budget)
    line = {"metric": METRIC, "value": 1.0 / step_s, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": n_timed, "warmup": min(args.warmup, 1), "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": 1, "image_tokens": h2 * w2, "text_tokens": T, "layers": [19, 38],
                       "note": "reference CPU path on the host cores, rank 0 only; whole steps, `steps` = the number actually timed"},
            "cpu_baseline": info,
            "e2e": {"value": 1.0 / step_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
class Inputs:
    """Synthetic inputs of one workload: per-sample seeds; prompt embeds + schedule come from rank 0."""

    def __init__(self, name, dev, rank, world):
        from textflux_b200 import B200FlowMatchEulerScheduler, calculate_shift
        h2, w2, T, desc, n_sched = WORKLOADS[name]
        self.name, self.h2, self.w2, self.T, self.desc, self.n_sched = name, h2, w2, T, desc, n_sched
        S, B = h2 * w2, 1
        self.S, self.B = S, B
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        self.latents0 = torch.randn(B, S, 64, generator=g, device=dev).to(torch.bfloat16)
        mil = torch.randn(B, S, 64, generator=g, device=dev).to(torch.bfloat16)
        mask = torch.zeros(B, h2, w2, 256, device=dev, dtype=torch.bfloat16)
        mask[:, h2 // 2:] = 1  # glyph part 0, scene part fully masked
        self.cond = torch.cat([mil, mask.reshape(B, S, 256)], dim=2).contiguous()
        g0 = torch.Generator(device=dev).manual_seed(999)
        prompt = torch.randn(1, T, 4096, generator=g0, device=dev).to(torch.bfloat16)
        pooled = torch.randn(1, 768, generator=g0, device=dev).to(torch.bfloat16)
        sch = B200FlowMatchEulerScheduler()
        self.mu = calculate_shift(S, sch.config.base_image_seq_len, sch.config.max_image_seq_len, sch.config.base_shift,
                                  sch.config.max_shift)
        sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n_sched, n_sched), device=dev, mu=self.mu)
        sig = sch.sigmas.clone()
        if world > 1:
            # the ONE collective of the job: text embeddings + pooled + guidance + sigma schedule from rank 0, packed
            # into a single NCCL broadcast over NVLink (textflux_b200/dist.py)
            from textflux_b200.dist import broadcast_conditioning
            prompt, pooled, sig, _ = broadcast_conditioning(prompt, pooled, sig, 30.0, src=0)
        self.sig_cpu = sig.tolist()
        self.ts = ((sig[:-1] * 1000)[:, None].expand(-1, B).to(torch.bfloat16) / 1000).contiguous()
        self.guidance = torch.full([B], 30.0, device=dev, dtype=torch.float32)
        ids = torch.zeros(h2, w2, 3)
        ids[..., 1] += torch.arange(h2)[:, None]
        ids[..., 2] += torch.arange(w2)[None, :]
        self.img_ids = ids.reshape(S, 3).to(dev, torch.bfloat16)
        self.txt_ids = torch.zeros(T, 3, device=dev, dtype=torch.bfloat16)
        self.prompt_b, self.pooled_b = prompt.expand(B, -1, -1).contiguous(), pooled.expand(B, -1).contiguous()


def parity_stats(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return {"rel_l2": ((a - b).norm() / b.norm()).item(),
            "cosine_dist": 1.0 - torch.nn.functional.cosine_similarity(a, b, dim=0).item(),
            "max_abs": (a - b).abs().max().item(), "ref_rms": b.pow(2).mean().sqrt().item()}


def run_ours(args, name):
    from textflux_b200 import B200FluxTransformer, B200FlowMatchEulerScheduler, synthetic_getter
    from textflux_b200.engine import FrozenConfig
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL writes its banner to stdout; keep stdout for the single JSON line by pointing fd 1 at stderr until then
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    full_depth = not (args.layers or args.single_layers)
    cfgd = dict(CFG12B, num_layers=args.layers or 19, num_single_layers=args.single_layers or 38)
    cfg = FrozenConfig(cfgd)
    getter = synthetic_getter(cfg, 1234, dev)
    eng_kw = {}
    if args.cta_group:
        eng_kw["gemm_cta_group"] = args.cta_group
    if args.mcast:
        eng_kw["gemm_mcast"] = args.mcast
    eng = B200FluxTransformer(cfg, getter, device=dev, **eng_kw)  # library defaults unless overridden on the command line
    for key, val in (("use_pdl", args.pdl), ("attn_variant", args.attn_variant or -1), ("attn_emu", args.attn_emu),
                     ("gemm_l2_hints", args.l2_hints), ("gemm_narrow_tiles", args.narrow_tiles), ("gemm_m_band", args.m_band),
                     ("gemm_k_snake", args.k_snake)):
        if val >= 0 or (key == "gemm_m_band" and val <= -100):
            eng.set_option(key, val)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_step(inp):
        def step(i, lat):
            # step i of an image's schedule; the schedule-wide modulation precompute (tfx_set_schedule, once per image)
            # is issued at every schedule start, i.e. INSIDE the timed region (the timed loop starts at k = 0)
            k = i % inp.n_sched
            if args.no_schedule:
                return eng.step(lat, inp.cond, inp.prompt_b, inp.pooled_b, inp.ts[k], inp.guidance, inp.img_ids, inp.txt_ids,
                                inp.sig_cpu[k], inp.sig_cpu[k + 1])
            if k == 0:
                eng.set_schedule(inp.ts, inp.guidance, inp.pooled_b, inp.S, inp.T)
            return eng.step_scheduled(k, lat, inp.cond, inp.prompt_b, inp.img_ids, inp.txt_ids, inp.sig_cpu[k], inp.sig_cpu[k + 1])
        return step

    def timed(fn, steps, warmup, lat0):
        """W untimed + K timed calls of fn(i, lat) bracketed by barrier + synchronize, CUDA events, max over ranks."""
        lat = lat0
        for i in range(warmup):
            lat = fn(i, lat)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            lat = fn(i, lat)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, lat

    # ============================================================ headline: device-timed, inputs resident in HBM
    inp = Inputs(name, dev, rank, world)
    S, T, B = inp.S, inp.T, inp.B
    step = make_step(inp)
    lat = inp.latents0
    for i in range(args.warmup):
        lat = step(i, lat)
    barrier()
    l0 = eng.counter("launches")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        lat = step(i, lat)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.counter("launches") - l0
    clocks = sampler.stop() if rank == 0 else None
    finite = bool(torch.isfinite(lat.float()).all())

    # ============================================================ end to end with HOST buffers (pinned)
    host = {k: v.cpu().pin_memory() for k, v in dict(lat=inp.latents0, cond=inp.cond, prompt=inp.prompt_b, pooled=inp.pooled_b,
                                                      ts=inp.ts, g=inp.guidance, img=inp.img_ids, txt=inp.txt_ids).items()}
    out_host = torch.empty_like(host["lat"]).pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("lat", "cond", "prompt", "pooled", "g", "img", "txt")) + B * 2
    d2h = out_host.numel() * out_host.element_size()

    def e2e_step(i):
        k = i % inp.n_sched
        d = {n: host[n].to(dev, non_blocking=True) for n in ("lat", "cond", "prompt", "pooled", "g", "img", "txt")}
        t = host["ts"][k].to(dev, non_blocking=True)
        if args.no_schedule:
            new = eng.step(d["lat"], d["cond"], d["prompt"], d["pooled"], t, d["g"], d["img"], d["txt"], inp.sig_cpu[k], inp.sig_cpu[k + 1])
        else:
            if k == 0:
                eng.set_schedule(host["ts"].to(dev, non_blocking=True), d["g"], d["pooled"], S, T)
            new = eng.step_scheduled(k, d["lat"], d["cond"], d["prompt"], d["img"], d["txt"], inp.sig_cpu[k], inp.sig_cpu[k + 1])
        out_host.copy_(new, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host["lat"].copy_(out_host)

    for i in range(min(args.warmup, 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # ============================================================ drop-in: what the UNMODIFIED pipeline calls
    sch = B200FlowMatchEulerScheduler()

    def dropin_step(i, x):
        # pipeline_flux_fill.py:2058-2064 (once per image) and :2077-2098 (per step), verbatim call sequence
        k = i % inp.n_sched
        if k == 0:
            sch.set_timesteps(sigmas=np.linspace(1.0, 1 / inp.n_sched, inp.n_sched), device=dev, mu=inp.mu)
        t = sch.timesteps[k]
        timestep = t.expand(x.shape[0]).to(x.dtype)
        v = eng(hidden_states=torch.cat((x, inp.cond), dim=2), timestep=timestep / 1000, guidance=inp.guidance,
                pooled_projections=inp.pooled_b, encoder_hidden_states=inp.prompt_b, txt_ids=inp.txt_ids, img_ids=inp.img_ids,
                joint_attention_kwargs=None, return_dict=False)[0]
        return sch.step(v, t, x, return_dict=False)[0]

    eng.set_option("mod_cache_reset", 1)
    hits0 = eng.counter("mod_cache_hits")
    dms_cold, _ = timed(dropin_step, args.steps, 0, inp.latents0)        # first image: every (t, g, pooled) triple is new
    dms_warm, lat_d = timed(dropin_step, args.steps, 0, inp.latents0)    # following images: modulation vectors cached
    dropin = {"value": args.steps * B * world / (dms_warm / 1e3), "unit": UNIT, "ms_per_step": dms_warm / args.steps,
              "first_image_value": args.steps * B * world / (dms_cold / 1e3), "first_image_ms_per_step": dms_cold / args.steps,
              "mod_cache_hits": eng.counter("mod_cache_hits") - hits0,
              "path": "B200FluxTransformer.forward (tfx_forward, one graph launch) + B200FlowMatchEulerScheduler.step (tfx_euler_step) "
                      "driven exactly as pipeline_flux_fill.py:2077-2098 does, incl. torch.cat(latents, cond) per step; "
                      "first image computes the adaLN vectors per step (6.5 GB GEMV), later images hit the device-side cache"}

    # ============================================================ per-kernel-family device time, live (eager launches)
    fam = {}
    if rank == 0:
        eng.set_option("profile", 1)
        lat2 = step(1, inp.latents0)
        torch.cuda.synchronize(dev)
        for f in ("gemm", "attn", "ln", "gemv", "misc"):
            fam[f] = {"us": eng.counter(f"prof_us_{f}"), "launches": eng.counter(f"prof_n_{f}")}
        eng.set_option("profile", 0)
        del lat2

    if world > 1:
        tmax = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_ms = tmax.tolist()
        e2e_s = e2e_ms / 1e3

    # ============================================================ the other BASELINE configurations, same run
    pk = peaks()
    L = cfg.num_layers + cfg.num_single_layers
    eager = None
    if rank == 0 and world == 1 and not args.no_eager and full_depth:
        eager = EagerReference(getter, dev)

    def config_entry(cin, ms_c, steps_c):
        fl = flops_per_step(cin.S, cin.T, cfg.num_layers, cfg.num_single_layers)
        ach = fl * cin.B * (steps_c / (ms_c / 1e3)) / 1e12
        return {"workload": cin.desc, "value": steps_c * cin.B * world / (ms_c / 1e3), "unit": UNIT, "ms_per_step": ms_c / steps_c,
                "steps": steps_c, "image_tokens": cin.S, "tflops_per_gpu": ach, "roofline_frac": ach / pk["sustained"]}

    def parity_for(cin):
        # engine noise_pred vs the reference's CUDA-eager bf16 forward, same weights, same inputs, full depth, step 0
        hs = torch.cat((cin.latents0, cin.cond), dim=2)
        v_eng = eng(hidden_states=hs, timestep=cin.ts[0], guidance=cin.guidance, pooled_projections=cin.pooled_b,
                    encoder_hidden_states=cin.prompt_b, txt_ids=cin.txt_ids, img_ids=cin.img_ids, return_dict=False)[0]
        v_ref = eager.forward(hs, cin.prompt_b, cin.pooled_b, cin.ts[0], cin.img_ids, cin.txt_ids, cin.guidance)
        return parity_stats(v_eng, v_ref)

    configs = {}
    others = [] if args.no_configs else [c for c in ("cfg2", "cfg4", "cfg5") if c != name]
    parity = None
    if eager is not None:
        parity = parity_for(inp)
        parity["against"] = eager.kind
    for cname in others:
        cin = Inputs(cname, dev, rank, world)
        ms_c, _ = timed(make_step(cin), args.config_steps, 3, cin.latents0)
        configs[cname] = config_entry(cin, ms_c, args.config_steps)
        if eager is not None:
            configs[cname]["parity"] = parity_for(cin)

    # ============================================================ the reference's CUDA-eager path, timed (headline workload)
    gpu_eager = None
    if eager is not None:
        n_e = max(1, min(args.steps, 5))
        gpu_eager = eager.time_steps(inp, n_e)
        # how far the reference's own bf16 result is from its fp32 result on these inputs: the floor parity is judged by
        if not args.no_fp32_floor:
            try:
                hs = torch.cat((inp.latents0, inp.cond), dim=2)
                v16 = eager.forward(hs, inp.prompt_b, inp.pooled_b, inp.ts[0], inp.img_ids, inp.txt_ids, inp.guidance)
                v32 = eager.forward_fp32(hs, inp.prompt_b, inp.pooled_b, inp.ts[0], inp.img_ids, inp.txt_ids, inp.guidance)
                v_eng = eng(hidden_states=hs, timestep=inp.ts[0], guidance=inp.guidance, pooled_projections=inp.pooled_b,
                            encoder_hidden_states=inp.prompt_b, txt_ids=inp.txt_ids, img_ids=inp.img_ids, return_dict=False)[0]
                parity["reference_bf16_vs_fp32"] = parity_stats(v16, v32)
                parity["engine_vs_fp32"] = parity_stats(v_eng, v32)
            except Exception as e:  # fp32 eager needs ~50 GB more; never let the floor take the bench line down
                parity["reference_bf16_vs_fp32"] = {"error": repr(e)[:200]}
        eager.release()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    once = None if args.no_image_stages else once_per_image(inp, dev, ms / args.steps)

    total_steps = args.steps * B * world
    value = total_steps / (ms / 1e3)
    flops = flops_per_step(S, T, cfg.num_layers, cfg.num_single_layers)
    N = S + T
    D = 3072
    achieved = flops * B * (args.steps / (ms / 1e3)) / 1e12  # per GPU
    tr = measured_traffic(name) if B == 1 and full_depth else None
    roof = {"bound": "tensor", "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["sustained"], "frac_of_burst": achieved / pk["burst"],
            "traffic": tr["bytes"] if tr else None,
            "traffic_note": (f"DRAM read {tr['read'] / 1e9:.1f} GB + write {tr['write'] / 1e9:.1f} GB per launch (= one step), ncu pass in "
                             f"{tr['file']}; algorithmic minimum 17.3 GB of weights (23.8 GB incl. the adaLN matrix) + inputs/outputs") if tr else None,
            "launch": "one sampling step (one CUDA-graph launch: every kernel of forward + fused Euler)",
            "flops_per_launch": flops * B, "peak_source": pk["source"] + ", sustained bf16 figure (timed inside a long step)",
            "kernel_families_us": fam}
    if fam.get("gemm", {}).get("us"):
        gemm_flops = flops - L * 4 * N * N * D
        roof["gemm_kernel"] = {"achieved": gemm_flops / (fam["gemm"]["us"] * 1e-6) / 1e12, "unit": "TFLOP/s",
                               "launches": fam["gemm"]["launches"],
                               "note": "tcgen05 GEMM family alone, eager launches timed with CUDA events"}
        if fam.get("attn", {}).get("us"):
            roof["attention_kernel"] = {"achieved": L * 4 * N * N * D / (fam["attn"]["us"] * 1e-6) / 1e12, "unit": "TFLOP/s",
                                        "launches": fam["attn"]["launches"]}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu, _, _ = cpu_reference_steps(inp.h2, inp.w2, T, inp.n_sched, warmup=0, steps=1, budget_s=0.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init 12B weights, seeded latents/cond/prompt embeds)",
            "config": {"workload": inp.desc, "batch_per_gpu": B, "global_batch": B * world, "image_tokens": S, "text_tokens": T,
                       "layers": [cfg.num_layers, cfg.num_single_layers], "parallelism": f"replica x{world}",
                       "l2": "23.8 GB of weights stream through the 126 MB L2 every step (inputs larger than L2)",
                       "engine_options": "library defaults" if not eng_kw else eng_kw,
                       "modulation": "per step" if args.no_schedule else "adaLN table of the schedule filled in 8-step passes when first needed, inside the timed region"},
            "clocks": clocks, "gpu_launches": launches, "finite": finite,
            "e2e": {"value": total_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roof, "dropin": dropin, "configs": configs, "parity": parity, "gpu_eager_baseline": gpu_eager,
            "once_per_image": once, "cpu_baseline": cpu}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def once_per_image(inp, dev, ms_per_step):
    """The stages FluxFillPipeline runs once per image around the loop, on the engine, NOT part of `value`:
    vae.encode of the masked canvas (:1528), vae.decode (:2128), T5-XXL and CLIP-L prompt encoding (:1438, :1483).  Synthetic weights of
    the real shapes, CUDA events, median of 5."""
    from textflux_b200 import B200AutoencoderKL, B200CLIPTextEncoder, B200T5Encoder
    from textflux_b200.text_encoders import CLIP_L_CONFIG, T5_XXL_CONFIG, clip_reference_names, t5_reference_names
    from textflux_b200.vae import FLUX_VAE_CONFIG, synthetic_state, vae_reference_names

    def med(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    out = {}
    try:
        H, W = inp.h2 * 16, inp.w2 * 16
        vae = B200AutoencoderKL.from_state_dict(FLUX_VAE_CONFIG, synthetic_state(vae_reference_names(FLUX_VAE_CONFIG), 31, dev), device=dev)
        g = torch.Generator(device=dev).manual_seed(5)
        image = (torch.rand(1, 3, H, W, generator=g, device=dev) * 2 - 1).to(torch.bfloat16)
        z = torch.randn(1, 16, H // 8, W // 8, generator=g, device=dev).to(torch.bfloat16)
        out["canvas"] = f"{H}x{W}"
        out["vae_encode_ms"], out["vae_decode_ms"] = med(lambda: vae.encode(image)), med(lambda: vae.decode(z))
        out["vae_finite"] = bool(torch.isfinite(vae.decode(z, return_dict=False)[0].float()).all())
        out["vae_launches"] = vae.counter("launches")
        del vae
        t5 = B200T5Encoder(T5_XXL_CONFIG, synthetic_state(t5_reference_names(T5_XXL_CONFIG), 7, dev).__getitem__, device=dev)
        ids = torch.randint(2, 32128, (1, 512), device=dev)
        ids[:, 60:] = 0
        out["t5_xxl_ms"] = med(lambda: t5(ids))
        out["t5_launches"] = t5.counter("launches")
        del t5
        clip = B200CLIPTextEncoder(CLIP_L_CONFIG, synthetic_state(clip_reference_names(CLIP_L_CONFIG), 8, dev).__getitem__, device=dev, cache=False)
        cids = torch.randint(3, 49406, (1, 77), device=dev)
        cids[:, 30:] = 49407
        out["clip_l_ms"] = med(lambda: clip(cids))
        del clip
        torch.cuda.empty_cache()
        stages = out["vae_encode_ms"] + out["vae_decode_ms"] + out["t5_xxl_ms"] + out["clip_l_ms"]
        out["stages_ms"] = stages
        out["share_of_a_30_step_image"] = stages / (stages + 30 * ms_per_step)
    except Exception as e:  # never let the side stages take the bench line down
        out["error"] = repr(e)[:300]
    return out


class EagerReference:
    """The reference's own CUDA-eager bf16 path with the engine's weights: the unmodified FluxTransformer2DModel +
    FlowMatchEulerDiscreteScheduler of baseline/_ref on CUDA tensors (kind "reference"); where the reference is not
    installed, the oracle port, which issues the same ATen calls (kind "port").  Checker and reported baseline only."""

    def __init__(self, getter, dev):
        from baseline import reference_arm as ra
        self.dev = dev
        self.ra = ra if ra.available() else None
        if self.ra is not None:
            self.kind = "reference (baseline/_ref FluxTransformer2DModel on CUDA, eager bf16)"
            self.model = ra.build_transformer(CFG12B, getter, dev, torch.bfloat16)
            self.sd = None
        else:
            from oracle import flux_oracle as fo
            from textflux_b200 import reference_names
            self.kind = "port (oracle/flux_oracle.py on CUDA tensors: the reference's ATen call sequence)"
            self.fo = fo
            self.sd = {n: getter(n) for n, _ in reference_names(fo.FLUX_FILL_12B)}
            self.model = None

    @torch.no_grad()
    def forward(self, hs, enc, pooled, t, img_ids, txt_ids, guidance):
        if self.model is not None:
            return self.model(hidden_states=hs, timestep=t, guidance=guidance, pooled_projections=pooled, encoder_hidden_states=enc,
                              txt_ids=txt_ids, img_ids=img_ids, joint_attention_kwargs=None, return_dict=False)[0]
        return self.fo.flux_forward(self.sd, self.fo.FLUX_FILL_12B, hs, enc, pooled, t, img_ids, txt_ids, guidance)

    @torch.no_grad()
    def forward_fp32(self, hs, enc, pooled, t, img_ids, txt_ids, guidance):
        """fp32 forward on the same bf16-rounded weights / inputs, fed the t*1000 and g*1000 the bf16 model actually sees
        (SURVEY.md §8d tolerance calibration)."""
        t32 = (t.to(torch.bfloat16) * 1000).float() / 1000
        g32 = (guidance.to(torch.bfloat16) * 1000).float() / 1000
        args = (hs.float(), enc.float(), pooled.float(), t32, img_ids.float(), txt_ids.float(), g32)
        if self.model is not None:
            self.model.to(torch.float32)
            try:
                out = self.model(hidden_states=args[0], timestep=args[3], guidance=args[6], pooled_projections=args[2],
                                 encoder_hidden_states=args[1], txt_ids=args[5], img_ids=args[4], return_dict=False)[0]
            finally:
                self.model.to(torch.bfloat16)  # bf16 -> fp32 -> bf16 is the identity on bf16-representable weights
            return out

        class F32(dict):
            def __init__(s, sd):
                s.sd = sd

            def __getitem__(s, k):
                return s.sd[k].float()

            def __contains__(s, k):
                return k in s.sd
        return self.fo.flux_forward(F32(self.sd), self.fo.FLUX_FILL_12B, *args)

    @torch.no_grad()
    def time_steps(self, inp, n):
        dev = self.dev
        if self.model is not None:
            sch = self.ra.build_scheduler()
            sch.set_timesteps(sigmas=np.linspace(1.0, 1 / inp.n_sched, inp.n_sched), device=dev, mu=inp.mu)
            ref_step = self.ra.reference_step_fn(self.model, sch)

            def step(i, x):
                return ref_step(i, x, inp.cond, inp.prompt_b, inp.pooled_b, inp.guidance, inp.txt_ids, inp.img_ids)[0]
        else:
            fo = self.fo
            sig, ts = fo.euler_set_timesteps(inp.n_sched, inp.S)
            sig, ts = sig.to(dev), ts.to(dev)

            def step(i, x):
                t = ts[i].expand(x.shape[0]).to(torch.bfloat16) / 1000
                v = self.forward(torch.cat((x, inp.cond), dim=2), inp.prompt_b, inp.pooled_b, t, inp.img_ids, inp.txt_ids, inp.guidance)
                return fo.euler_step(v, sig[i], sig[i + 1], x)
        x = step(0, inp.latents0)  # warm-up (cuBLAS heuristics, SDPA backend selection, allocator)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            x = step(1 + i, x)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / n
        fl = flops_per_step(inp.S, inp.T)
        return {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": n, "kind": self.kind,
                "tflops": fl / (ms * 1e-3) / 1e12, "workload": inp.desc,
                "note": "the reference's stock CUDA path (cuBLAS addmm + SDPA, ~22.8k eager ATen dispatches per step), same weights and inputs as the engine"}

    def release(self):
        self.model = None
        self.sd = None
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30, help="timed steps; 30 = one whole schedule of the benchmark config")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--config-steps", type=int, default=10, help="timed steps of each secondary configuration")
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="--impl reference: stop timing whole CPU steps after this many seconds")
    ap.add_argument("--cta-group", type=int, default=0, help="override the library's GEMM cta_group (1, 2)")
    ap.add_argument("--mcast", type=int, default=0, help="CTA pairs per cluster sharing A by TMA multicast (0, 2, 4)")
    ap.add_argument("--layers", type=int, default=0, help="debug: override the 19 double blocks (invalidates the number)")
    ap.add_argument("--single-layers", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=-1, help="override programmatic dependent launch (0/1)")
    ap.add_argument("--attn-variant", type=int, default=0, help="override the attention schedule")
    ap.add_argument("--attn-emu", type=int, default=-1, help="override: softmax column pairs per 8 on the FMA-pipe exp2 (0, 2, 3, 4)")
    ap.add_argument("--narrow-tiles", type=int, default=-1, help="override: allow 224-wide GEMM tiles (0/1)")
    ap.add_argument("--k-snake", type=int, default=-1, help="override GemmParams::k_snake on the banded wide-K GEMMs (0 | 1)")
    ap.add_argument("--m-band", type=int, default=-1, help="override the tile order of the wide-K GEMMs (0 = M-fastest, b = bands of b M tiles)")
    ap.add_argument("--l2-hints", type=int, default=-1, help="override the GEMM L2 eviction hints (0..3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the CUDA-eager reference baseline and the parity legs")
    ap.add_argument("--no-fp32-floor", action="store_true")
    ap.add_argument("--no-image-stages", action="store_true", help="skip the once-per-image stages (VAE, prompt encoders)")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--no-schedule", action="store_true", help="recompute the adaLN modulation every step inside the fused step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
