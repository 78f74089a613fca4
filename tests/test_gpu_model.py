"""GPU parity of the whole hot path (engine through the C ABI) against the committed reference fixtures and the
CPU oracle.  Tolerances follow SURVEY.md §8d: the engine may differ from the reference's bf16 result by at most
2x what the reference's own bf16 result differs from its fp32 result on the same inputs, and must be at least
as close to fp32 (x1.5); cosine distance of the outputs < 1e-4 (the reference's own slow-test criterion,
tests/pipelines/flux/test_pipeline_flux.py:296-298)."""
import numpy as np
import pytest
import torch

from oracle import flux_oracle as fo

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def _cosdist(a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    return 1.0 - torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def _engine(cfg, sd, **kw):
    from textflux_b200 import B200FluxTransformer
    return B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0", **kw)


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("cta_group,attn,graph,mcast", [(1, 4, False, 0), (1, 5, True, 0), (2, 8, True, 0), (2, 5, False, 0),
                                                        (2, 9, True, 2), (2, 9, False, 4), (1, 9, True, 0)])
def test_tiny_forward_vs_reference_golden(golden, cta_group, attn, graph, mcast):
    g = golden("tiny_forward.pt")
    cfg = fo.FluxConfig(**g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"])
    eng = _engine(cfg, sd, gemm_cta_group=cta_group, use_graph=graph, gemm_mcast=mcast)
    eng.set_option("attn_variant", attn)  # all attention schedules
    eng.set_option("use_pdl", 0 if (cta_group == 2 and attn == 5) else 1)  # with and without programmatic dependent launch
    inp = _cuda(g["inputs"])
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    for _ in range(2):  # second call replays the captured graph
        out = eng(hidden_states=hs, timestep=g["timestep"].cuda(), guidance=g["guidance"].cuda(),
                  pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"],
                  img_ids=inp["img_ids"], joint_attention_kwargs=None, return_dict=False)[0]
        torch.cuda.synchronize()
        assert out.shape == g["sample"].shape and out.dtype == torch.bfloat16
        ref16, ref32 = g["sample"], g["sample_fp32"]
        base = _rel(ref16, ref32)
        e16, e32 = _rel(out, ref16), _rel(out, ref32)
        print(f"tiny forward cta_group={cta_group} attn_variant={attn}: ref16-vs-fp32 {base:.3e} engine-vs-ref16 {e16:.3e} "
              f"engine-vs-fp32 {e32:.3e} cosdist {_cosdist(out, ref16):.2e}")
        assert torch.isfinite(out.float()).all()
        assert e16 <= 2.0 * base, (e16, base)
        assert e32 <= 1.5 * base, (e32, base)
        assert _cosdist(out, ref16) < 1e-4
    assert eng.counter("launches") > 0


def test_tiny_loop_dropin_and_fused(golden):
    """4 Euler steps: (a) forward + scheduler.step exactly as pipeline_flux_fill.py:2077-2098 drives them,
    (b) the fused one-launch-per-step path.  Both must agree bit-for-bit with each other and track the reference."""
    from textflux_b200 import B200FlowMatchEulerScheduler, calculate_shift
    g = golden("tiny_loop.pt")
    cfg = fo.FluxConfig(**g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"])
    eng = _engine(cfg, sd)
    inp = _cuda(g["inputs"])
    n = g["steps"]
    sch = B200FlowMatchEulerScheduler()
    mu = calculate_shift(64, sch.config.base_image_seq_len, sch.config.max_image_seq_len, sch.config.base_shift,
                         sch.config.max_shift)
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n, n), device="cuda", mu=mu)
    assert torch.equal(sch.timesteps.cpu(), g["timesteps"]) and torch.equal(sch.sigmas.cpu(), g["sigmas"])
    latents = inp["latents"]
    guidance = torch.full([1], g["guidance_scale"], device="cuda", dtype=torch.float32).expand(1)
    lat_a = []
    for i, t in enumerate(sch.timesteps):
        timestep = t.expand(1).to(latents.dtype)
        v = eng(hidden_states=torch.cat((latents, inp["cond"]), dim=2), timestep=timestep / 1000, guidance=guidance,
                pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"],
                img_ids=inp["img_ids"], joint_attention_kwargs=None, return_dict=False)[0]
        if i == 0:
            base = 5.7e-3  # reference bf16-vs-fp32 rel-L2 on the tiny model
            assert _rel(v, g["noise_preds"][0]) < 2 * base
        latents = sch.step(v, t, latents, return_dict=False)[0]
        lat_a.append(latents)
    seen = []
    final = eng.denoise(inp["latents"], inp["cond"], inp["prompt_embeds"], inp["pooled"], inp["txt_ids"], inp["img_ids"],
                        g["guidance_scale"], n, callback=lambda i, x: seen.append(x.clone()))
    torch.cuda.synchronize()
    for a, b in zip(lat_a, seen):
        assert torch.equal(a, b)
    assert torch.equal(final, lat_a[-1])
    # per-step modulation (no schedule precompute) gives the same bits
    final2 = eng.denoise(inp["latents"], inp["cond"], inp["prompt_embeds"], inp["pooled"], inp["txt_ids"], inp["img_ids"],
                         g["guidance_scale"], n, precompute_modulation=False)
    assert torch.equal(final2, final)
    for i in range(n):
        assert _cosdist(lat_a[i], g["latents"][i]) < 1e-4, i
        assert _rel(lat_a[i], g["latents"][i]) < 2e-2, i


def test_real_dim_blocks_vs_reference_golden(golden):
    """One double + one single block at the real FLUX dims (D=3072, 24 heads x 128, rope 16/56/56)."""
    g = golden("real_dim_blocks.pt")
    cfg = fo.FluxConfig(**g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"])
    inp = fo.synthetic_inputs(cfg, *g["grid"], g["T"], batch=1, seed0=g["input_seed0"])
    sd32 = {k: v.float() for k, v in sd.items()}
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    t32 = (g["timestep"].to(torch.bfloat16) * 1000).float() / 1000
    g32 = (g["guidance"].to(torch.bfloat16) * 1000).float() / 1000
    with torch.no_grad():
        ref32 = fo.flux_forward(sd32, cfg, hs.float(), inp["prompt_embeds"].float(), inp["pooled"].float(), t32,
                                inp["img_ids"].float(), inp["txt_ids"].float(), g32)
    base = _rel(g["sample"], ref32)
    for cta_group, mcast in ((1, 0), (2, 0), (2, 2), (2, 4)):
        eng = _engine(cfg, sd, gemm_cta_group=cta_group, gemm_mcast=mcast)
        d = _cuda(inp)
        out = eng(hidden_states=hs.cuda(), timestep=g["timestep"].cuda(), guidance=g["guidance"].cuda(),
                  pooled_projections=d["pooled"], encoder_hidden_states=d["prompt_embeds"], txt_ids=d["txt_ids"],
                  img_ids=d["img_ids"], return_dict=False)[0]
        torch.cuda.synchronize()
        e16, e32 = _rel(out, g["sample"]), _rel(out, ref32)
        print(f"real-dim blocks cta_group={cta_group} mcast={mcast}: ref16-vs-fp32 {base:.3e} engine-vs-ref16 {e16:.3e} engine-vs-fp32 {e32:.3e}")
        assert e16 <= 2.0 * base and e32 <= 1.5 * base, (e16, e32, base)
        assert _cosdist(out, g["sample"]) < 1e-4
        del eng


def test_batch_rows_are_independent(golden):
    """Samples of a batch never interact: B=2 equals two B=1 calls bit-for-bit."""
    g = golden("tiny_forward.pt")
    cfg = fo.FluxConfig(**g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"])
    eng = _engine(cfg, sd)
    inp = _cuda(g["inputs"])
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    kw = dict(txt_ids=inp["txt_ids"], img_ids=inp["img_ids"], return_dict=False)
    both = eng(hidden_states=hs, timestep=g["timestep"].cuda(), guidance=g["guidance"].cuda(),
               pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"], **kw)[0]
    for b in range(2):
        one = eng(hidden_states=hs[b:b + 1], timestep=g["timestep"][b:b + 1].cuda(), guidance=g["guidance"][b:b + 1].cuda(),
                  pooled_projections=inp["pooled"][b:b + 1], encoder_hidden_states=inp["prompt_embeds"][b:b + 1], **kw)[0]
        assert torch.equal(one[0], both[b])


def test_full_width_reduced_depth_vs_oracle():
    """BASELINE config 2 shape (S=2048, T=512, D=3072) with 2+2 blocks: the oracle finishes this in seconds on CPU."""
    cfg = fo.FluxConfig(num_layers=2, num_single_layers=2)
    sd = fo.init_state_dict(cfg, seed=21)
    inp = fo.synthetic_inputs(cfg, 64, 32, 512, batch=1, seed0=4000)
    t = (torch.tensor([871.3]).to(torch.bfloat16) / 1000)
    gd = torch.full([1], 30.0)
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    with torch.no_grad():
        ref16 = fo.flux_forward(sd, cfg, hs, inp["prompt_embeds"], inp["pooled"], t, inp["img_ids"], inp["txt_ids"], gd)
    eng = _engine(cfg, sd)
    d = _cuda(inp)
    out = eng(hidden_states=hs.cuda(), timestep=t.cuda(), guidance=gd.cuda(), pooled_projections=d["pooled"],
              encoder_hidden_states=d["prompt_embeds"], txt_ids=d["txt_ids"], img_ids=d["img_ids"], return_dict=False)[0]
    out2 = eng(hidden_states=hs.cuda(), timestep=t.cuda(), guidance=gd.cuda(), pooled_projections=d["pooled"],
               encoder_hidden_states=d["prompt_embeds"], txt_ids=d["txt_ids"], img_ids=d["img_ids"], return_dict=False)[0]
    torch.cuda.synchronize()
    assert torch.equal(out, out2)  # deterministic
    e16 = _rel(out, ref16)
    print(f"full-width 2+2: engine-vs-oracle-bf16 rel-L2 {e16:.3e} cosdist {_cosdist(out, ref16):.2e}")
    assert e16 < 1.2e-2 and _cosdist(out, ref16) < 1e-4


def test_full_size_12b_properties():
    """BASELINE.json configs[1] at FULL size (19 + 38 blocks, D = 3072, S = 2048, T = 512; random 12B weights made on the
    device): the oracle cannot run this in seconds, so parity is checked through size-independent properties --
    determinism, forward + scheduler.step == the fused step == the scheduled step (bit-exact), batch rows independent,
    and equivariance under a permutation of the image tokens (rows and img_ids permuted together), which exercises the
    token indexing / RoPE / attention row order end to end."""
    from textflux_b200 import B200FlowMatchEulerScheduler, B200FluxTransformer, calculate_shift, synthetic_getter
    cfg = fo.FLUX_FILL_12B
    dev = torch.device("cuda", 0)
    eng = B200FluxTransformer(cfg.to_dict(), synthetic_getter(cfg, 4321, dev), device=dev)
    h2, w2, T, S = 64, 32, 512, 2048
    inp = _cuda(fo.synthetic_inputs(cfg, h2, w2, T, batch=2, seed0=7000))
    sch = B200FlowMatchEulerScheduler()
    mu = calculate_shift(S, sch.config.base_image_seq_len, sch.config.max_image_seq_len, sch.config.base_shift, sch.config.max_shift)
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / 30, 30), device="cuda", mu=mu)
    t0 = sch.timesteps[0]
    ts = (t0.expand(2).to(torch.bfloat16) / 1000)
    gd = torch.full([2], 30.0, device="cuda")
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    kw = dict(txt_ids=inp["txt_ids"], return_dict=False)

    def fwd(sl, img_ids, hidden):
        return eng(hidden_states=hidden, timestep=ts[sl], guidance=gd[sl], pooled_projections=inp["pooled"][sl],
                   encoder_hidden_states=inp["prompt_embeds"][sl], img_ids=img_ids, **kw)[0]

    one = slice(0, 1)
    v = fwd(one, inp["img_ids"], hs[one])
    assert torch.isfinite(v.float()).all() and v.shape == (1, S, 64)
    assert torch.equal(v, fwd(one, inp["img_ids"], hs[one]))                                   # deterministic (graph replay)
    both = fwd(slice(0, 2), inp["img_ids"], hs)
    # batch rows independent.  The default attention schedule may cut the last partial wave of (head, query-pair) units into
    # KV shares whose boundaries depend on the number of units, i.e. on the batch size: same math, different fp32 summation
    # order, so B = 2 tracks B = 1 to rounding (amplified through 57 random blocks) instead of bit-for-bit ...
    print(f"12B full size: B=2 row 0 vs B=1 rel-L2 {_rel(both[0], v[0]):.3e}")
    assert _rel(both[0], v[0]) < 5e-2 and _cosdist(both[0], v[0]) < 1e-3
    eng.set_option("attn_variant", 5)  # ... and bit-for-bit under the schedule with a fixed summation order
    assert torch.equal(fwd(slice(0, 2), inp["img_ids"], hs)[0], fwd(one, inp["img_ids"], hs[one])[0])
    eng.set_option("attn_variant", 0)
    # forward + scheduler.step  ==  fused step  ==  scheduled step
    sig = sch.sigmas_cpu
    x_ref = sch.step(v, t0, inp["latents"][one], return_dict=False)[0]
    x_fused, v_fused = eng.step(inp["latents"][one], inp["cond"][one], inp["prompt_embeds"][one], inp["pooled"][one], ts[one], gd[one],
                                inp["img_ids"], inp["txt_ids"], sig[0], sig[1], return_noise_pred=True)
    assert torch.equal(v_fused, v) and torch.equal(x_fused, x_ref)
    tall = (sch.timesteps[:, None].expand(-1, 1).to(torch.bfloat16) / 1000).contiguous()
    eng.set_schedule(tall, gd[one], inp["pooled"][one], S, T)
    x_sched = eng.step_scheduled(0, inp["latents"][one], inp["cond"][one], inp["prompt_embeds"][one], inp["img_ids"], inp["txt_ids"],
                                 sig[0], sig[1])
    assert torch.equal(x_sched, x_ref)
    # permutation equivariance over image tokens (exact in real arithmetic; here up to the summation order of attention)
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(3)).cuda()
    vp = fwd(one, inp["img_ids"][perm], hs[one][:, perm])
    rel = _rel(vp, v[:, perm])
    print(f"12B full size: permutation equivariance rel-L2 {rel:.3e}, cosdist {_cosdist(vp, v[:, perm]):.2e}")
    # measured 1.4e-2 / 1e-4: bf16 rounding differences of the re-ordered attention sums, amplified through 57 random blocks;
    # an indexing error (wrong row order, wrong RoPE row) gives rel-L2 ~ 1
    assert rel < 5e-2 and _cosdist(vp, v[:, perm]) < 1e-3


def test_schedule_precompute_batch_and_long_schedule():
    """tfx_set_schedule over 11 steps x 3 samples (33 rows -> 5 GEMV passes) equals the per-step path bit-for-bit."""
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=9)
    eng = _engine(cfg, sd)
    inp = _cuda(fo.synthetic_inputs(cfg, 4, 6, 16, batch=3, seed0=50))
    a = eng.denoise(inp["latents"], inp["cond"], inp["prompt_embeds"], inp["pooled"], inp["txt_ids"], inp["img_ids"], 3.5, 11)
    b = eng.denoise(inp["latents"], inp["cond"], inp["prompt_embeds"], inp["pooled"], inp["txt_ids"], inp["img_ids"], 3.5, 11,
                    precompute_modulation=False)
    torch.cuda.synchronize()
    assert torch.isfinite(a.float()).all() and torch.equal(a, b)
    with pytest.raises(Exception):
        eng.step_scheduled(11, inp["latents"], inp["cond"], inp["prompt_embeds"], inp["img_ids"], inp["txt_ids"], 0.5, 0.4)


def test_overshoot_scheduler_bit_exact_vs_reference_golden(golden):
    """TextFlux's default sampler: 6 steps of the reference StochasticRFOvershotDiscreteScheduler (CPU generator seed
    777) reproduced bit-for-bit by the CUDA step, including predicted_x1."""
    from textflux_b200 import B200StochasticRFOvershotScheduler, calculate_shift
    d = golden("overshoot.pt")
    sch = B200StochasticRFOvershotScheduler()
    sch.set_c(2.0)
    sch.set_overshot_func(lambda t, dt: t + dt)
    mu = calculate_shift(d["S"], 256, 4096, 0.5, 1.15)
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / d["n"], d["n"]), device="cuda", mu=mu)
    g = torch.Generator().manual_seed(d["input_seed"])
    x = torch.randn(2, d["S"], 64, generator=g).to(torch.bfloat16).cuda()
    vs = [torch.randn(2, d["S"], 64, generator=g).to(torch.bfloat16).cuda() for _ in range(d["n"])]
    gen = torch.Generator().manual_seed(d["noise_seed"])
    for i, t in enumerate(sch.timesteps):
        x, x1 = sch.step(vs[i], t, x, generator=gen, return_dict=False)
        assert torch.equal(x.cpu(), d["prev_samples"][i]), i
        assert torch.equal(x1.cpu(), d["predicted_x1"][i]), i
    assert sch.step_index == d["n"]


def test_engine_input_validation():
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=1)
    eng = _engine(cfg, sd)
    inp = _cuda(fo.synthetic_inputs(cfg, 4, 4, 16))
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    with pytest.raises(ValueError):
        eng(hidden_states=hs[..., :100], timestep=torch.zeros(1, device="cuda"), guidance=torch.ones(1, device="cuda"),
            pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"],
            img_ids=inp["img_ids"])
    with pytest.raises(ValueError):
        eng(hidden_states=hs, timestep=torch.zeros(1, device="cuda"), guidance=None, pooled_projections=inp["pooled"],
            encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"], img_ids=inp["img_ids"])
    with pytest.raises(RuntimeError):
        eng.to("cpu")
