"""Generate tests/golden/*.pt by running the REAL reference (read-only import) in the build container.

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The fixtures pin oracle/flux_oracle.py (and through it the CUDA
engine) to outputs of the reference itself; the reference tree cannot travel to the GPU box,
the fixtures do.  Weights are not stored: they are regenerated from oracle.init_state_dict
(seeded torch CPU generators, identical on every box with the same torch build).
"""
from __future__ import annotations

import os

import torch

from . import flux_oracle as fo
from .ref_loader import build_reference_scheduler, build_reference_transformer

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _hook_taps(model):
    taps = {}
    hs = []
    for i, blk in enumerate(model.transformer_blocks):
        def f(_m, _in, out, i=i):
            taps[f"double.{i}.enc"], taps[f"double.{i}.x"] = out[0].clone(), out[1].clone()
        hs.append(blk.register_forward_hook(f))
    for i, blk in enumerate(model.single_transformer_blocks):
        def f(_m, _in, out, i=i):
            taps[f"single.{i}"] = out.clone()
        hs.append(blk.register_forward_hook(f))

    def f_temb(_m, _in, out):
        taps["temb"] = out.clone()

    def f_rope(_m, _in, out):
        taps["rope_cos"], taps["rope_sin"] = out[0].clone(), out[1].clone()

    hs.append(model.time_text_embed.register_forward_hook(f_temb))
    hs.append(model.pos_embed.register_forward_hook(f_rope))
    return taps, hs


@torch.no_grad()
def tiny_forward():
    """BASELINE.json configs[0]: tiny 2+2 model, 8x8 packed tokens (16x16 latent), 16 text tokens, B=2."""
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=1234, dtype=torch.bfloat16)
    m = build_reference_transformer(cfg, sd, torch.bfloat16)
    inp = fo.synthetic_inputs(cfg, 8, 8, 16, batch=2, seed0=1000)
    t = torch.tensor([612.5, 612.5]).to(torch.bfloat16) / 1000  # what pipeline :2082,2086 hands over
    g = torch.full([1], 30.0, dtype=torch.float32).expand(2)
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    taps, hooks = _hook_taps(m)
    out = m(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"],
            encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"], img_ids=inp["img_ids"],
            return_dict=False)[0]
    for h in hooks:
        h.remove()
    # fp32 reference on the same bf16-rounded weights/inputs, fed the t*1000 / g*1000 the bf16 model sees
    m32 = build_reference_transformer(cfg, sd, torch.float32)
    t32 = (t.to(torch.bfloat16) * 1000).float() / 1000
    g32 = (g.to(torch.bfloat16) * 1000).float() / 1000
    out32 = m32(hidden_states=hs.float(), timestep=t32, guidance=g32, pooled_projections=inp["pooled"].float(),
                encoder_hidden_states=inp["prompt_embeds"].float(), txt_ids=inp["txt_ids"].float(),
                img_ids=inp["img_ids"].float(), return_dict=False)[0]
    return dict(config=cfg.to_dict(), weight_seed=1234, input_seed0=1000, grid=(8, 8), T=16, batch=2,
                timestep=t, guidance=g, inputs=inp, sample=out, taps=taps, sample_fp32=out32)


@torch.no_grad()
def tiny_loop(steps: int = 4):
    """4 Euler steps of the reference transformer + reference scheduler, driven as pipeline :2077-2098 does."""
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=1234, dtype=torch.bfloat16)
    m = build_reference_transformer(cfg, sd, torch.bfloat16)
    sch = build_reference_scheduler()
    inp = fo.synthetic_inputs(cfg, 8, 8, 16, batch=1, seed0=2000)
    S = 64
    import numpy as np
    sigmas = np.linspace(1.0, 1 / steps, steps)
    mu = fo.calculate_shift(S, sch.config.base_image_seq_len, sch.config.max_image_seq_len,
                            sch.config.base_shift, sch.config.max_shift)
    sch.set_timesteps(sigmas=sigmas, mu=mu)
    latents = inp["latents"]
    guidance = torch.full([1], 30.0, dtype=torch.float32).expand(1)
    preds, lats = [], []
    for t in sch.timesteps:
        timestep = t.expand(1).to(latents.dtype)
        v = m(hidden_states=torch.cat((latents, inp["cond"]), dim=2), timestep=timestep / 1000, guidance=guidance,
              pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"],
              txt_ids=inp["txt_ids"], img_ids=inp["img_ids"], return_dict=False)[0]
        latents = sch.step(v, t, latents, return_dict=False)[0]
        preds.append(v)
        lats.append(latents)
    return dict(config=cfg.to_dict(), weight_seed=1234, input_seed0=2000, grid=(8, 8), T=16, steps=steps,
                guidance_scale=30.0, sigmas=sch.sigmas.clone(), timesteps=sch.timesteps.clone(),
                inputs=inp, noise_preds=preds, latents=lats)


@torch.no_grad()
def schedules():
    """set_timesteps outputs of the real scheduler for the BASELINE configs' (n, S)."""
    import numpy as np
    out = {}
    for n, S in [(30, 2048), (30, 4608), (30, 4736), (30, 4096), (50, 8192), (50, 12288), (1, 64), (4, 64)]:
        sch = build_reference_scheduler()
        mu = fo.calculate_shift(S, sch.config.base_image_seq_len, sch.config.max_image_seq_len,
                                sch.config.base_shift, sch.config.max_shift)
        sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n, n), mu=mu)
        out[(n, S)] = dict(mu=mu, sigmas=sch.sigmas.clone(), timesteps=sch.timesteps.clone())
    return out


@torch.no_grad()
def real_dim_blocks():
    """One double + one single block at the real FLUX dims (D=3072, 24x128, rope 16/56/56), short sequences."""
    cfg = fo.FluxConfig(num_layers=1, num_single_layers=1)
    sd = fo.init_state_dict(cfg, seed=77, dtype=torch.bfloat16)
    m = build_reference_transformer(cfg, sd, torch.bfloat16)
    inp = fo.synthetic_inputs(cfg, 8, 16, 64, batch=1, seed0=3000)
    t = torch.tensor([984.79]).to(torch.bfloat16) / 1000
    g = torch.full([1], 30.0, dtype=torch.float32)
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    taps, hooks = _hook_taps(m)
    out = m(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"],
            encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"], img_ids=inp["img_ids"],
            return_dict=False)[0]
    for h in hooks:
        h.remove()
    keep = {k: v for k, v in taps.items() if k in ("temb", "double.0.x", "double.0.enc", "single.0")}
    return dict(config=cfg.to_dict(), weight_seed=77, input_seed0=3000, grid=(8, 16), T=64, timestep=t, guidance=g,
                sample=out, taps=keep)


@torch.no_grad()
def overshoot_steps():
    """The TextFlux default sampler (demo.py:15, scripts/batch_eval.sh): 6 steps of the real
    StochasticRFOvershotDiscreteScheduler on seeded bf16 tensors, c = 2.0, overshot_func(t, dt) = t + dt."""
    import numpy as np
    from .ref_loader import import_reference
    d = import_reference()
    from diffusers.schedulers.scheduling_stochastic_rf_discrete_overshot import StochasticRFOvershotDiscreteScheduler
    sch = StochasticRFOvershotDiscreteScheduler(use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15)
    sch.set_c(2.0)
    sch.set_overshot_func(lambda t, dt: t + dt)
    n, S = 6, 96
    mu = fo.calculate_shift(S)
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n, n), mu=mu)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(2, S, 64, generator=g).to(torch.bfloat16)
    vs = [torch.randn(2, S, 64, generator=g).to(torch.bfloat16) for _ in range(n)]
    gen = torch.Generator().manual_seed(777)
    prevs, x1s = [], []
    for i, t in enumerate(sch.timesteps):
        x, x1 = sch.step(vs[i], t, x, generator=gen, return_dict=False)
        prevs.append(x)
        x1s.append(x1)
    return dict(n=n, S=S, c=2.0, sigmas=sch.sigmas.clone(), timesteps=sch.timesteps.clone(), input_seed=123, noise_seed=777,
                prev_samples=prevs, predicted_x1=x1s)


@torch.no_grad()
def conditioning():
    """the REAL FluxFillPipeline._pack_latents / _unpack_latents / _prepare_latent_image_ids /
    prepare_mask_latents (latent-input branch, no VAE) and the :2127 de-normalisation on seeded tensors."""
    from types import SimpleNamespace
    from .ref_loader import import_reference
    d = import_reference()
    P = d.FluxFillPipeline
    shift, scale, vs = 0.1159, 0.3611, 8  # FLUX VAE config (diffusers/scripts/convert_flux_to_diffusers.py:299-300)
    out = dict(shift_factor=shift, scaling_factor=scale, vae_scale_factor=vs, cases=[])
    for ci, (B, h, w, n_img, mask_dtype) in enumerate([(2, 8, 12, 1, torch.float32), (1, 16, 8, 2, torch.bfloat16), (1, 6, 10, 1, torch.float32)]):
        g = torch.Generator().manual_seed(4000 + ci)
        lat = torch.randn(B, 16, h, w, generator=g).to(torch.bfloat16)
        mil = (torch.randn(B, 16, h, w, generator=g) * 3).to(torch.bfloat16)
        mask = (torch.rand(B, 1, h * vs, w * vs, generator=g) > 0.5).to(mask_dtype)
        me = SimpleNamespace(vae_scale_factor=vs, vae=SimpleNamespace(config=SimpleNamespace(shift_factor=shift, scaling_factor=scale)),
                             _pack_latents=P._pack_latents)
        mask_p, mil_p = P.prepare_mask_latents(me, mask, mil, B, 16, n_img, h * vs, w * vs, torch.bfloat16, "cpu", None)
        packed = P._pack_latents(lat, B, 16, h, w)
        unpacked = P._unpack_latents(packed, h * vs, w * vs, vs)
        decode_in = (unpacked / scale) + shift
        ids = P._prepare_latent_image_ids(B, h // 2, w // 2, "cpu", torch.bfloat16)
        out["cases"].append(dict(B=B, h=h, w=w, num_images_per_prompt=n_img, latents=lat, masked_image_latents=mil, mask=mask,
                                 mask_packed=mask_p.clone(), masked_image_latents_packed=mil_p.clone(), latents_packed=packed.clone(),
                                 unpacked=unpacked.clone(), decode_in=decode_in.clone(), img_ids=ids.clone()))
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.save(tiny_forward(), os.path.join(GOLDEN, "tiny_forward.pt"))
    torch.save(tiny_loop(), os.path.join(GOLDEN, "tiny_loop.pt"))
    torch.save(schedules(), os.path.join(GOLDEN, "schedules.pt"))
    torch.save(real_dim_blocks(), os.path.join(GOLDEN, "real_dim_blocks.pt"))
    torch.save(overshoot_steps(), os.path.join(GOLDEN, "overshoot.pt"))
    torch.save(conditioning(), os.path.join(GOLDEN, "conditioning.pt"))
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
