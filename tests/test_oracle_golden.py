"""CPU: the oracle restatement reproduces the committed reference outputs (tests/golden/, made by
oracle/make_golden.py from the real reference) BIT-EXACTLY."""
import os

import pytest
import torch

from oracle import flux_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg(d):
    return fo.FluxConfig(**d)


@torch.no_grad()
def test_tiny_forward_bit_exact(golden):
    g = golden("tiny_forward.pt")
    cfg = _cfg(g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"], dtype=torch.bfloat16)
    inp = g["inputs"]
    regen = fo.synthetic_inputs(cfg, *g["grid"], g["T"], batch=g["batch"], seed0=g["input_seed0"])
    for k in inp:
        assert torch.equal(inp[k], regen[k]), k
    taps = {}
    out = fo.flux_forward(sd, cfg, torch.cat([inp["latents"], inp["cond"]], dim=2), inp["prompt_embeds"],
                          inp["pooled"], g["timestep"], inp["img_ids"], inp["txt_ids"], g["guidance"], taps=taps)
    assert out.dtype == torch.bfloat16 and out.shape == (2, 64, 64)
    assert torch.equal(out, g["sample"])
    for k, v in g["taps"].items():
        assert torch.equal(taps[k], v), k


@torch.no_grad()
def test_tiny_forward_fp32_matches_reference_fp32(golden):
    g = golden("tiny_forward.pt")
    cfg = _cfg(g["config"])
    sd = {k: v.float() for k, v in fo.init_state_dict(cfg, seed=g["weight_seed"]).items()}
    inp = g["inputs"]
    t32 = (g["timestep"].to(torch.bfloat16) * 1000).float() / 1000
    g32 = (g["guidance"].to(torch.bfloat16) * 1000).float() / 1000
    out = fo.flux_forward(sd, cfg, torch.cat([inp["latents"], inp["cond"]], dim=2).float(),
                          inp["prompt_embeds"].float(), inp["pooled"].float(), t32, inp["img_ids"].float(),
                          inp["txt_ids"].float(), g32)
    assert torch.equal(out, g["sample_fp32"])
    # tolerance calibration: reference-bf16 vs reference-fp32 on the tiny model
    rel = ((g["sample"].float() - out).norm() / out.norm()).item()
    assert rel < 2e-2, rel


@torch.no_grad()
def test_tiny_loop_bit_exact(golden):
    g = golden("tiny_loop.pt")
    cfg = _cfg(g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"])
    inp = g["inputs"]
    rec = []
    sig, ts = fo.euler_set_timesteps(g["steps"], 64)
    assert torch.equal(sig, g["sigmas"]) and torch.equal(ts, g["timesteps"])
    last = fo.denoise_loop(sd, cfg, inp["latents"], inp["cond"], inp["prompt_embeds"], inp["pooled"],
                           inp["txt_ids"], inp["img_ids"], g["guidance_scale"], g["steps"], record=rec)
    for i, (v, x) in enumerate(rec):
        assert torch.equal(v, g["noise_preds"][i]), i
        assert torch.equal(x, g["latents"][i]), i
    assert torch.equal(last, g["latents"][-1])


def test_schedules_bit_exact(golden):
    g = golden("schedules.pt")
    for (n, S), ref in g.items():
        assert fo.calculate_shift(S) == ref["mu"]
        sig, ts = fo.euler_set_timesteps(n, S)
        assert torch.equal(sig, ref["sigmas"]), (n, S)
        assert torch.equal(ts, ref["timesteps"]), (n, S)
    # SURVEY §8a13 measured mu values
    assert abs(fo.calculate_shift(2048) - 0.803) < 1e-3
    assert abs(fo.calculate_shift(8192) - 1.843) < 1e-3


@torch.no_grad()
def test_real_dim_blocks_bit_exact(golden):
    g = golden("real_dim_blocks.pt")
    cfg = _cfg(g["config"])
    sd = fo.init_state_dict(cfg, seed=g["weight_seed"])
    inp = fo.synthetic_inputs(cfg, *g["grid"], g["T"], batch=1, seed0=g["input_seed0"])
    taps = {}
    out = fo.flux_forward(sd, cfg, torch.cat([inp["latents"], inp["cond"]], dim=2), inp["prompt_embeds"],
                          inp["pooled"], g["timestep"], inp["img_ids"], inp["txt_ids"], g["guidance"], taps=taps)
    assert torch.equal(out, g["sample"])
    for k, v in g["taps"].items():
        assert torch.equal(taps[k], v), k


def test_bf16_double_rounding_quirk():
    """the model sees bf16(bf16(t)/1000)*1000 -> 984 for t=984.79; guidance 30 -> 29952."""
    t = torch.tensor(984.79)
    seen = (t.to(torch.bfloat16) / 1000).to(torch.bfloat16) * 1000
    assert seen.item() == 984.0
    assert (torch.tensor(30.0).to(torch.bfloat16) * 1000).item() == 29952.0


def test_euler_step_rounding():
    """(sigma_next-sigma) is a 0-dim fp32 tensor: type promotion rounds it to bf16, the product is rounded to
    bf16, then added to the fp32 sample (scheduler :322-330)."""
    g = torch.Generator().manual_seed(0)
    v = torch.randn(4, 64, generator=g).to(torch.bfloat16)
    x = torch.randn(4, 64, generator=g).to(torch.bfloat16)
    s0, s1 = torch.tensor(0.731), torch.tensor(0.702)
    got = fo.euler_step(v, s0, s1, x)
    dt = (s1 - s0).to(torch.bfloat16).float()
    want = (x.float() + (dt * v.float()).to(torch.bfloat16).float()).to(torch.bfloat16)
    assert torch.equal(got, want)


def test_pack_unpack_index_exact():
    """token s=i*(w/2)+j, channel c*4+di*2+dj; unpack inverts pack."""
    B, C, h, w = 1, 16, 6, 10
    lat = torch.arange(B * C * h * w, dtype=torch.float32).view(B, C, h, w)
    p = fo.pack_latents(lat)
    assert p.shape == (B, (h // 2) * (w // 2), C * 4)
    for (i, j, c, di, dj) in [(0, 0, 0, 0, 0), (2, 3, 5, 1, 0), (1, 4, 15, 1, 1)]:
        assert p[0, i * (w // 2) + j, c * 4 + di * 2 + dj] == lat[0, c, 2 * i + di, 2 * j + dj]
    assert torch.equal(fo.unpack_latents(p, h * 8, w * 8), lat)
    ids = fo.prepare_latent_image_ids(3, 5, torch.float32)
    assert ids[1 * 5 + 2].tolist() == [0.0, 1.0, 2.0]
    m = torch.arange(16 * 32, dtype=torch.float32).view(1, 1, 16, 32)
    pm = fo.pack_mask(m)
    assert pm.shape == (1, 1 * 2, 256)
    # channel (py*8+px)*4 + di*2 + dj of token (i,j) is pixel (8*(2i+di)+py, 8*(2j+dj)+px)
    assert pm[0, 1, (3 * 8 + 5) * 4 + 1 * 2 + 0] == m[0, 0, 8 * 1 + 3, 8 * 2 + 5]


def test_conditioning_glue_matches_reference_golden(golden):
    """pack / unpack / mask-pack / VAE (de)normalisation restatements against outputs of the REAL
    FluxFillPipeline helpers (tests/golden/conditioning.pt, written by oracle/make_golden.py::conditioning)."""
    d = golden("conditioning.pt")
    sf, sc, vs = d["shift_factor"], d["scaling_factor"], d["vae_scale_factor"]
    for c in d["cases"]:
        H, W = c["h"] * vs, c["w"] * vs
        mp, lp = fo.prepare_mask_latents(c["mask"], c["masked_image_latents"], c["B"], 16, c["num_images_per_prompt"], H, W,
                                         torch.bfloat16, sf, sc, vs)
        assert torch.equal(mp, c["mask_packed"]) and torch.equal(lp, c["masked_image_latents_packed"])
        assert torch.equal(fo.pack_latents(c["latents"]), c["latents_packed"])
        un = fo.unpack_latents(c["latents_packed"], H, W, vs)
        assert torch.equal(un, c["unpacked"]) and torch.equal(un, c["latents"])
        assert torch.equal(fo.denormalize_vae_latents(un, sf, sc), c["decode_in"])
        assert torch.equal(fo.prepare_latent_image_ids(c["h"] // 2, c["w"] // 2), c["img_ids"])


def test_flop_model_matches_survey():
    for S, tf in [(2048, 37.6650989568), (4608, 84.49265762304), (8192, 165.474467315712)]:
        assert abs(fo.flops_per_step(fo.FLUX_FILL_12B, S, 512) / 1e12 - tf) < 1e-6


# ---------------------------------------------------------------------------------------------- AutoencoderKL oracle
@pytest.mark.parametrize("name", ["vae_small", "vae_flux"])
def test_vae_oracle_matches_reference_golden(name):
    """oracle/vae_oracle.py restates AutoencoderKL encode / decode bit-exactly (fixtures written by the real reference modules)."""
    from oracle import vae_oracle as vo
    d = torch.load(os.path.join(GOLDEN, name + ".pt"))
    cfg = vo.VaeConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in d["config"].items()})
    sd32 = vo.init_state_dict(cfg, seed=d["seed"])
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        sd = {k: v.to(dtype) for k, v in sd32.items()}
        m = vo.encode_moments(sd, cfg, d["image"].to(dtype))
        assert torch.equal(m, d[f"moments_{tag}"])
        assert torch.equal(vo.decode(sd, cfg, d["z"].to(dtype)), d[f"decoded_{tag}"])
        noise = torch.randn(d[f"sample_{tag}"].shape, generator=torch.Generator().manual_seed(5), dtype=dtype)
        assert torch.equal(vo.gaussian_sample(m, noise), d[f"sample_{tag}"])


# ---------------------------------------------------------------------------------------------- prompt-encoder oracle
def test_textenc_t5_oracle_matches_transformers_golden():
    from oracle import textenc_oracle as to
    d = torch.load(os.path.join(GOLDEN, "textenc_t5.pt"))
    cfg = to.T5Cfg(**d["config"])
    sd32 = to.init_state_dict(to.t5_spec(cfg), d["seed"])
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        assert torch.equal(to.t5_encode({k: v.to(dtype) for k, v in sd32.items()}, cfg, d["input_ids"]), d[f"last_hidden_{tag}"])


def test_textenc_clip_oracle_matches_transformers_golden():
    from oracle import textenc_oracle as to
    d = torch.load(os.path.join(GOLDEN, "textenc_clip.pt"))
    cfg = to.ClipCfg(**d["config"])
    sd32 = to.init_state_dict(to.clip_spec(cfg), d["seed"])
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        lh, po = to.clip_encode({k: v.to(dtype) for k, v in sd32.items()}, cfg, d["input_ids"])
        assert torch.equal(lh, d[f"last_hidden_{tag}"]) and torch.equal(po, d[f"pooled_{tag}"])


def test_t5_bucket_table_of_the_engine_equals_the_reference_bucket_function():
    """The host-side table the engine hands to tfx_textenc_encode (textflux_b200.text_encoders.t5_bucket_lut) against
    T5Attention._relative_position_bucket evaluated over the whole [T, T] grid -- index math, bit-exact."""
    from oracle import textenc_oracle as to
    from textflux_b200.text_encoders import t5_bucket_lut
    for T in (1, 7, 48, 77, 512):
        lut = t5_bucket_lut(T, 32, 128)
        ctx = torch.arange(T)[:, None]
        mem = torch.arange(T)[None, :]
        ref = to.t5_relative_position_bucket(mem - ctx, 32, 128)
        assert torch.equal(lut[(mem - ctx) + T - 1].long(), ref)
