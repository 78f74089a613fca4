"""The drop-in itself: the UNMODIFIED reference FluxFillPipeline.__call__ (pipeline_flux_fill.py:1850-2137, installed under
baseline/_ref) run stock on CUDA, then again with `textflux_b200.attach(pipe)` -- same seeds, same inputs, tiny
AutoencoderKL, text_encoder=None with prompt_embeds passed, output_type="latent"  -- for the Euler
scheduler and for TextFlux's default overshoot sampler; plus the host-side semantics around the boundary (modulation
cache, callbacks / interrupt, PEFT-wrapped transformers, LoRA hot-swap bookkeeping)."""
import numpy as np
import pytest
import torch

from baseline import reference_arm as ra
from oracle import flux_oracle as fo

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not ra.available(), reason="reference not installed under baseline/_ref")


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def _cosdist(a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    return 1.0 - torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def _tiny_pipeline(overshoot: bool, vae_channels=(8, 8, 16, 16), vae_groups=4):
    ra.import_reference()
    from diffusers import AutoencoderKL, FlowMatchEulerDiscreteScheduler, FluxFillPipeline, FluxTransformer2DModel
    torch.manual_seed(0)
    vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                        up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=vae_channels, layers_per_block=1,
                        latent_channels=16, norm_num_groups=vae_groups, use_quant_conv=False, use_post_quant_conv=False,
                        shift_factor=0.1159, scaling_factor=0.3611, sample_size=64)
    cfg = fo.TINY
    tr = FluxTransformer2DModel(**cfg.to_dict())
    tr.load_state_dict(fo.init_state_dict(cfg, seed=1234, dtype=torch.float32))
    sch = FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15)
    pipe = FluxFillPipeline(scheduler=sch, vae=vae, text_encoder=None, tokenizer=None, text_encoder_2=None, tokenizer_2=None,
                            transformer=tr)
    pipe = pipe.to("cuda", torch.bfloat16)
    if overshoot:  # run_inference.py:79-91
        from diffusers import StochasticRFOvershotDiscreteScheduler
        pipe.scheduler = StochasticRFOvershotDiscreteScheduler.from_config(pipe.scheduler.config)
        pipe.scheduler.set_c(2.0)
        pipe.scheduler.set_overshot_func(lambda t, dt: t + dt)
    pipe.set_progress_bar_config(disable=True)
    return pipe


def _call(pipe, steps, H=128, W=64, output_type="latent"):
    g = torch.Generator().manual_seed(1)
    image = torch.rand(1, 3, H, W, generator=g)
    mask = torch.zeros(1, 1, H, W)
    mask[:, :, H // 2:] = 1  # TextFlux layout: glyph half kept, scene half to be filled
    pe = torch.randn(1, 16, fo.TINY.joint_attention_dim, generator=g).to(torch.bfloat16).cuda()
    pp = torch.randn(1, fo.TINY.pooled_projection_dim, generator=g).to(torch.bfloat16).cuda()
    seen = []

    def cb(p, i, t, kw):
        seen.append(kw["latents"].clone())
        return {}

    torch.manual_seed(77)  # the overshoot scheduler draws from the global CUDA generator (the pipeline passes none to step)
    torch.cuda.manual_seed(77)
    out = pipe(prompt_embeds=pe, pooled_prompt_embeds=pp, image=image, mask_image=mask, height=H, width=W,
               num_inference_steps=steps, guidance_scale=30.0, generator=torch.Generator().manual_seed(3),
               output_type=output_type, callback_on_step_end=cb, return_dict=False)[0]
    torch.cuda.synchronize()
    return out, seen


@needs_ref
@pytest.mark.parametrize("overshoot", [False, True], ids=["euler", "overshoot"])
def test_unmodified_pipeline_with_engine_attached(overshoot):
    import textflux_b200
    from textflux_b200 import B200FluxTransformer
    steps = 6
    pipe = _tiny_pipeline(overshoot)
    ref_out, ref_seen = _call(pipe, steps)
    stock_type = type(pipe.transformer).__name__
    textflux_b200.attach(pipe)   # zero edits to reference files, library defaults
    assert stock_type == "FluxTransformer2DModel" and isinstance(pipe.transformer, B200FluxTransformer)
    l0 = pipe.transformer.counter("launches")
    out, seen = _call(pipe, steps)
    assert pipe.transformer.counter("launches") > l0  # the engine's kernels ran inside the unmodified __call__
    assert out.shape == ref_out.shape and out.dtype == ref_out.dtype and len(seen) == len(ref_seen) == steps
    for i in range(steps):
        rel, cos = _rel(seen[i], ref_seen[i]), _cosdist(seen[i], ref_seen[i])
        print(f"{'overshoot' if overshoot else 'euler'} step {i}: latents rel-L2 {rel:.3e} cosdist {cos:.2e}")
        assert cos < 1e-4 and rel < 2e-2, (i, rel, cos)   # tiny-model floor: reference bf16-vs-fp32 5.7e-3
    assert _cosdist(out, ref_out) < 1e-4
    # a second image through the same pipeline is served from the modulation cache and gives the same bits
    hits0 = pipe.transformer.counter("mod_cache_hits")
    out2, _ = _call(pipe, steps)
    assert pipe.transformer.counter("mod_cache_hits") - hits0 == steps
    assert torch.equal(out2, out)


@needs_ref
def test_unmodified_pipeline_with_engine_and_vae_attached():
    """The whole GPU side of FluxFillPipeline.__call__ on the engine: vae.encode of the masked image (:1528), the denoising loop,
    vae.decode (:2128) -- a VAE of the DownEncoderBlock2D family at channel counts the engine implements (multiples of 64)."""
    import textflux_b200
    from textflux_b200 import B200AutoencoderKL, B200FluxTransformer
    steps = 4
    pipe = _tiny_pipeline(False, vae_channels=(64, 64, 128, 128), vae_groups=32)
    ref_img, ref_seen = _call(pipe, steps, output_type="pt")
    textflux_b200.attach(pipe, vae=True)
    assert isinstance(pipe.vae, B200AutoencoderKL) and isinstance(pipe.transformer, B200FluxTransformer)
    v0 = pipe.vae.counter("launches")
    img, seen = _call(pipe, steps, output_type="pt")
    assert pipe.vae.counter("launches") > v0
    assert img.shape == ref_img.shape == (1, 3, 128, 64)
    print(f"engine + VAE attached: first-step latents rel-L2 {_rel(seen[0], ref_seen[0]):.3e}, image rel-L2 {_rel(img, ref_img):.3e}, "
          f"cosdist {_cosdist(img, ref_img):.2e}")
    assert _cosdist(seen[0], ref_seen[0]) < 1e-4   # conditioning came through the engine's vae.encode
    assert _cosdist(img, ref_img) < 1e-3 and _rel(img, ref_img) < 5e-2
    # the tiny VAE of the other tests (8..16 channels) is outside what the engine implements: "auto" leaves the reference VAE in place
    pipe2 = _tiny_pipeline(False)
    textflux_b200.attach(pipe2)
    assert type(pipe2.vae).__name__ == "AutoencoderKL"
    with pytest.raises(ValueError):
        textflux_b200.attach(_tiny_pipeline(False), vae=True)


def test_attach_swaps_the_prompt_encoders():
    """attach(pipe, text_encoders=True): transformers' CLIPTextModel / T5EncoderModel on pipe.text_encoder / text_encoder_2 become the
    engine's mirrors and answer the two calls of pipeline_flux_fill.py:1438 / :1483 like the stock modules (the tokenizers need vocabulary
    files that are not in the image, so the pipeline object here is the attribute bag attach() works on, not FluxFillPipeline)."""
    from types import SimpleNamespace
    from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel
    import textflux_b200
    from textflux_b200 import B200CLIPTextEncoder, B200FluxTransformer, B200T5Encoder
    torch.manual_seed(0)
    t5 = T5EncoderModel(T5Config(vocab_size=300, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, feed_forward_proj="gated-gelu",
                                 dropout_rate=0.0, tie_word_embeddings=False, is_encoder_decoder=False, use_cache=False)).to("cuda", torch.bfloat16).eval()
    ccfg = CLIPTextConfig(vocab_size=500, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                          max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)
    ccfg._attn_implementation = "eager"
    clip = CLIPTextModel(ccfg).to("cuda", torch.bfloat16).eval()
    cfg = fo.TINY
    ra_ok = ra.available()
    if ra_ok:
        ra.import_reference()
        from diffusers import FlowMatchEulerDiscreteScheduler, FluxTransformer2DModel
        tr = FluxTransformer2DModel(**cfg.to_dict())
        tr.load_state_dict(fo.init_state_dict(cfg, seed=1234, dtype=torch.float32))
        tr = tr.to("cuda", torch.bfloat16)
        sch = FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15)
    else:
        pytest.skip("reference not installed under baseline/_ref")
    pipe = SimpleNamespace(transformer=tr, scheduler=sch, vae=None, text_encoder=clip, text_encoder_2=t5)
    ids5 = torch.randint(2, 300, (1, 64), device="cuda")
    idsc = torch.randint(3, 498, (1, 77), device="cuda")
    idsc[:, 40:] = 499
    with torch.no_grad():
        ref5 = t5(ids5, output_hidden_states=False)[0]
        refc = clip(idsc, output_hidden_states=False).pooler_output
    textflux_b200.attach(pipe, text_encoders=True)
    assert isinstance(pipe.text_encoder, B200CLIPTextEncoder) and isinstance(pipe.text_encoder_2, B200T5Encoder) and isinstance(pipe.transformer, B200FluxTransformer)
    out5 = pipe.text_encoder_2(ids5, output_hidden_states=False)[0]
    outc = pipe.text_encoder(idsc, output_hidden_states=False).pooler_output
    print(f"attach: T5 rel-L2 {_rel(out5, ref5):.3e} cosdist {_cosdist(out5, ref5):.2e} | CLIP pooled rel-L2 {_rel(outc, refc):.3e} cosdist {_cosdist(outc, refc):.2e}")
    assert out5.shape == ref5.shape and outc.shape == refc.shape and out5.dtype == outc.dtype == torch.bfloat16
    assert _cosdist(out5, ref5) < 1e-3 and _cosdist(outc, refc) < 1e-3


def _engine(cfg, sd, **kw):
    from textflux_b200 import B200FluxTransformer
    return B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0", **kw)


def _fwd(eng, inp, t, g):
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    return eng(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"],
               txt_ids=inp["txt_ids"], img_ids=inp["img_ids"], return_dict=False)[0]


def test_modulation_cache_is_exact_and_bounded():
    """Cached modulation vectors are the bits a fresh computation gives; a slot is only served for identical
    (timestep, guidance, pooled) bits; the cache holds `mod_cache_slots` entries with round-robin replacement."""
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=3)
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 4, 4, 16, batch=2, seed0=9).items()}
    g = torch.tensor([30.0, 3.5], device="cuda")
    ts = [(torch.tensor([a, b]).to(torch.bfloat16) / 1000).cuda() for a, b in ((900.0, 900.0), (700.0, 512.0), (100.0, 33.0))]
    cold = _engine(cfg, sd)
    cold.set_option("mod_cache_slots", 0)
    want = [_fwd(cold, inp, t, g) for t in ts]
    eng = _engine(cfg, sd)
    eng.set_option("mod_cache_slots", 2)
    got = [_fwd(eng, inp, t, g) for t in ts]                  # 3 misses into 2 slots: ts[0] is evicted
    assert eng.counter("mod_cache_hits") == 0 and eng.counter("mod_cache_valid") == 2
    got += [_fwd(eng, inp, ts[2], g), _fwd(eng, inp, ts[1], g)]   # 2 hits
    assert eng.counter("mod_cache_hits") == 2
    got.append(_fwd(eng, inp, ts[0], g))                          # evicted -> recomputed
    assert eng.counter("mod_cache_hits") == 2
    for a, b in zip(got, want + [want[2], want[1], want[0]]):
        assert torch.equal(a, b)
    # any differing bit of the key is a miss: guidance, one pooled element
    h = eng.counter("mod_cache_hits")
    _fwd(eng, inp, ts[0], torch.tensor([30.0, 3.75], device="cuda"))
    inp2 = dict(inp, pooled=inp["pooled"].clone())
    inp2["pooled"][1, 5] += 0.5
    out_p = _fwd(eng, inp2, ts[0], g)
    assert eng.counter("mod_cache_hits") == h
    assert torch.equal(out_p, _fwd(cold, inp2, ts[0], g))
    eng.set_option("mod_cache_reset", 1)
    assert eng.counter("mod_cache_valid") == 0


def test_denoise_callback_replaces_tensors_and_interrupt():
    """pipeline_flux_fill.py:2078-2079 (interrupt) and :2105-2112 (callback_on_step_end may replace latents / prompt_embeds)."""
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=4)
    eng = _engine(cfg, sd)
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 4, 4, 16, batch=1, seed0=21).items()}
    args = (inp["latents"], inp["cond"], inp["prompt_embeds"], inp["pooled"], inp["txt_ids"], inp["img_ids"], 30.0, 5)
    plain = eng.denoise(*args)
    steps_seen = []
    new_pe = (inp["prompt_embeds"].float() * 0.5).to(torch.bfloat16)

    def cb(engine, i, t, kw):
        steps_seen.append((i, float(t)))
        assert set(kw) == {"latents", "prompt_embeds"}
        if i == 1:
            return {"latents": kw["latents"] * 0 + 1, "prompt_embeds": new_pe}
        return {}

    swapped = eng.denoise(*args, callback_on_step_end=cb, callback_on_step_end_tensor_inputs=("latents", "prompt_embeds"))
    assert [i for i, _ in steps_seen] == [0, 1, 2, 3, 4] and not torch.equal(swapped, plain)
    # the same thing by hand: 2 steps, replace, 3 more steps with the new prompt embeddings
    from textflux_b200 import B200FlowMatchEulerScheduler, calculate_shift
    sch = B200FlowMatchEulerScheduler()
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / 5, 5), device="cuda", mu=calculate_shift(16, 256, 4096, 0.5, 1.15))
    ts = (sch.timesteps[:, None].to(torch.bfloat16) / 1000).contiguous()
    g = torch.full([1], 30.0, device="cuda")
    x, pe = inp["latents"], inp["prompt_embeds"]
    for i in range(5):
        x = eng.step(x, inp["cond"], pe, inp["pooled"], ts[i], g, inp["img_ids"], inp["txt_ids"], sch.sigmas_cpu[i], sch.sigmas_cpu[i + 1])
        if i == 1:
            x, pe = x * 0 + 1, new_pe
    assert torch.equal(swapped, x)

    def stop(engine, i, t, kw):
        if i == 1:
            engine.interrupt = True
        return {}

    seen, plain_seen = [], []
    halted = eng.denoise(*args, callback_on_step_end=stop, callback=lambda i, x: seen.append(i))
    assert seen == [0, 1]
    assert torch.equal(eng.denoise(*args, callback=lambda i, x: plain_seen.append(x.clone())), plain)
    assert torch.equal(halted, plain_seen[1]) and not torch.equal(halted, plain)
    with pytest.raises(ValueError):
        eng.denoise(*args, callback_on_step_end=cb, callback_on_step_end_tensor_inputs=("noise_pred",))


class _FakeLoraLinear(torch.nn.Module):
    """Same attribute surface as peft.tuners.lora.layer.Linear (peft is not installed here): base_layer, lora_A / lora_B
    ModuleDicts keyed by adapter name, scaling dict, active_adapters; forward = base(x) + scaling * B(A(x))."""

    def __init__(self, base: torch.nn.Linear, r: int, scaling: float, seed: int):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.base_layer = base
        self.lora_A = torch.nn.ModuleDict({"default": torch.nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = torch.nn.ModuleDict({"default": torch.nn.Linear(r, base.out_features, bias=False)})
        with torch.no_grad():
            self.lora_A["default"].weight.copy_(torch.randn(r, base.in_features, generator=g) * 0.05)
            self.lora_B["default"].weight.copy_(torch.randn(base.out_features, r, generator=g) * 0.05)
        self.scaling = {"default": scaling}
        self.active_adapters = ["default"]
        self.merged = False

    def forward(self, x):
        return self.base_layer(x) + self.scaling["default"] * self.lora_B["default"](self.lora_A["default"](x))


@needs_ref
def test_attach_folds_peft_wrapped_transformer():
    """run_inference_lora.py:52-65 leaves PEFT layers inside pipe.transformer (state-dict keys ...base_layer.weight /
    ...lora_A.default.weight): from_reference folds them, and the result tracks the reference's UNFUSED forward."""
    from textflux_b200 import B200FluxTransformer
    ra.import_reference()
    from diffusers import FluxTransformer2DModel
    cfg = fo.TINY
    tr = FluxTransformer2DModel(**cfg.to_dict())
    tr.load_state_dict(fo.init_state_dict(cfg, seed=1234, dtype=torch.float32))
    targets = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "attn.add_q_proj", "ff.net.2", "ff_context.net.0.proj")
    n = 0
    for name, mod in list(tr.named_modules()):
        if isinstance(mod, torch.nn.Linear) and name.endswith(targets) and ".lora_" not in name:
            parent = tr.get_submodule(name.rpartition(".")[0])
            setattr(parent, name.rpartition(".")[2], _FakeLoraLinear(mod, 4, 2.0, seed=100 + n))
            n += 1
    assert n > 10
    tr = tr.to("cuda", torch.bfloat16).eval()
    assert any(".base_layer.weight" in k for k in tr.state_dict())
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 8, 8, 16, batch=1, seed0=5).items()}
    t = (torch.tensor([612.5]).to(torch.bfloat16) / 1000).cuda()
    g = torch.full([1], 30.0, device="cuda")
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    with torch.no_grad():
        ref = tr(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"],
                 txt_ids=inp["txt_ids"], img_ids=inp["img_ids"], return_dict=False)[0]
        base_sd = {k.replace(".base_layer.", "."): v for k, v in tr.state_dict().items() if ".lora_" not in k}
    eng = B200FluxTransformer.from_reference(tr, device="cuda:0")
    out = _fwd(eng, inp, t, g)
    no_lora = _fwd(_engine(cfg, base_sd), inp, t, g)
    rel, rel_base = _rel(out, ref), _rel(no_lora, ref)
    print(f"PEFT-wrapped tiny model: folded engine vs unfused reference rel-L2 {rel:.3e} (engine without the adapter: {rel_base:.3e})")
    assert rel < 1.2e-2 and _cosdist(out, ref) < 1e-4 and rel_base > 1.5 * rel  # measured 4.3e-3 with / 9.4e-3 without the adapter
    assert len(eng._lora_modules) == n   # recorded, so unload / swap restore them
    eng.unload_lora_weights(base_sd.__getitem__)
    assert torch.equal(_fwd(eng, inp, t, g), no_lora)
