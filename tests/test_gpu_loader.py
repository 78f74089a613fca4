"""GPU: an engine built straight from checkpoint files (textflux_b200.loader) is the engine built from the state dict --
same packed weights, so bit-identical outputs; a LoRA file folded at load equals folding by hand."""
import json

import pytest
import torch

from oracle import flux_oracle as fo

pytestmark = pytest.mark.gpu


def _run(eng, inp, t, g):
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    out = eng(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"],
              txt_ids=inp["txt_ids"], img_ids=inp["img_ids"], return_dict=False)[0]
    torch.cuda.synchronize()
    return out


def test_engine_from_files_equals_engine_from_state_dict(tmp_path):
    from textflux_b200 import B200FluxTransformer, fold_lora
    from textflux_b200 import loader as ld
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=31, dtype=torch.bfloat16)
    names = sorted(sd)
    wm = {}
    for i, part in enumerate([names[0::2], names[1::2]]):
        fn = f"diffusion_pytorch_model-{i + 1:05d}-of-00002.safetensors"
        ld.save_safetensors({k: sd[k] for k in part}, str(tmp_path / fn))
        wm.update({k: fn for k in part})
    json.dump({"metadata": {}, "weight_map": wm}, open(tmp_path / ld.SAFE_WEIGHTS_INDEX_NAME, "w"))
    json.dump(dict(cfg.to_dict(), _class_name="FluxTransformer2DModel", _diffusers_version="0.32.0.dev0"), open(tmp_path / ld.CONFIG_NAME, "w"))
    g = torch.Generator().manual_seed(9)
    lora = {}
    for m in ("transformer_blocks.0.attn.to_v", "transformer_blocks.1.attn.add_q_proj", "single_transformer_blocks.1.attn.to_k",
              "single_transformer_blocks.0.proj_mlp", "transformer_blocks.0.ff_context.net.2"):
        o, i = sd[m + ".weight"].shape
        lora[f"transformer.{m}.lora_A.weight"] = (torch.randn(4, i, generator=g) * 0.1).to(torch.bfloat16)
        lora[f"transformer.{m}.lora_B.weight"] = (torch.randn(o, 4, generator=g) * 0.1).to(torch.bfloat16)
    ld.save_safetensors(lora, str(tmp_path / ld.LORA_WEIGHT_NAME_SAFE))

    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 8, 8, 16, batch=2, seed0=500).items()}
    t = (torch.tensor([700.0, 700.0]).to(torch.bfloat16) / 1000).cuda()
    gd = torch.full([2], 30.0).cuda()

    a = _run(ld.load_transformer(str(tmp_path), device="cuda:0"), inp, t, gd)
    b = _run(B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0"), inp, t, gd)
    assert torch.equal(a, b)

    a = _run(ld.load_transformer(str(tmp_path), device="cuda:0", lora=str(tmp_path), lora_scale=0.7), inp, t, gd)
    get = fold_lora(lambda n: sd[n].cuda(), lora, scale=0.7)
    b = _run(B200FluxTransformer(cfg.to_dict(), get, device="cuda:0"), inp, t, gd)
    assert torch.equal(a, b)
    c = _run(B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0"), inp, t, gd)
    assert not torch.equal(a, c)  # the adapter does change the output
    # and the folded engine follows the unfused reference math (W x + scale * B (A x)) within the bf16 tolerance of the path
    sd_l = {k: v.clone() for k, v in sd.items()}
    for k in lora:
        if k.endswith(".lora_A.weight"):
            m = k[len("transformer."):-len(".lora_A.weight")]
            sd_l[m + ".weight"] = (sd[m + ".weight"].float() + 0.7 * (lora[f"transformer.{m}.lora_B.weight"].float() @ lora[k].float())).to(torch.bfloat16)
    with torch.no_grad():
        ref = fo.flux_forward(sd_l, cfg, torch.cat([inp["latents"], inp["cond"]], dim=2).cpu(), inp["prompt_embeds"].cpu(), inp["pooled"].cpu(),
                              t.cpu(), inp["img_ids"].cpu(), inp["txt_ids"].cpu(), gd.cpu())
    rel = ((a.float().cpu() - ref.float()).norm() / ref.float().norm()).item()
    assert rel < 2e-2, rel


def test_lora_hot_swap_on_a_live_engine(tmp_path):
    """load_lora_weights / unload_lora_weights rewrite the packed weights in place under a captured step graph:
    outputs equal a fresh engine built with the same fold, and the base outputs come back bit-exactly."""
    from textflux_b200 import B200FluxTransformer, fold_lora
    cfg = fo.TINY
    sd = {k: v.cuda() for k, v in fo.init_state_dict(cfg, seed=32, dtype=torch.bfloat16).items()}
    g = torch.Generator().manual_seed(10)
    lora = {}
    for m in ("transformer_blocks.1.attn.to_q", "single_transformer_blocks.0.proj_out", "transformer_blocks.0.ff.net.0.proj"):
        o, i = sd[m + ".weight"].shape
        lora[f"transformer.{m}.lora_A.weight"] = (torch.randn(4, i, generator=g) * 0.1).to(torch.bfloat16)
        lora[f"transformer.{m}.lora_B.weight"] = (torch.randn(o, 4, generator=g) * 0.1).to(torch.bfloat16)
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 8, 8, 16, batch=1, seed0=600).items()}
    t = (torch.tensor([500.0]).to(torch.bfloat16) / 1000).cuda()
    gd = torch.full([1], 30.0).cuda()
    eng = B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0")
    base = _run(eng, inp, t, gd).clone()
    base2 = _run(eng, inp, t, gd).clone()  # graph replay
    assert torch.equal(base, base2)
    eng.load_lora_weights(sd.__getitem__, lora, scale=0.9)
    with_lora = _run(eng, inp, t, gd).clone()
    fresh = _run(B200FluxTransformer(cfg.to_dict(), fold_lora(sd.__getitem__, lora, scale=0.9), device="cuda:0"), inp, t, gd)
    assert torch.equal(with_lora, fresh) and not torch.equal(with_lora, base)
    eng.unload_lora_weights(sd.__getitem__)
    assert torch.equal(_run(eng, inp, t, gd), base)


def test_pipeline_directory_loads_every_stage(tmp_path):
    """loader.load_components on a FluxFillPipeline-shaped directory: each engine equals the one built from the state dict."""
    from oracle import textenc_oracle as to
    from oracle import vae_oracle as vo
    from textflux_b200 import B200AutoencoderKL, B200CLIPTextEncoder, B200FluxTransformer, B200T5Encoder
    from textflux_b200 import loader as ld

    def write(sub, sd, config, single, shards=1, index=None):
        d = tmp_path / sub
        d.mkdir()
        json.dump(config, open(d / "config.json", "w"))
        sd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
        if shards == 1:
            ld.save_safetensors(sd, str(d / single))
        else:
            names, wm = sorted(sd), {}
            for i in range(shards):
                fn = single.replace(".safetensors", f"-{i + 1:05d}-of-{shards:05d}.safetensors")
                ld.save_safetensors({k: sd[k] for k in names[i::shards]}, str(d / fn))
                wm.update({k: fn for k in names[i::shards]})
            json.dump({"metadata": {}, "weight_map": wm}, open(d / index, "w"))
        return sd

    cfg = fo.TINY
    sd_t = write("transformer", fo.init_state_dict(cfg, seed=31, dtype=torch.bfloat16), dict(cfg.to_dict(), _class_name="FluxTransformer2DModel"),
                 ld.SAFETENSORS_WEIGHTS_NAME)
    vcfg = vo.SMALL_VAE
    sd_v = write("vae", vo.init_state_dict(vcfg, seed=5), dict(vcfg.reference_kwargs(), _class_name="AutoencoderKL"), ld.SAFETENSORS_WEIGHTS_NAME)
    ccfg, tcfg = to.CLIP_TINY, to.T5_TINY
    sd_c = write("text_encoder", to.init_state_dict(to.clip_spec(ccfg), 11), dict(ccfg.to_dict(), hidden_act="quick_gelu", model_type="clip_text_model",
                                                                                  architectures=["CLIPTextModel"]), ld.TRANSFORMERS_WEIGHTS_NAME)
    sd_5 = write("text_encoder_2", to.init_state_dict(to.t5_spec(tcfg), 12), dict(tcfg.to_dict(), feed_forward_proj="gated-gelu", model_type="t5",
                                                                                  architectures=["T5EncoderModel"]),
                 ld.TRANSFORMERS_WEIGHTS_NAME, shards=2, index=ld.TRANSFORMERS_INDEX_NAME)
    (tmp_path / "scheduler").mkdir()
    json.dump({"_class_name": "FlowMatchEulerDiscreteScheduler", "_diffusers_version": "0.32.0", "base_image_seq_len": 256, "base_shift": 0.5,
               "max_image_seq_len": 4096, "max_shift": 1.15, "num_train_timesteps": 1000, "shift": 3.0, "use_dynamic_shifting": True,
               "some_future_key": 1}, open(tmp_path / "scheduler" / "scheduler_config.json", "w"))

    parts = ld.load_components(str(tmp_path), device="cuda:0")
    assert set(parts) == {"transformer", "vae", "text_encoder", "text_encoder_2", "scheduler"}
    assert isinstance(parts["transformer"], B200FluxTransformer) and isinstance(parts["vae"], B200AutoencoderKL)
    assert isinstance(parts["text_encoder"], B200CLIPTextEncoder) and isinstance(parts["text_encoder_2"], B200T5Encoder)
    assert parts["scheduler"].config["shift"] == 3.0 and parts["scheduler"].config["use_dynamic_shifting"] is True

    cuda = lambda sd: {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(4)
    ids_c = torch.randint(0, ccfg.vocab_size - 1, (2, 77), generator=g)
    ids_c[:, -1] = ccfg.vocab_size - 1
    ref_c = B200CLIPTextEncoder(dict(ccfg.to_dict(), hidden_act="quick_gelu"), cuda(sd_c).__getitem__, device="cuda:0")
    a, b = parts["text_encoder"](ids_c.cuda()), ref_c(ids_c.cuda())
    assert torch.equal(a[0], b[0]) and torch.equal(a.pooler_output, b.pooler_output)
    ids_5 = torch.randint(0, tcfg.vocab_size, (2, 64), generator=g)
    ref_5 = B200T5Encoder(dict(tcfg.to_dict(), feed_forward_proj="gated-gelu"), cuda(sd_5).__getitem__, device="cuda:0")
    assert torch.equal(parts["text_encoder_2"](ids_5.cuda())[0], ref_5(ids_5.cuda())[0])
    ref_v = B200AutoencoderKL.from_state_dict(vcfg.reference_kwargs(), cuda(sd_v), device="cuda:0")
    x = torch.randn(1, 3, 64, 96, generator=g).to(torch.bfloat16).cuda()
    za, zb = parts["vae"].encode(x).latent_dist.mode(), ref_v.encode(x).latent_dist.mode()
    assert torch.equal(za, zb) and torch.equal(parts["vae"].decode(za).sample, ref_v.decode(zb).sample)
    ref_t = B200FluxTransformer.from_state_dict(cfg.to_dict(), sd_t, device="cuda:0")
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 8, 8, 16, batch=1, seed0=500).items()}
    t, gd = (torch.tensor([700.0]).to(torch.bfloat16) / 1000).cuda(), torch.full([1], 30.0).cuda()
    assert torch.equal(_run(parts["transformer"], inp, t, gd), _run(ref_t, inp, t, gd))
    torch.cuda.synchronize()

    from baseline import reference_arm as ra
    if ra.available():  # the reference's own constructor accepts the engines as its registered modules
        ra.import_reference()
        from diffusers import FluxFillPipeline
        pipe = FluxFillPipeline(tokenizer=None, tokenizer_2=None, **parts)
        assert pipe.transformer is parts["transformer"] and pipe.vae is parts["vae"] and pipe.text_encoder_2 is parts["text_encoder_2"]
        assert pipe.vae_scale_factor == 2 ** (len(vcfg.block_out_channels) - 1)
