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
