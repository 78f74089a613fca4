"""Build container only (skipped where /root/reference is absent): oracle vs the live reference modules."""
import pytest
import torch

from oracle import flux_oracle as fo
from oracle.ref_loader import (build_reference_scheduler, build_reference_transformer, import_reference,
                               reference_available)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


@torch.no_grad()
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_forward_bit_exact_vs_live_reference(dtype):
    cfg = fo.FluxConfig(in_channels=384, out_channels=64, num_layers=1, num_single_layers=2, attention_head_dim=64,
                        num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=48,
                        axes_dims_rope=(16, 24, 24))
    sd = fo.init_state_dict(cfg, seed=5, dtype=dtype)
    m = build_reference_transformer(cfg, sd, dtype)
    inp = fo.synthetic_inputs(cfg, 4, 6, 10, batch=3, seed0=11, dtype=dtype)
    t = torch.tensor([0.3, 0.9, 0.55]).to(dtype)
    g = torch.tensor([30.0, 3.5, 1.0])
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    ref = m(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"],
            encoder_hidden_states=inp["prompt_embeds"], txt_ids=inp["txt_ids"], img_ids=inp["img_ids"],
            return_dict=False)[0]
    got = fo.flux_forward(sd, cfg, hs, inp["prompt_embeds"], inp["pooled"], t, inp["img_ids"], inp["txt_ids"], g)
    assert torch.equal(ref, got)


def test_pack_and_ids_vs_live_pipeline():
    import_reference()
    from diffusers.pipelines.flux.pipeline_flux_fill import FluxFillPipeline as P
    lat = torch.randn(2, 16, 8, 12)
    assert torch.equal(P._pack_latents(lat, 2, 16, 8, 12), fo.pack_latents(lat))
    packed = fo.pack_latents(lat)
    assert torch.equal(P._unpack_latents(packed, 64, 96, 8), fo.unpack_latents(packed, 64, 96, 8))
    ids = P._prepare_latent_image_ids(2, 4, 6, "cpu", torch.bfloat16)
    assert torch.equal(ids, fo.prepare_latent_image_ids(4, 6, torch.bfloat16))


def test_scheduler_step_vs_live():
    import numpy as np
    sch = build_reference_scheduler()
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / 5, 5), mu=fo.calculate_shift(1024))
    sig, ts = fo.euler_set_timesteps(5, 1024)
    assert torch.equal(sig, sch.sigmas) and torch.equal(ts, sch.timesteps)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 32, 64, generator=g).to(torch.bfloat16)
    for i, t in enumerate(sch.timesteps):
        v = torch.randn(1, 32, 64, generator=g).to(torch.bfloat16)
        want = sch.step(v, t, x, return_dict=False)[0]
        got = fo.euler_step(v, sig[i], sig[i + 1], x)
        assert torch.equal(want, got)
        x = want
