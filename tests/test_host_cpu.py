"""CPU: host-side logic and the C-ABI library surface (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import flux_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from textflux_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from textflux_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "textflux_b200.h")).read()
    declared = set(re.findall(r"\b(tfx_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_error_convention_without_gpu(lib):
    """Bad arguments return TFX_ERR_INVALID and a message; nothing throws across the boundary."""
    from textflux_b200 import _lib
    assert lib.tfx_create(None, 0, None) == 1
    assert b"null" in lib.tfx_last_error(None)
    assert lib.tfx_euler_step(None, None, None, 4, 0.5, 0.4, None) == 1
    with pytest.raises(ValueError):
        _lib.check(lib.tfx_op_gemv(None, 1, 8, None, None, 8, None, 0, None))
    if not torch.cuda.is_available():
        cfg = _lib.TfxConfig(384, 64, 1, 1, 64, 4, 128, 32, 1, (C.c_int32 * 3)(8, 28, 28))
        h = C.c_void_p()
        assert lib.tfx_create(C.byref(cfg), 0, C.byref(h)) == 3  # TFX_ERR_CUDA: no device, no CPU fallback
        assert h.value is None


def test_engine_refuses_cpu():
    from textflux_b200 import B200FluxTransformer
    cfg = fo.TINY
    with pytest.raises(RuntimeError):
        B200FluxTransformer(cfg.to_dict(), lambda n: None, device="cpu")


def test_scheduler_host_logic_matches_reference_golden(golden):
    from textflux_b200 import B200FlowMatchEulerScheduler, calculate_shift
    g = golden("schedules.pt")
    for (n, S), ref in g.items():
        sch = B200FlowMatchEulerScheduler()
        mu = calculate_shift(S, sch.config.base_image_seq_len, sch.config.max_image_seq_len, sch.config.base_shift,
                             sch.config.max_shift)
        assert mu == ref["mu"]
        sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n, n), mu=mu)
        assert torch.equal(sch.sigmas, ref["sigmas"]) and torch.equal(sch.timesteps, ref["timesteps"])
        assert sch.order == 1 and sch.step_index is None
        assert sch.index_for_timestep(sch.timesteps[min(2, n - 1)]) == min(2, n - 1)
    sch = B200FlowMatchEulerScheduler()
    with pytest.raises(ValueError):
        sch.set_timesteps(num_inference_steps=4)  # dynamic shifting needs mu, like the reference
    with pytest.raises(ValueError):
        sch.step(torch.zeros(1), 3, torch.zeros(1))  # integer timestep rejected, like the reference
    import inspect
    assert "sigmas" in inspect.signature(sch.set_timesteps).parameters  # retrieve_timesteps inspects this (:1305-1314)


def test_packer_is_a_bit_exact_row_regrouping():
    from textflux_b200 import pack_weights, reference_names
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=3)
    assert {n for n, _ in reference_names(cfg)} == set(sd)
    for n, shp in reference_names(cfg):
        assert tuple(sd[n].shape) == shp, n
    P = pack_weights(cfg, sd.__getitem__, "cpu")
    D = cfg.inner_dim
    assert torch.equal(P["d1.qkv_x.w"][D:2 * D], sd["transformer_blocks.1.attn.to_k.weight"])
    assert torch.equal(P["d1.qkv_c.w"][:D], sd["transformer_blocks.1.attn.add_q_proj.weight"])
    assert torch.equal(P["d0.qkv_c.b"][0, 2 * D:], sd["transformer_blocks.0.attn.add_v_proj.bias"])
    assert torch.equal(P["s1.qkvmlp.w"][3 * D:], sd["single_transformer_blocks.1.proj_mlp.weight"])
    assert torch.equal(P["s0.out.w"], sd["single_transformer_blocks.0.proj_out.weight"])
    assert torch.equal(P["d0.rms_k_c"][0], sd["transformer_blocks.0.attn.norm_added_k.weight"])
    L, Ls = cfg.num_layers, cfg.num_single_layers
    assert P["mod.w"].shape == ((12 * L + 3 * Ls + 2) * D, D)
    assert torch.equal(P["mod.w"][(12 * 1 + 6) * D:(12 * 1 + 12) * D], sd["transformer_blocks.1.norm1_context.linear.weight"])
    assert torch.equal(P["mod.w"][(12 * L + 3) * D:(12 * L + 6) * D], sd["single_transformer_blocks.1.norm.linear.weight"])
    assert torch.equal(P["mod.b"][0, (12 * L + 3 * Ls) * D:], sd["norm_out.linear.bias"])
    # every reference value lands exactly once
    assert sum(t.numel() for t in P.values()) == sum(t.numel() for t in sd.values())


def test_lora_fold_matches_unfused_math():
    from textflux_b200 import fold_lora
    cfg = fo.TINY
    sd = fo.init_state_dict(cfg, seed=4)
    D = cfg.inner_dim
    g = torch.Generator().manual_seed(0)
    name = "transformer_blocks.0.attn.to_q"
    lora = {f"transformer.{name}.lora_A.weight": torch.randn(16, D, generator=g) * 0.02,
            f"transformer.{name}.lora_B.weight": torch.randn(D, 16, generator=g) * 0.02,
            f"transformer.{name}.alpha": torch.tensor(8.0)}
    get = fold_lora(sd.__getitem__, lora, scale=1.0)
    W = get(name + ".weight")
    ref = sd[name + ".weight"].float() + 0.5 * lora[f"transformer.{name}.lora_B.weight"] @ lora[f"transformer.{name}.lora_A.weight"]
    assert torch.equal(W, ref.to(torch.bfloat16))
    assert torch.equal(get("transformer_blocks.0.attn.to_k.weight"), sd["transformer_blocks.0.attn.to_k.weight"])
    x = torch.randn(32, D, generator=g).to(torch.bfloat16)
    fused = torch.nn.functional.linear(x.float(), W.float())
    unfused = torch.nn.functional.linear(x.float(), ref)
    assert ((fused - unfused).norm() / unfused.norm()).item() < 4e-3


def test_overshoot_scheduler_host_arithmetic_matches_reference_golden(golden):
    """sigmas (float64 shift) and the per-step scalars (t_o - t, a, b) of the overshoot sampler, bit-exact."""
    from textflux_b200 import B200StochasticRFOvershotScheduler, calculate_shift
    d = golden("overshoot.pt")
    sch = B200StochasticRFOvershotScheduler()
    mu = calculate_shift(d["S"], sch.config.base_image_seq_len, sch.config.max_image_seq_len, sch.config.base_shift,
                         sch.config.max_shift)
    sch.set_timesteps(sigmas=np.linspace(1.0, 1 / d["n"], d["n"]), mu=mu)
    assert torch.equal(sch.sigmas, d["sigmas"]) and torch.equal(sch.timesteps, d["timesteps"])
    for i in range(d["n"]):
        coef, a, b, sg = sch._scalars(i)
        sigma, sn = d["sigmas"][i], d["sigmas"][i + 1]
        t = 1 - sigma
        step = sigma - sn
        tn = min(t + step, 1)
        to_ = min(tn + step * 2.0, 1)
        ra = tn / to_
        rb = ((1 - tn) ** 2 - (ra - tn) ** 2) ** 0.5
        assert coef == float(to_ - t) and a == float(ra) and b == float(rb) and sg == float(sigma)
    with pytest.raises(ValueError):
        sch.set_attn_map(torch.ones(4))


def test_oracle_overshoot_step_bit_exact(golden):
    d = golden("overshoot.pt")
    sig, ts = fo.overshoot_set_timesteps(d["n"], d["S"])
    assert torch.equal(sig, d["sigmas"]) and torch.equal(ts, d["timesteps"])
    g = torch.Generator().manual_seed(d["input_seed"])
    x = torch.randn(2, d["S"], 64, generator=g).to(torch.bfloat16)
    vs = [torch.randn(2, d["S"], 64, generator=g).to(torch.bfloat16) for _ in range(d["n"])]
    gen = torch.Generator().manual_seed(d["noise_seed"])
    for i in range(d["n"]):
        eps = torch.randn(x.shape, generator=gen, dtype=torch.float32)
        x, x1 = fo.overshoot_step(vs[i], sig[i], sig[i + 1], x, eps)
        assert torch.equal(x, d["prev_samples"][i]) and torch.equal(x1, d["predicted_x1"][i])
