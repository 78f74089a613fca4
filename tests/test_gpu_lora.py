"""LoRA at the real dimensions of BASELINE config 4 (TextFlux-LoRA-beta: rank 16 on the 12 target module suffixes of
scripts/train_lora.py:511-524, alpha/r = 1), against the math the reference actually runs at inference -- adapters kept
UNFUSED by PEFT, `W x + (alpha/r) B (A x)` in eager bf16 (run_inference_lora.py:52-65) -- evaluated on CUDA by the oracle
port with its Linear swapped for the unfused form."""
import contextlib

import pytest
import torch

from oracle import flux_oracle as fo

pytestmark = pytest.mark.gpu

TARGETS = ("attn.to_k", "attn.to_q", "attn.to_v", "attn.to_out.0", "attn.add_k_proj", "attn.add_q_proj", "attn.add_v_proj",
           "attn.to_add_out", "ff.net.0.proj", "ff.net.2", "ff_context.net.0.proj", "ff_context.net.2")


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


def _cosdist(a, b):
    return 1.0 - torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def make_lora(sd, rank, seed, b_std=0.02, device="cuda"):
    """A [r, in] ~ N(0, 0.02^2), B [out, r] ~ N(0, b_std^2) on every module whose name ends with a target."""
    g = torch.Generator(device=device).manual_seed(seed)
    lora = {}
    for k in sd:
        if k.endswith(".weight") and k[: -len(".weight")].endswith(TARGETS):
            m = k[: -len(".weight")]
            o, i = sd[k].shape
            lora[f"transformer.{m}.lora_A.weight"] = (torch.randn(rank, i, generator=g, device=device) * 0.02).to(torch.bfloat16)
            lora[f"transformer.{m}.lora_B.weight"] = (torch.randn(o, rank, generator=g, device=device) * b_std).to(torch.bfloat16)
    return lora


@contextlib.contextmanager
def unfused_lora(lora, scale=1.0):
    """Swap the oracle's Linear for PEFT's forward: base(x) + scaling * lora_B(lora_A(x)) (peft/tuners/lora/layer.py Linear.forward)."""
    orig = fo._linear

    def lin(sd, name, x):
        y = orig(sd, name, x)
        ka = f"transformer.{name}.lora_A.weight"
        if ka in lora:
            A = lora[ka].to(x.dtype)
            Bm = lora[f"transformer.{name}.lora_B.weight"].to(x.dtype)
            y = y + torch.nn.functional.linear(torch.nn.functional.linear(x, A), Bm) * scale
        return y

    fo._linear = lin
    try:
        yield
    finally:
        fo._linear = orig


def _setup(S_grid=(64, 64), T=512, seed=77):
    cfg = fo.FluxConfig(num_layers=1, num_single_layers=1)  # D = 3072, 24 heads x 128: one double + one single block
    sd = {k: v.cuda() for k, v in fo.init_state_dict(cfg, seed=seed, dtype=torch.bfloat16).items()}
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, *S_grid, T, batch=1, seed0=4100).items()}
    t = (torch.tensor([702.0]).to(torch.bfloat16) / 1000).cuda()
    g = torch.full([1], 30.0).cuda()
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    return cfg, sd, inp, t, g, hs


def _oracle(sd, cfg, hs, inp, t, g, fp32=False):
    with torch.no_grad():
        if fp32:
            sd = {k: v.float() for k, v in sd.items()}
            t32 = (t.to(torch.bfloat16) * 1000).float() / 1000
            g32 = (g.to(torch.bfloat16) * 1000).float() / 1000
            return fo.flux_forward(sd, cfg, hs.float(), inp["prompt_embeds"].float(), inp["pooled"].float(), t32,
                                   inp["img_ids"].float(), inp["txt_ids"].float(), g32)
        return fo.flux_forward(sd, cfg, hs, inp["prompt_embeds"], inp["pooled"], t, inp["img_ids"], inp["txt_ids"], g)


def _engine_fwd(eng, hs, inp, t, g):
    return eng(hidden_states=hs, timestep=t, guidance=g, pooled_projections=inp["pooled"], encoder_hidden_states=inp["prompt_embeds"],
               txt_ids=inp["txt_ids"], img_ids=inp["img_ids"], return_dict=False)[0]


def test_config4_rank16_fold_vs_unfused_reference_math():
    """1024x1024 (S = 4096, T = 512), D = 3072, rank 16 on all 12 targets of the double block and q/k/v of the single block."""
    from textflux_b200 import B200FluxTransformer, fold_lora
    cfg, sd, inp, t, g, hs = _setup()
    lora = make_lora(sd, 16, seed=5)
    assert len(lora) == 2 * (12 + 3)
    with unfused_lora(lora):
        ref16 = _oracle(sd, cfg, hs, inp, t, g)
        ref32 = _oracle(sd, cfg, hs, inp, t, g, fp32=True)
    base16 = _oracle(sd, cfg, hs, inp, t, g)
    eng = B200FluxTransformer(cfg.to_dict(), fold_lora(sd.__getitem__, lora), device="cuda:0")
    out = _engine_fwd(eng, hs, inp, t, g)
    torch.cuda.synchronize()
    floor = _rel(ref16, ref32)
    e16, e32, effect = _rel(out, ref16), _rel(out, ref32), _rel(base16, ref16)
    print(f"config-4 LoRA r16 at D=3072, N=4608: unfused bf16-vs-fp32 floor {floor:.3e} | folded engine vs unfused bf16 {e16:.3e}, vs fp32 {e32:.3e} | "
          f"adapter effect (no adapter vs unfused) {effect:.3e} | cosdist {_cosdist(out, ref16):.2e}")
    assert e16 <= 2.0 * floor and e32 <= 1.5 * floor, (e16, e32, floor)
    assert _cosdist(out, ref16) < 1e-4
    assert effect > 5 * e32  # the adapter is resolved, not rounded away


def test_swap_after_fold_at_construction_restores_first_adapter_modules():
    """an engine built with adapter A folded at construction (fold_lora getter / load_transformer(lora=...)), then
    switched to adapter B, must equal a fresh engine built with B only -- A's modules that B does not touch go back to base."""
    from textflux_b200 import B200FluxTransformer, fold_lora
    cfg = fo.TINY
    sd = {k: v.cuda() for k, v in fo.init_state_dict(cfg, seed=41, dtype=torch.bfloat16).items()}
    g = torch.Generator().manual_seed(3)

    def adapter(mods, s):
        d = {}
        for m in mods:
            o, i = sd[m + ".weight"].shape
            d[f"transformer.{m}.lora_A.weight"] = (torch.randn(4, i, generator=g) * s).to(torch.bfloat16)
            d[f"transformer.{m}.lora_B.weight"] = (torch.randn(o, 4, generator=g) * s).to(torch.bfloat16)
        return d

    lora_a = adapter(("transformer_blocks.0.attn.to_q", "transformer_blocks.1.ff.net.2", "single_transformer_blocks.1.attn.to_v"), 0.1)
    lora_b = adapter(("transformer_blocks.0.attn.to_q", "single_transformer_blocks.0.proj_mlp"), 0.12)
    inp = {k: v.cuda() for k, v in fo.synthetic_inputs(cfg, 8, 8, 16, batch=1, seed0=70).items()}
    t = (torch.tensor([400.0]).to(torch.bfloat16) / 1000).cuda()
    gd = torch.full([1], 30.0).cuda()
    hs = torch.cat([inp["latents"], inp["cond"]], dim=2)
    eng = B200FluxTransformer(cfg.to_dict(), fold_lora(sd.__getitem__, lora_a), device="cuda:0")
    assert len(eng._lora_modules) == 3
    with_a = _engine_fwd(eng, hs, inp, t, gd).clone()
    eng.load_lora_weights(sd.__getitem__, lora_b)
    with_b = _engine_fwd(eng, hs, inp, t, gd).clone()
    fresh_b = _engine_fwd(B200FluxTransformer(cfg.to_dict(), fold_lora(sd.__getitem__, lora_b), device="cuda:0"), hs, inp, t, gd)
    assert torch.equal(with_b, fresh_b) and not torch.equal(with_b, with_a)
    eng.unload_lora_weights(sd.__getitem__)
    base = _engine_fwd(B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0"), hs, inp, t, gd)
    assert torch.equal(_engine_fwd(eng, hs, inp, t, gd), base)


# ---------------------------------------------------------------------------------------------- the unfused side path
def _lib_and_chk():
    from textflux_b200 import _lib
    return _lib.load(), _lib.check


def _linear_lora(A, W, bias, la, lb, mode=0, cta_group=2, m_band=0, gate=None, res=None):
    lib, chk = _lib_and_chk()
    M, K = A.shape
    N = W.shape[0]
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    if res is not None:
        out.copy_(res)
    t = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
    chk(lib.tfx_op_linear_lora(A.data_ptr(), A.stride(0), W.data_ptr(), bias.data_ptr(), None if la is None else la.data_ptr(),
                               None if lb is None else lb.data_ptr(), t.data_ptr(), out.data_ptr(), N, M, N, K, mode,
                               None if gate is None else gate.data_ptr(), None if res is None else out.data_ptr(), cta_group, m_band,
                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out, t


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("M,N,K,rank", [(300, 320, 192, 16), (2560, 3072, 3072, 48), (4608, 9216, 3072, 64), (128, 256, 64, 4)])
def test_linear_with_side_adapter_matches_unfused_math(M, N, K, rank, cta_group):
    """Y = bf16(x W^T + bf16(x A^T) B^T + bias): the T GEMM is bit-exact against torch's bf16 rounding of the fp32 product up to
    accumulation order, the main GEMM carries the extension k-block in the same fp32 accumulator."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + rank)
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    la = torch.zeros(64, K, device="cuda", dtype=torch.bfloat16)
    lb = torch.zeros(N, 64, device="cuda", dtype=torch.bfloat16)
    la[:rank] = (torch.randn(rank, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    lb[:, :rank] = (torch.randn(N, rank, generator=g, device="cuda") * 0.3).to(torch.bfloat16)
    out, t = _linear_lora(x, W, b, la, lb, cta_group=cta_group)
    t_ref = x.float() @ la.float().T
    assert _rel(t, t_ref) < 4e-3 and torch.equal(t[:, rank:], torch.zeros_like(t[:, rank:]))
    ref = x.float() @ W.float().T + t.float() @ lb.float().T + b.float()
    base = x.float() @ W.float().T + b.float()
    assert _rel(out, ref) < 4e-3, _rel(out, ref)
    assert _rel(out, base) > 10 * _rel(out, ref)  # the adapter is in the result
    # no adapter through the same entry point == the plain GEMM, bit for bit
    plain, _ = _linear_lora(x, W, b, None, None, cta_group=cta_group)
    assert _rel(plain, base) < 4e-3


@pytest.mark.parametrize("band", [1, 3, 5, 40, -1, -3, -7, -40])
@pytest.mark.parametrize("M,N,K", [(2560, 3072, 1024), (1300, 640, 256), (5120, 3072, 512)])
def test_banded_tile_order_is_bit_identical(M, N, K, band):
    """GemmParams::m_band only permutes which CTA pair computes which tile."""
    g = torch.Generator(device="cuda").manual_seed(M + band)
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    ref, _ = _linear_lora(x, W, b, None, None, m_band=0)
    out, _ = _linear_lora(x, W, b, None, None, m_band=band)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("band", [1, 5, -1])
@pytest.mark.parametrize("M,N,K,rank", [(5120, 3072, 4096, 0), (1300, 640, 1000, 0), (2560, 3072, 2048, 16)])
def test_alternating_k_direction(M, N, K, rank, band):
    """GemmParams::k_snake: the tiles of every second band accumulate their k-blocks back to front (the extension block of the LoRA
    side path included) -- the same sum in another order, so equal to the forward walk up to fp32 accumulation order, deterministic,
    as close to the fp32 product as the forward walk, and bit-identical to the forward walk inside the even bands."""
    g = torch.Generator(device="cuda").manual_seed(M + K + band)
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    la = lb = None
    if rank:
        la = (torch.randn(64, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
        la[rank:] = 0
        lb = (torch.randn(N, 64, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    fwd, _ = _linear_lora(x, W, b, la, lb, m_band=band)
    snake = band + (1000 if band > 0 else -1000)
    out, _ = _linear_lora(x, W, b, la, lb, m_band=snake)
    again, _ = _linear_lora(x, W, b, la, lb, m_band=snake)
    assert torch.equal(out, again)
    assert not torch.equal(out, fwd)
    if band > 0:  # rows of band 0 (M tiles 0 .. band - 1) sum forwards either way
        assert torch.equal(out[: 256 * band], fwd[: 256 * band])
    else:
        assert torch.equal(out[:, : 192 * -band], fwd[:, : 192 * -band])  # tiles are at least 192 columns wide
    ref = x.float() @ W.float().t() + b.float()
    if rank:
        ref = ref + (x.float() @ la.float().t()).to(torch.bfloat16).float() @ lb.float().t()
    assert abs(_rel(out, ref) - _rel(fwd, ref)) < 2e-4 and _rel(out, ref) < 4e-3 and _rel(out, fwd) < 1e-3


def test_config4_rank16_side_path_vs_unfused_reference_math():
    """Same comparison as the fold test above, through the side path: the engine computes the reference's unfused form."""
    from textflux_b200 import B200FluxTransformer
    cfg, sd, inp, t, g, hs = _setup()
    lora = make_lora(sd, 16, seed=5)
    with unfused_lora(lora):
        ref16 = _oracle(sd, cfg, hs, inp, t, g)
        ref32 = _oracle(sd, cfg, hs, inp, t, g, fp32=True)
    base16 = _oracle(sd, cfg, hs, inp, t, g)
    eng = B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0")
    base_out = _engine_fwd(eng, hs, inp, t, g).clone()
    plan = eng.load_lora_weights(sd.__getitem__, lora, mode="side")
    assert set(plan.values()) == {"side"} and len(plan) == 8 + 1  # 8 packed matrices of the double block, qkvmlp of the single
    out = _engine_fwd(eng, hs, inp, t, g).clone()
    torch.cuda.synchronize()
    floor = _rel(ref16, ref32)
    e16, e32, effect = _rel(out, ref16), _rel(out, ref32), _rel(base16, ref16)
    print(f"config-4 LoRA r16 side path: floor {floor:.3e} | engine vs unfused bf16 {e16:.3e}, vs fp32 {e32:.3e} | adapter effect {effect:.3e}")
    assert e16 <= 2.0 * floor and e32 <= 1.5 * floor, (e16, e32, floor)
    assert _cosdist(out, ref16) < 1e-4 and effect > 5 * e32
    eng.unload_lora_weights(sd.__getitem__)
    assert torch.equal(_engine_fwd(eng, hs, inp, t, g), base_out)  # base weights were never touched


def test_auto_mode_folds_large_deltas_and_keeps_ulp_sized_ones_unfused():
    """A delta far above W's bf16 ulp folds; one near the ulp would be rounded away by a fold and takes the side path, where the
    engine still resolves it (closer to the unfused fp32 result than a folded engine is)."""
    from textflux_b200 import B200FluxTransformer, fold_lora
    from textflux_b200.packer import fold_noise
    cfg, sd, inp, t, g, hs = _setup(S_grid=(32, 32), T=128)
    big, small = make_lora(sd, 16, seed=8, b_std=0.02), make_lora(sd, 16, seed=9, b_std=2e-4)
    cfgns = cfg
    n_big, n_small = fold_noise(cfgns, sd.__getitem__, big), fold_noise(cfgns, sd.__getitem__, small)
    print("fold noise, large delta:", {k: round(v, 4) for k, v in n_big.items()})
    print("fold noise, ulp-sized delta:", {k: round(v, 3) for k, v in n_small.items()})
    eng = B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0")
    assert set(eng.load_lora_weights(sd.__getitem__, big, mode="auto").values()) == {"fold"}
    assert set(eng.load_lora_weights(sd.__getitem__, small, mode="auto").values()) == {"side"}
    out_side = _engine_fwd(eng, hs, inp, t, g).clone()
    with unfused_lora(small):
        ref32 = _oracle(sd, cfg, hs, inp, t, g, fp32=True)
    base32 = _oracle(sd, cfg, hs, inp, t, g, fp32=True)
    folded = B200FluxTransformer(cfg.to_dict(), fold_lora(sd.__getitem__, small), device="cuda:0")
    out_fold = _engine_fwd(folded, hs, inp, t, g)
    # the adapter's own contribution, as each engine realises it, against the fp32 truth of that contribution
    base_eng = _engine_fwd(B200FluxTransformer.from_state_dict(cfg.to_dict(), sd, device="cuda:0"), hs, inp, t, g)
    d_true = (ref32 - base32).float()
    d_side, d_fold = (out_side.float() - base_eng.float()), (out_fold.float() - base_eng.float())
    c_side = torch.nn.functional.cosine_similarity(d_side.flatten(), d_true.flatten(), dim=0).item()
    c_fold = torch.nn.functional.cosine_similarity(d_fold.flatten(), d_true.flatten(), dim=0).item()
    print(f"ulp-sized adapter: cosine of realised vs true contribution, side {c_side:.3f}, fold {c_fold:.3f}")
    assert c_side > c_fold
