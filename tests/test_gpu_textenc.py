"""Prompt encoders on the GPU: T5 v1.1 encoder and CLIP text model through the C ABI against (a) golden
fixtures written by the installed transformers modules (oracle/make_golden_textenc.py) and (b) the oracle restatement run on CUDA
tensors at the real widths (T5-XXL: d_model 4096, 64 heads, d_ff 10240, T = 512; CLIP-L: 768 / 12 heads / 3072, T = 77).

Bar: engine-vs-reference-bf16 rel-L2 <= 2x the reference's own bf16-vs-fp32 rel-L2 on the same inputs,
engine-vs-fp32 <= 1.5x that, cosine distance < 1e-3 against fp32."""
import os

import pytest
import torch

from oracle import textenc_oracle as to

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


def _cosdist(a, b):
    return 1.0 - torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def _bar(name, out, ref16, ref32):
    floor = _rel(ref16, ref32)
    e16, e32 = _rel(out, ref16), _rel(out, ref32)
    print(f"{name}: reference bf16-vs-fp32 floor {floor:.3e} | engine vs bf16 {e16:.3e}, vs fp32 {e32:.3e}, cosdist vs fp32 {_cosdist(out, ref32):.2e}")
    assert e16 <= 2.0 * floor and e32 <= 1.5 * floor, (name, e16, e32, floor)
    assert _cosdist(out, ref32) < 1e-3


def _t5(cfg, sd):
    from textflux_b200 import B200T5Encoder
    return B200T5Encoder(dict(cfg.to_dict(), feed_forward_proj="gated-gelu"), {k: v.cuda() for k, v in sd.items()}.__getitem__, device="cuda:0")


def _clip(cfg, sd, **kw):
    from textflux_b200 import B200CLIPTextEncoder
    return B200CLIPTextEncoder(dict(cfg.to_dict(), hidden_act="quick_gelu"), {k: v.cuda() for k, v in sd.items()}.__getitem__, device="cuda:0", **kw)


def test_t5_against_transformers_golden():
    d = torch.load(os.path.join(GOLDEN, "textenc_t5.pt"))
    cfg = to.T5Cfg(**d["config"])
    enc = _t5(cfg, to.init_state_dict(to.t5_spec(cfg), d["seed"]))
    out = enc(d["input_ids"].cuda(), output_hidden_states=False)
    assert out[0].shape == d["last_hidden_bf16"].shape and out[0].dtype == torch.bfloat16 and out.last_hidden_state is out[0]
    _bar("T5 tiny", out[0].cpu(), d["last_hidden_bf16"], d["last_hidden_f32"])
    assert enc.counter("launches") > 0 and enc.dtype == torch.bfloat16


def test_clip_against_transformers_golden():
    d = torch.load(os.path.join(GOLDEN, "textenc_clip.pt"))
    cfg = to.ClipCfg(**d["config"])
    enc = _clip(cfg, to.init_state_dict(to.clip_spec(cfg), d["seed"]))
    out = enc(d["input_ids"].cuda(), output_hidden_states=False)
    _bar("CLIP tiny last_hidden_state", out.last_hidden_state.cpu(), d["last_hidden_bf16"], d["last_hidden_f32"])
    _bar("CLIP tiny pooler_output", out.pooler_output.cpu(), d["pooled_bf16"], d["pooled_f32"])
    # the pooled row IS a row of the last hidden state, at the reference's EOS index
    idx = to.clip_pooled_index(d["input_ids"], cfg.eos_token_id)
    assert torch.equal(out.pooler_output, out.last_hidden_state[torch.arange(2), idx.cuda()])
    # constant prompt (run_inference.py:27-40): the second call with the same ids is served from the cache, bit for bit, without launches
    l0 = enc.counter("launches")
    again = enc(d["input_ids"].cuda())
    assert enc.counter("launches") == l0 and enc.cache_hits == 1 and torch.equal(again.pooler_output, out.pooler_output)


@pytest.mark.parametrize("T", [1, 5, 16, 32, 33, 100])
def test_short_and_ragged_sequence_lengths(T):
    """Sequence lengths below and off the 16-row MMA block and the 32-row query block (the score tile doubles as the output staging
    tile: T = 32 once overflowed it), two prompts, both encoders, against the CUDA-run oracle."""
    tcfg, ccfg = to.T5_TINY, to.CLIP_TINY
    tsd = to.init_state_dict(to.t5_spec(tcfg), 5, device="cuda")
    csd = to.init_state_dict(to.clip_spec(ccfg), 6, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(T)
    ids = torch.randint(2, tcfg.vocab_size, (2, T), generator=g, device="cuda")
    out = _t5(tcfg, tsd)(ids)[0]
    _bar(f"T5 T={T}", out, to.t5_encode({k: v.to(torch.bfloat16) for k, v in tsd.items()}, tcfg, ids), to.t5_encode(tsd, tcfg, ids))
    if T <= ccfg.max_position_embeddings:
        cids = torch.randint(3, ccfg.vocab_size - 2, (2, T), generator=g, device="cuda")
        cids[0, T // 2] = ccfg.vocab_size - 1
        cids[1, T - 1] = ccfg.vocab_size - 1
        o = _clip(ccfg, csd, cache=False)(cids)
        lh16, po16 = to.clip_encode({k: v.to(torch.bfloat16) for k, v in csd.items()}, ccfg, cids)
        lh32, po32 = to.clip_encode(csd, ccfg, cids)
        _bar(f"CLIP T={T} last_hidden_state", o.last_hidden_state, lh16, lh32)
        _bar(f"CLIP T={T} pooler_output", o.pooler_output, po16, po32)


def test_t5_xxl_width_vs_cuda_oracle():
    """Two layers at T5-XXL's real dimensions, 512 tokens: the shapes every GEMM and the attention kernel see in production."""
    cfg = to.T5Cfg(num_layers=2)
    sd32 = to.init_state_dict(to.t5_spec(cfg), 7, device="cuda")
    sd16 = {k: v.to(torch.bfloat16) for k, v in sd32.items()}
    g = torch.Generator(device="cuda").manual_seed(3)
    ids = torch.randint(2, cfg.vocab_size, (1, 512), generator=g, device="cuda")
    ids[:, 40:] = 0
    ids[:, 39] = 1
    enc = _t5(cfg, sd32)
    out = enc(ids)[0]
    _bar("T5-XXL width, 2 layers, T=512", out, to.t5_encode(sd16, cfg, ids), to.t5_encode(sd32, cfg, ids))


def test_t5_xxl_full_depth_vs_cuda_oracle():
    """The whole T5-XXL encoder (24 layers, 4.76 B parameters), one 512-token prompt padded the way the pipeline pads it."""
    cfg = to.T5_XXL
    sd16 = to.init_state_dict(to.t5_spec(cfg), 7, dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    ids = torch.randint(2, cfg.vocab_size, (1, 512), generator=g, device="cuda")
    ids[:, 70:] = 0
    ids[:, 69] = 1
    enc = _t5(cfg, sd16)
    out = enc(ids)[0]
    ref16 = to.t5_encode(sd16, cfg, ids)
    del enc
    sd32 = {k: v.float() for k, v in sd16.items()}
    del sd16
    torch.cuda.empty_cache()
    _bar("T5-XXL, 24 layers, T=512", out, ref16, to.t5_encode(sd32, cfg, ids))


def test_clip_l_vs_cuda_oracle():
    """The whole CLIP-L text tower (12 layers, 768 wide), T = 77 (not a multiple of the 16-row MMA block), two prompts."""
    cfg = to.CLIP_L
    sd32 = to.init_state_dict(to.clip_spec(cfg), 8, device="cuda")
    sd16 = {k: v.to(torch.bfloat16) for k, v in sd32.items()}
    g = torch.Generator(device="cuda").manual_seed(4)
    ids = torch.randint(3, cfg.vocab_size - 2, (2, 77), generator=g, device="cuda")
    ids[:, 0] = cfg.vocab_size - 2
    ids[0, 20:] = cfg.vocab_size - 1
    ids[1, 76] = cfg.vocab_size - 1
    enc = _clip(cfg, sd32, cache=False)
    out = enc(ids)
    lh16, po16 = to.clip_encode(sd16, cfg, ids)
    lh32, po32 = to.clip_encode(sd32, cfg, ids)
    _bar("CLIP-L last_hidden_state", out.last_hidden_state, lh16, lh32)
    _bar("CLIP-L pooler_output", out.pooler_output, po16, po32)


def test_text_encoders_reject_what_they_do_not_implement():
    from textflux_b200 import B200CLIPTextEncoder, B200T5Encoder
    cfg = to.T5_TINY
    sd = to.init_state_dict(to.t5_spec(cfg), 1)
    with pytest.raises(ValueError):
        B200T5Encoder(dict(cfg.to_dict(), feed_forward_proj="relu"), sd.__getitem__, device="cuda:0")
    with pytest.raises(RuntimeError):
        B200T5Encoder(dict(cfg.to_dict(), feed_forward_proj="gated-gelu"), sd.__getitem__, device="cpu")
    enc = _t5(cfg, sd)
    with pytest.raises(NotImplementedError):
        enc(torch.ones(1, 16, dtype=torch.long, device="cuda"), attention_mask=torch.tensor([[1] * 8 + [0] * 8], device="cuda"))
    with pytest.raises(ValueError):
        enc(torch.ones(16, dtype=torch.long, device="cuda"))
    ccfg = to.CLIP_TINY
    with pytest.raises(ValueError):
        B200CLIPTextEncoder(dict(ccfg.to_dict(), hidden_act="gelu"), to.init_state_dict(to.clip_spec(ccfg), 1).__getitem__, device="cuda:0")
