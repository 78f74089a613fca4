"""CPU/torch restatement of the prompt encoders FluxFillPipeline calls.

TEST INFRASTRUCTURE ONLY: imported by tests/, oracle/make_golden_textenc.py and nothing under textflux_b200/.

The algorithm lives in a THIRD-PARTY dependency that is not under /root/reference: `transformers` (reference pin 4.43.3,
/root/reference/requirements.txt:4; installed in this image: 5.5.0).  Call sites in the reference:
  diffusers/src/diffusers/pipelines/flux/pipeline_flux_fill.py:1411-1458  _get_t5_prompt_embeds: tokenizer_2(max_length=512, padding="max_length")
        -> text_encoder_2(input_ids, output_hidden_states=False)[0]            (T5EncoderModel, google/t5-v1_1-xxl encoder)
  :1460-1503  _get_clip_prompt_embeds: tokenizer(max_length=77) -> text_encoder(input_ids, output_hidden_states=False).pooler_output
        (CLIPTextModel, openai/clip-vit-large-patch14 text tower);  run_inference.py:27-40 fixes the CLIP prompt to a constant template.
Restated from transformers/models/t5/modeling_t5.py (T5LayerNorm, T5Attention incl. _relative_position_bucket / compute_bias,
T5DenseGatedActDense, T5Block, T5Stack) and models/clip/modeling_clip.py (CLIPTextEmbeddings, eager_attention_forward, CLIPMLP,
CLIPEncoderLayer, CLIPTextTransformer incl. the EOS pooling rule).
Pinned bit-exactly (torch.equal, fp32 and bf16) to the INSTALLED transformers modules run in the build container with
attn_implementation="eager": tests/golden/textenc_*.pt (oracle/make_golden_textenc.py), tests/test_oracle_golden.py::test_textenc_*.
The 4.43.3 pin itself cannot be installed offline; its T5 path is the same arithmetic, its CLIP attention scales q before the
product and takes the softmax in the input dtype (a rounding-order difference inside the bf16 floor).
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class T5Cfg:
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    num_heads: int = 64
    num_layers: int = 24
    d_ff: int = 10240
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6

    def to_dict(self):
        return asdict(self)


@dataclass(frozen=True)
class ClipCfg:
    vocab_size: int = 49408
    hidden_size: int = 768
    num_attention_heads: int = 12
    num_hidden_layers: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 77
    layer_norm_eps: float = 1e-5
    eos_token_id: int = 2  # openai/clip-vit-large-patch14's config.json: the legacy value that selects the argmax pooling rule

    def to_dict(self):
        return asdict(self)


T5_XXL, T5_TINY = T5Cfg(), T5Cfg(vocab_size=384, d_model=256, num_heads=4, num_layers=2, d_ff=512)
CLIP_L, CLIP_TINY = ClipCfg(), ClipCfg(vocab_size=1000, hidden_size=256, num_attention_heads=4, num_hidden_layers=2, intermediate_size=512)


# ------------------------------------------------------------------------------------------------------------------ T5
def t5_spec(cfg: T5Cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    inner = cfg.num_heads * cfg.d_kv
    out = [("shared.weight", (cfg.vocab_size, cfg.d_model), "e")]
    for i in range(cfg.num_layers):
        a, f = f"encoder.block.{i}.layer.0", f"encoder.block.{i}.layer.1"
        out.append((f"{a}.SelfAttention.q.weight", (inner, cfg.d_model), "wq"))
        for m in ("k", "v"):
            out.append((f"{a}.SelfAttention.{m}.weight", (inner, cfg.d_model), "w"))
        out.append((f"{a}.SelfAttention.o.weight", (cfg.d_model, inner), "w"))
        if i == 0:
            out.append((f"{a}.SelfAttention.relative_attention_bias.weight", (cfg.relative_attention_num_buckets, cfg.num_heads), "r"))
        out.append((f"{a}.layer_norm.weight", (cfg.d_model,), "g"))
        out.append((f"{f}.DenseReluDense.wi_0.weight", (cfg.d_ff, cfg.d_model), "w"))
        out.append((f"{f}.DenseReluDense.wi_1.weight", (cfg.d_ff, cfg.d_model), "w"))
        out.append((f"{f}.DenseReluDense.wo.weight", (cfg.d_model, cfg.d_ff), "w"))
        out.append((f"{f}.layer_norm.weight", (cfg.d_model,), "g"))
    out.append(("encoder.final_layer_norm.weight", (cfg.d_model,), "g"))
    return out


def clip_spec(cfg: ClipCfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    D, Fd = cfg.hidden_size, cfg.intermediate_size
    out = [("text_model.embeddings.token_embedding.weight", (cfg.vocab_size, D), "e"),
           ("text_model.embeddings.position_embedding.weight", (cfg.max_position_embeddings, D), "p")]
    for i in range(cfg.num_hidden_layers):
        L = f"text_model.encoder.layers.{i}"
        for m in ("k_proj", "v_proj", "q_proj", "out_proj"):
            out += [(f"{L}.self_attn.{m}.weight", (D, D), "w"), (f"{L}.self_attn.{m}.bias", (D,), "b")]
        out += [(f"{L}.layer_norm1.weight", (D,), "g"), (f"{L}.layer_norm1.bias", (D,), "b")]
        out += [(f"{L}.mlp.fc1.weight", (Fd, D), "w"), (f"{L}.mlp.fc1.bias", (Fd,), "b")]
        out += [(f"{L}.mlp.fc2.weight", (D, Fd), "w"), (f"{L}.mlp.fc2.bias", (D,), "b")]
        out += [(f"{L}.layer_norm2.weight", (D,), "g"), (f"{L}.layer_norm2.bias", (D,), "b")]
    out += [("text_model.final_layer_norm.weight", (D,), "g"), ("text_model.final_layer_norm.bias", (D,), "b")]
    return out


def init_state_dict(spec, seed: int, dtype=torch.float32, device="cpu") -> Dict[str, Tensor]:
    """Synthetic weights: embeddings N(0,1), linears N(0, 1/fan_in), biases N(0, 0.02^2), norm weights 1 + N(0, 0.1^2),
    relative-attention bias N(0, 0.5^2).  T5 does not scale its scores by d_kv^-0.5 (the factor lives in the trained weights; T5's own
    initialiser gives q a std of (d_model * d_kv)^-0.5): its q projections get that extra d_kv^-0.5 here, otherwise the logits have a
    std of 8, every softmax is one-hot and 24 layers amplify bf16 rounding into noise (reference bf16-vs-fp32 rel-L2 0.9)."""
    sd = {}
    for idx, (name, shape, kind) in enumerate(spec):
        g = torch.Generator(device=device).manual_seed(seed * 100003 + idx)
        t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if kind == "w":
            t = t * (1.0 / shape[1]) ** 0.5
        elif kind == "wq":
            t = t * (1.0 / (shape[1] * 64)) ** 0.5
        elif kind == "b":
            t = t * 0.02
        elif kind == "g":
            t = 1.0 + 0.1 * t
        elif kind == "r":
            t = t * 0.5
        elif kind == "p":
            t = t * 0.1
        sd[name] = t.to(dtype)
    return sd


def t5_relative_position_bucket(relative_position: Tensor, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """T5Attention._relative_position_bucket with bidirectional=True (the encoder), operation for operation."""
    relative_buckets = 0
    num_buckets //= 2
    relative_buckets += (relative_position > 0).to(torch.long) * num_buckets
    relative_position = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    relative_position_if_large = max_exact + (
        torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    relative_position_if_large = torch.min(relative_position_if_large, torch.full_like(relative_position_if_large, num_buckets - 1))
    relative_buckets += torch.where(is_small, relative_position, relative_position_if_large)
    return relative_buckets


def new_gelu(x: Tensor) -> Tensor:
    """transformers.activations.NewGELUActivation ("gelu_new", T5 v1.1's dense_act_fn): the tanh GELU written out op by op."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def _t5_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """T5LayerNorm.forward."""
    variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(variance + eps)
    if w.dtype in (torch.float16, torch.bfloat16):
        x = x.to(w.dtype)
    return w * x


@torch.no_grad()
def t5_encode(sd: Dict[str, Tensor], cfg: T5Cfg, input_ids: Tensor) -> Tensor:
    """T5EncoderModel(input_ids)[0] with attention_mask=None (the pipeline passes none: padding tokens are attended)."""
    B, T = input_ids.shape
    H, dk = cfg.num_heads, cfg.d_kv
    x = F.embedding(input_ids, sd["shared.weight"])
    ctx = torch.arange(T, dtype=torch.long, device=input_ids.device)[:, None]
    mem = torch.arange(T, dtype=torch.long, device=input_ids.device)[None, :]
    buckets = t5_relative_position_bucket(mem - ctx, cfg.relative_attention_num_buckets, cfg.relative_attention_max_distance)
    bias = F.embedding(buckets, sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]).permute([2, 0, 1]).unsqueeze(0)
    for i in range(cfg.num_layers):
        a, f = f"encoder.block.{i}.layer.0", f"encoder.block.{i}.layer.1"
        n = _t5_norm(x, sd[f"{a}.layer_norm.weight"], cfg.layer_norm_epsilon)
        q = F.linear(n, sd[f"{a}.SelfAttention.q.weight"]).view(B, T, H, dk).transpose(1, 2)
        k = F.linear(n, sd[f"{a}.SelfAttention.k.weight"]).view(B, T, H, dk).transpose(1, 2)
        v = F.linear(n, sd[f"{a}.SelfAttention.v.weight"]).view(B, T, H, dk).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(3, 2))
        scores += bias
        w = F.softmax(scores.float(), dim=-1).type_as(scores)
        o = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, T, -1)
        x = x + F.linear(o, sd[f"{a}.SelfAttention.o.weight"])
        n = _t5_norm(x, sd[f"{f}.layer_norm.weight"], cfg.layer_norm_epsilon)
        g = new_gelu(F.linear(n, sd[f"{f}.DenseReluDense.wi_0.weight"]))
        h = g * F.linear(n, sd[f"{f}.DenseReluDense.wi_1.weight"])
        x = x + F.linear(h, sd[f"{f}.DenseReluDense.wo.weight"])
    return _t5_norm(x, sd["encoder.final_layer_norm.weight"], cfg.layer_norm_epsilon)


# ------------------------------------------------------------------------------------------------------------------ CLIP
def clip_pooled_index(input_ids: Tensor, eos_token_id: int) -> Tensor:
    """CLIPTextTransformer.forward's EOS rule: legacy configs (eos_token_id == 2) take argmax(input_ids) (the EOS id is the largest
    in the vocabulary), newer ones the first position equal to eos_token_id."""
    if eos_token_id == 2:
        return input_ids.to(torch.int).argmax(dim=-1)
    return (input_ids.to(torch.int) == eos_token_id).int().argmax(dim=-1)


@torch.no_grad()
def clip_encode(sd: Dict[str, Tensor], cfg: ClipCfg, input_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """CLIPTextModel(input_ids) -> (last_hidden_state, pooler_output), eager attention, causal mask, no padding mask."""
    B, T = input_ids.shape
    D, H = cfg.hidden_size, cfg.num_attention_heads
    dh = D // H
    P = "text_model."
    x = F.embedding(input_ids, sd[P + "embeddings.token_embedding.weight"]) + sd[P + "embeddings.position_embedding.weight"][:T][None]
    mask = torch.full((T, T), torch.finfo(x.dtype).min, dtype=x.dtype, device=x.device).triu(1)[None, None]
    for i in range(cfg.num_hidden_layers):
        L = f"{P}encoder.layers.{i}."
        n = F.layer_norm(x, (D,), sd[L + "layer_norm1.weight"], sd[L + "layer_norm1.bias"], cfg.layer_norm_eps)
        q = F.linear(n, sd[L + "self_attn.q_proj.weight"], sd[L + "self_attn.q_proj.bias"]).view(B, T, H, dh).transpose(1, 2)
        k = F.linear(n, sd[L + "self_attn.k_proj.weight"], sd[L + "self_attn.k_proj.bias"]).view(B, T, H, dh).transpose(1, 2)
        v = F.linear(n, sd[L + "self_attn.v_proj.weight"], sd[L + "self_attn.v_proj.bias"]).view(B, T, H, dh).transpose(1, 2)
        w = torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5
        w = w + mask
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(B, T, -1).contiguous()
        x = x + F.linear(o, sd[L + "self_attn.out_proj.weight"], sd[L + "self_attn.out_proj.bias"])
        n = F.layer_norm(x, (D,), sd[L + "layer_norm2.weight"], sd[L + "layer_norm2.bias"], cfg.layer_norm_eps)
        h = F.linear(n, sd[L + "mlp.fc1.weight"], sd[L + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)  # QuickGELUActivation
        x = x + F.linear(h, sd[L + "mlp.fc2.weight"], sd[L + "mlp.fc2.bias"])
    x = F.layer_norm(x, (D,), sd[P + "final_layer_norm.weight"], sd[P + "final_layer_norm.bias"], cfg.layer_norm_eps)
    idx = clip_pooled_index(input_ids, cfg.eos_token_id)
    return x, x[torch.arange(B, device=x.device), idx]
