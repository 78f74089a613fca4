"""B200T5Encoder / B200CLIPTextEncoder: drop-ins for `pipe.text_encoder_2` (T5EncoderModel) and `pipe.text_encoder` (CLIPTextModel)
as FluxFillPipeline calls them (pipelines/flux/pipeline_flux_fill.py:1411-1503):

    prompt_embeds        = self.text_encoder_2(text_input_ids.to(device), output_hidden_states=False)[0]          # :1438
    pooled_prompt_embeds = self.text_encoder(text_input_ids.to(device), output_hidden_states=False).pooler_output  # :1483-1486
    dtype = self.text_encoder_2.dtype / self.text_encoder.dtype                                                    # :1440, :1489

Tokenisation stays with transformers' tokenizers (host side, untouched).  Everything from input_ids on runs through the C ABI
(tfx_textenc_*, include/textflux_b200.h); there is no CPU or torch fallback.  TextFlux's CLIP prompt is a constant template
(run_inference.py:27-40), so the CLIP mirror keeps the pooled embedding of every token sequence it has seen.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Callable, Dict, Optional, Union

import torch

from . import _lib
from .engine import FrozenConfig

Tensor = torch.Tensor


def _cfg_dict(config) -> dict:
    if isinstance(config, dict):
        return dict(config)
    if hasattr(config, "to_dict"):
        return dict(config.to_dict())
    return {k: getattr(config, k) for k in dir(config) if not k.startswith("_") and not callable(getattr(config, k))}


class _Output(tuple):
    """What the pipeline indexes: `[0]`, `.last_hidden_state`, `.pooler_output` (transformers' ModelOutput duck type)."""

    def __new__(cls, last_hidden_state, pooler_output=None):
        items = (last_hidden_state,) if pooler_output is None else (last_hidden_state, pooler_output)
        self = super().__new__(cls, items)
        self.last_hidden_state = last_hidden_state
        self.pooler_output = pooler_output
        return self


class _TextEncoderBase(torch.nn.Module):
    def __init__(self, tc: "_lib.TfxTextEncConfig", packed: Dict[str, Tensor], device: torch.device):
        super().__init__()
        self._lib = _lib.load()
        self._dev = device
        h = C.c_void_p()
        _lib.check(self._lib.tfx_textenc_create(C.byref(tc), device.index, C.byref(h)))
        self._h = h
        self.register_buffer("_anchor", torch.zeros(1, dtype=torch.bfloat16, device=device), persistent=False)
        self._weights = packed
        for name, t in packed.items():
            _lib.check(self._lib.tfx_textenc_set_weight(self._h, name.encode(), t.data_ptr(), t.shape[0], t.shape[1]), self._h, textenc=True)

    @staticmethod
    def _device(device) -> torch.device:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("textflux_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        return torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev

    @property
    def device(self) -> torch.device:
        return self._dev

    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    def to(self, *args, **kwargs):
        dev, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if dev is not None and (torch.device(dev).type != "cuda" or torch.device(dev).index not in (None, self._dev.index)):
            raise RuntimeError(f"textflux_b200: the text encoder lives on {self._dev}")
        if dtype is not None and dtype != torch.bfloat16:
            raise RuntimeError("textflux_b200: the text encoders compute in bf16 only")
        return self

    def counter(self, key: str) -> int:
        v = C.c_int64()
        _lib.check(self._lib.tfx_textenc_get_counter(self._h, key.encode(), C.byref(v)), self._h, textenc=True)
        return v.value

    def __del__(self):
        h, lib = getattr(self, "_h", None), getattr(self, "_lib", None)
        if h and lib:
            lib.tfx_textenc_destroy(h)
            self._h = None

    def _ids(self, input_ids: Tensor) -> Tensor:
        if input_ids is None or input_ids.ndim != 2:
            raise ValueError("input_ids [batch, sequence] is required")
        if input_ids.dtype not in (torch.int64, torch.int32):
            raise ValueError(f"input_ids dtype {input_ids.dtype} unsupported")
        return input_ids.to(device=self._dev, dtype=torch.int32).contiguous()

    def _encode(self, ids: Tensor, lut: Optional[Tensor], pooled_index: Optional[Tensor]):
        B, T = ids.shape
        D = self.config["d_model" if "d_model" in self.config else "hidden_size"]
        out = torch.empty(B, T, D, device=self._dev, dtype=torch.bfloat16)
        pooled = torch.empty(B, D, device=self._dev, dtype=torch.bfloat16) if pooled_index is not None else None
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.tfx_textenc_encode(self._h, ids.data_ptr(), B, T, None if lut is None else lut.data_ptr(), out.data_ptr(),
                                                    None if pooled_index is None else pooled_index.data_ptr(),
                                                    None if pooled is None else pooled.data_ptr(),
                                                    torch.cuda.current_stream(self._dev).cuda_stream), self._h, textenc=True)
        return out, pooled


def t5_bucket_lut(T: int, num_buckets: int, max_distance: int) -> Tensor:
    """Bucket of every relative position d = key - query in [-(T-1), T-1], at index d + T - 1: T5Attention._relative_position_bucket
    (transformers models/t5/modeling_t5.py, bidirectional) with the same torch ops in the same order -- host-side index math."""
    rp = torch.arange(-(T - 1), T, dtype=torch.long)
    nb = num_buckets // 2
    buckets = (rp > 0).to(torch.long) * nb
    rp = torch.abs(rp)
    max_exact = nb // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return (buckets + torch.where(is_small, rp, large)).to(torch.int32)


T5_XXL_CONFIG = dict(vocab_size=32128, d_model=4096, d_kv=64, num_heads=64, num_layers=24, d_ff=10240, relative_attention_num_buckets=32,
                     relative_attention_max_distance=128, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu")
CLIP_L_CONFIG = dict(vocab_size=49408, hidden_size=768, num_attention_heads=12, num_hidden_layers=12, intermediate_size=3072,
                     max_position_embeddings=77, layer_norm_eps=1e-5, eos_token_id=2, hidden_act="quick_gelu")


def t5_reference_names(cfg) -> list:
    """(name, shape) of the T5EncoderModel.state_dict() tensors the encoder reads."""
    inner, D, F = cfg["num_heads"] * cfg["d_kv"], cfg["d_model"], cfg["d_ff"]
    out = [("shared.weight", (cfg["vocab_size"], D)),
           ("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", (cfg.get("relative_attention_num_buckets", 32), cfg["num_heads"]))]
    for i in range(cfg["num_layers"]):
        a, f = f"encoder.block.{i}.layer.0", f"encoder.block.{i}.layer.1"
        out += [(f"{a}.SelfAttention.{m}.weight", (inner, D)) for m in ("q", "k", "v")]
        out += [(f"{a}.SelfAttention.o.weight", (D, inner)), (f"{a}.layer_norm.weight", (D,)), (f"{f}.layer_norm.weight", (D,)),
                (f"{f}.DenseReluDense.wi_0.weight", (F, D)), (f"{f}.DenseReluDense.wi_1.weight", (F, D)), (f"{f}.DenseReluDense.wo.weight", (D, F))]
    out.append(("encoder.final_layer_norm.weight", (D,)))
    return out


def clip_reference_names(cfg) -> list:
    """(name, shape) of the CLIPTextModel.state_dict() tensors the encoder reads."""
    D, F = cfg["hidden_size"], cfg["intermediate_size"]
    out = [("text_model.embeddings.token_embedding.weight", (cfg["vocab_size"], D)),
           ("text_model.embeddings.position_embedding.weight", (cfg["max_position_embeddings"], D))]
    for i in range(cfg["num_hidden_layers"]):
        L = f"text_model.encoder.layers.{i}."
        for m in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out += [(f"{L}self_attn.{m}.weight", (D, D)), (f"{L}self_attn.{m}.bias", (D,))]
        out += [(L + "layer_norm1.weight", (D,)), (L + "layer_norm1.bias", (D,)), (L + "layer_norm2.weight", (D,)), (L + "layer_norm2.bias", (D,)),
                (L + "mlp.fc1.weight", (F, D)), (L + "mlp.fc1.bias", (F,)), (L + "mlp.fc2.weight", (D, F)), (L + "mlp.fc2.bias", (D,))]
    out += [("text_model.final_layer_norm.weight", (D,)), ("text_model.final_layer_norm.bias", (D,))]
    return out


class B200T5Encoder(_TextEncoderBase):
    """Drop-in for T5EncoderModel (T5 v1.1 encoder: gated-gelu, RMS layer norm, relative position bias on layer 0 shared by all)."""

    def __init__(self, config, get: Callable[[str], Tensor], device: Union[str, torch.device] = "cuda"):
        cfg = _cfg_dict(config)
        ffp = cfg.get("feed_forward_proj", "gated-gelu")
        if ffp != "gated-gelu" or cfg.get("is_gated_act", True) is not True:
            raise ValueError(f"textflux_b200: only the gated-gelu T5 v1.1 feed-forward is implemented, got feed_forward_proj={ffp!r}")
        if cfg["d_kv"] != 64:
            raise ValueError(f"textflux_b200: T5 d_kv={cfg['d_kv']} unsupported (64)")
        dev = self._device(device)
        H, L = cfg["num_heads"], cfg["num_layers"]
        P: Dict[str, Tensor] = {}

        def put(name, t):
            t = t.to(device=dev, dtype=torch.bfloat16)
            P[name] = (t.reshape(1, -1) if t.ndim == 1 else t).contiguous()

        emb = "shared.weight"
        try:
            put("embed", get(emb))
        except KeyError:
            put("embed", get("encoder.embed_tokens.weight"))
        put("rel_bias", get("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"))
        for i in range(L):
            a, f = f"encoder.block.{i}.layer.0", f"encoder.block.{i}.layer.1"
            put(f"l{i}.ln1.w", get(f"{a}.layer_norm.weight"))
            put(f"l{i}.qkv.w", torch.cat([get(f"{a}.SelfAttention.{m}.weight").to(dev) for m in ("q", "k", "v")], dim=0))
            put(f"l{i}.o.w", get(f"{a}.SelfAttention.o.weight"))
            put(f"l{i}.ln2.w", get(f"{f}.layer_norm.weight"))
            put(f"l{i}.wi0.w", get(f"{f}.DenseReluDense.wi_0.weight"))
            put(f"l{i}.wi1.w", get(f"{f}.DenseReluDense.wi_1.weight"))
            put(f"l{i}.wo.w", get(f"{f}.DenseReluDense.wo.weight"))
        put("final_ln.w", get("encoder.final_layer_norm.weight"))
        tc = _lib.TfxTextEncConfig(0, cfg["vocab_size"], cfg["d_model"], cfg["d_kv"], H, L, cfg["d_ff"], 0,
                                   cfg.get("relative_attention_num_buckets", 32), cfg.get("relative_attention_max_distance", 128),
                                   float(cfg.get("layer_norm_epsilon", 1e-6)))
        with torch.cuda.device(dev):
            super().__init__(tc, P, dev)
        self.config = FrozenConfig(cfg)
        self._luts: Dict[int, Tensor] = {}

    @classmethod
    def from_reference(cls, module: torch.nn.Module, device="cuda") -> "B200T5Encoder":
        sd = module.state_dict()
        return cls(module.config, sd.__getitem__, device=device)

    @torch.no_grad()
    def forward(self, input_ids: Tensor = None, attention_mask: Optional[Tensor] = None, output_hidden_states: bool = False, **kw):
        if output_hidden_states:
            raise NotImplementedError("textflux_b200: only the last hidden state is produced (the pipeline asks for nothing else)")
        if attention_mask is not None and not bool((attention_mask != 0).all()):
            raise NotImplementedError("textflux_b200: padding masks are not implemented -- FluxFillPipeline passes none, pad tokens are attended")
        ids = self._ids(input_ids)
        T = ids.shape[1]
        if T not in self._luts:
            self._luts[T] = t5_bucket_lut(T, self.config["relative_attention_num_buckets"], self.config["relative_attention_max_distance"]).to(self._dev)
        out, _ = self._encode(ids, self._luts[T], None)
        return _Output(out)


class B200CLIPTextEncoder(_TextEncoderBase):
    """Drop-in for CLIPTextModel (pre-LN transformer, causal mask, quick_gelu, EOS pooling)."""

    def __init__(self, config, get: Callable[[str], Tensor], device: Union[str, torch.device] = "cuda", cache: bool = True):
        cfg = _cfg_dict(config)
        if cfg.get("hidden_act", "quick_gelu") != "quick_gelu":
            raise ValueError(f"textflux_b200: CLIP hidden_act={cfg['hidden_act']!r} unsupported (quick_gelu)")
        D, H, L = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_hidden_layers"]
        if D // H != 64 or D % H:
            raise ValueError(f"textflux_b200: CLIP head dimension {D / H} unsupported (64)")
        dev = self._device(device)
        P: Dict[str, Tensor] = {}

        def put(name, t):
            t = t.to(device=dev, dtype=torch.bfloat16)
            P[name] = (t.reshape(1, -1) if t.ndim == 1 else t).contiguous()

        pre = "text_model."
        put("embed", get(pre + "embeddings.token_embedding.weight"))
        put("pos", get(pre + "embeddings.position_embedding.weight"))
        for i in range(L):
            lp = f"{pre}encoder.layers.{i}."
            for s, n in (("ln1", "layer_norm1"), ("ln2", "layer_norm2"), ("o", "self_attn.out_proj"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
                put(f"l{i}.{s}.w", get(lp + n + ".weight"))
                put(f"l{i}.{s}.b", get(lp + n + ".bias"))
            put(f"l{i}.qkv.w", torch.cat([get(f"{lp}self_attn.{m}_proj.weight").to(dev) for m in ("q", "k", "v")], dim=0))
            put(f"l{i}.qkv.b", torch.cat([get(f"{lp}self_attn.{m}_proj.bias").to(dev) for m in ("q", "k", "v")], dim=0))
        put("final_ln.w", get(pre + "final_layer_norm.weight"))
        put("final_ln.b", get(pre + "final_layer_norm.bias"))
        tc = _lib.TfxTextEncConfig(1, cfg["vocab_size"], D, 64, H, L, cfg["intermediate_size"], cfg["max_position_embeddings"], 0, 0,
                                   float(cfg.get("layer_norm_eps", 1e-5)))
        with torch.cuda.device(dev):
            super().__init__(tc, P, dev)
        self.config = FrozenConfig(cfg)
        self._cache: Optional[Dict[bytes, tuple]] = {} if cache else None
        self.cache_hits = 0

    @classmethod
    def from_reference(cls, module: torch.nn.Module, device="cuda", **kw) -> "B200CLIPTextEncoder":
        sd = module.state_dict()
        return cls(module.config, sd.__getitem__, device=device, **kw)

    def pooled_index(self, input_ids: Tensor) -> Tensor:
        """CLIPTextTransformer.forward's EOS rule (host-side index logic on the token ids)."""
        eos = self.config.get("eos_token_id", 2)
        ids = input_ids.to(torch.int)
        return (ids.argmax(dim=-1) if eos == 2 else (ids == eos).int().argmax(dim=-1)).to(torch.int32)

    @torch.no_grad()
    def forward(self, input_ids: Tensor = None, attention_mask: Optional[Tensor] = None, output_hidden_states: bool = False, **kw):
        if output_hidden_states:
            raise NotImplementedError("textflux_b200: only last_hidden_state / pooler_output are produced")
        if attention_mask is not None and not bool((attention_mask != 0).all()):
            raise NotImplementedError("textflux_b200: padding masks are not implemented -- FluxFillPipeline passes none")
        ids = self._ids(input_ids)
        key = None
        if self._cache is not None:
            key = ids.cpu().numpy().tobytes() + bytes(ids.shape)
            if key in self._cache:
                self.cache_hits += 1
                lh, po = self._cache[key]
                return _Output(lh.clone(), po.clone())
        out, pooled = self._encode(ids, None, self.pooled_index(ids).contiguous())
        if key is not None:
            if len(self._cache) >= 64:
                self._cache.pop(next(iter(self._cache)))
            self._cache[key] = (out.clone(), pooled.clone())
        return _Output(out, pooled)
