"""CPU oracle for the TextFlux hot path (TEST INFRASTRUCTURE ONLY).

A functional, dependency-free (torch only) restatement of the reference's per-step
denoising path:

    FluxTransformer2DModel.forward          D/models/transformers/transformer_flux.py:1028-1212
    FluxTransformerBlock.forward            D/models/transformers/transformer_flux.py:794-841
    FluxSingleTransformerBlock.forward      D/models/transformers/transformer_flux.py:715-739
    FluxAttnProcessor2_0.__call__           D/models/attention_processor.py:1979-2060
    RMSNorm.forward                         D/models/normalization.py:532-549
    AdaLayerNormZero / Single / Continuous  D/models/normalization.py:159-171,196-203,361-366
    apply_rotary_emb / FluxPosEmbed         D/models/embeddings.py:879-925,946-973,813-876
    get_timestep_embedding + embed MLPs     D/models/embeddings.py:27-78,976-1040,1318-1339,1906-1932
    FeedForward / GELU(tanh)                D/models/attention.py:1185-1243, D/models/activations.py:65-90
    FlowMatchEulerDiscreteScheduler         D/schedulers/scheduling_flow_match_euler_discrete.py:181-338
    calculate_shift, pack/unpack, ids, loop D/pipelines/flux/pipeline_flux_fill.py:1248-1258,1728-1765,2077-2119

(`D/` = /root/reference/diffusers/src/diffusers.)

Every op is the same ATen call, in the same order and at the same dtype, as the reference
executes, so on the same torch build the oracle is BIT-EXACT against the imported reference
modules (pinned by tests/test_oracle_vs_reference.py in the build container and by the
committed fixtures under tests/golden/ everywhere else).  The reference's own tests hold no
golden values for this path, so the pin is "outputs of the reference itself
run here" (oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (textflux_b200/) never does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (mirrors the @register_to_config signature, transformer_flux.py:865-879)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FluxConfig:
    patch_size: int = 1
    in_channels: int = 384
    out_channels: int = 64
    num_layers: int = 19
    num_single_layers: int = 38
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, ...] = (16, 56, 56)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def to_dict(self) -> dict:
        return dict(
            patch_size=self.patch_size, in_channels=self.in_channels, out_channels=self.out_channels,
            num_layers=self.num_layers, num_single_layers=self.num_single_layers,
            attention_head_dim=self.attention_head_dim, num_attention_heads=self.num_attention_heads,
            joint_attention_dim=self.joint_attention_dim, pooled_projection_dim=self.pooled_projection_dim,
            guidance_embeds=self.guidance_embeds, axes_dims_rope=tuple(self.axes_dims_rope),
        )


#: FLUX.1-Fill-dev 12B (SURVEY.md §8 header; convert_flux_to_diffusers.py:280-289)
FLUX_FILL_12B = FluxConfig()
#: BASELINE.json configs[0]: 2 double + 2 single blocks, dim 256, 16x16 latent, 16 text tokens
TINY = FluxConfig(in_channels=384, out_channels=64, num_layers=2, num_single_layers=2,
                  attention_head_dim=64, num_attention_heads=4, joint_attention_dim=128,
                  pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(8, 28, 28))


# --------------------------------------------------------------------------------------
# state-dict layout and deterministic synthetic weights (§8d)
# --------------------------------------------------------------------------------------
def state_dict_spec(cfg: FluxConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor of the reference state dict; kind in {w,b,rms}."""
    D, dh = cfg.inner_dim, cfg.attention_head_dim
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def lin(name, o, i):
        out.append((name + ".weight", (o, i), "w"))
        out.append((name + ".bias", (o,), "b"))

    lin("time_text_embed.timestep_embedder.linear_1", D, 256)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    if cfg.guidance_embeds:
        lin("time_text_embed.guidance_embedder.linear_1", D, 256)
        lin("time_text_embed.guidance_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim)
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", D, cfg.joint_attention_dim)
    lin("x_embedder", D, cfg.in_channels)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        lin(p + "norm1.linear", 6 * D, D)
        lin(p + "norm1_context.linear", 6 * D, D)
        for n in ("to_q", "to_k", "to_v", "add_k_proj", "add_v_proj", "add_q_proj"):
            lin(p + "attn." + n, D, D)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            out.append((p + "attn." + n + ".weight", (dh,), "rms"))
        lin(p + "attn.to_out.0", D, D)
        lin(p + "attn.to_add_out", D, D)
        lin(p + "ff.net.0.proj", 4 * D, D)
        lin(p + "ff.net.2", D, 4 * D)
        lin(p + "ff_context.net.0.proj", 4 * D, D)
        lin(p + "ff_context.net.2", D, 4 * D)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        lin(p + "norm.linear", 3 * D, D)
        lin(p + "proj_mlp", 4 * D, D)
        lin(p + "proj_out", D, 5 * D)
        for n in ("to_q", "to_k", "to_v"):
            lin(p + "attn." + n, D, D)
        for n in ("norm_q", "norm_k"):
            out.append((p + "attn." + n + ".weight", (dh,), "rms"))
    lin("norm_out.linear", 2 * D, D)
    lin("proj_out", cfg.patch_size * cfg.patch_size * cfg.out_channels, D)
    return out


def init_state_dict(cfg: FluxConfig, seed: int = 1234, dtype: torch.dtype = torch.bfloat16,
                    device: str = "cpu", w_std: float = 0.02, b_std: float = 0.02,
                    rms_std: float = 0.1, only_prefixes: Optional[Sequence[str]] = None) -> Dict[str, Tensor]:
    """Synthetic weights of SURVEY.md §8d: W,b ~ N(0, 0.02^2); RMSNorm weight ~ 1 + N(0, 0.1^2).

    One generator per tensor (seeded from the name's position) so a subset (only_prefixes)
    reproduces the same values as the full dict, on any device.
    """
    sd: Dict[str, Tensor] = {}
    for idx, (name, shape, kind) in enumerate(state_dict_spec(cfg)):
        if only_prefixes is not None and not any(name.startswith(p) for p in only_prefixes):
            continue
        g = torch.Generator(device=device).manual_seed(seed * 100003 + idx)
        t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if kind == "w":
            t = t * w_std
        elif kind == "b":
            t = t * b_std
        else:
            t = 1.0 + t * rms_std
        sd[name] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: Tensor, embedding_dim: int = 256) -> Tensor:
    """embeddings.py:27-78 with flip_sin_to_cos=True, downscale_freq_shift=0, scale=1 (:1322)."""
    half = embedding_dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - 0)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = 1 * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def _mlp2(sd, prefix: str, x: Tensor) -> Tensor:
    """TimestepEmbedding / PixArtAlphaTextProjection: Linear -> SiLU -> Linear (embeddings.py:1009-1021,1927-1932)."""
    x = F.linear(x, sd[prefix + ".linear_1.weight"], sd[prefix + ".linear_1.bias"])
    x = F.silu(x)
    return F.linear(x, sd[prefix + ".linear_2.weight"], sd[prefix + ".linear_2.bias"])


def time_text_embed(sd, cfg: FluxConfig, timestep: Tensor, guidance: Optional[Tensor], pooled: Tensor) -> Tensor:
    """CombinedTimestepGuidanceTextProjEmbeddings.forward (embeddings.py:1327-1339)."""
    t_proj = get_timestep_embedding(timestep)
    t_emb = _mlp2(sd, "time_text_embed.timestep_embedder", t_proj.to(dtype=pooled.dtype))
    if cfg.guidance_embeds:
        g_proj = get_timestep_embedding(guidance)
        g_emb = _mlp2(sd, "time_text_embed.guidance_embedder", g_proj.to(dtype=pooled.dtype))
        tg = t_emb + g_emb
    else:
        tg = t_emb
    p = _mlp2(sd, "time_text_embed.text_embedder", pooled)
    return tg + p


def rope_1d(dim: int, pos: Tensor, theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """get_1d_rotary_pos_embed(use_real=True, repeat_interleave_real=True, float64) (embeddings.py:813-876)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device)[: dim // 2] / dim)) / 1.0
    freqs = torch.outer(pos, freqs)
    cos = freqs.cos().repeat_interleave(2, dim=1).float()
    sin = freqs.sin().repeat_interleave(2, dim=1).float()
    return cos, sin


def flux_pos_embed(ids: Tensor, axes_dim: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """FluxPosEmbed.forward (embeddings.py:952-973)."""
    pos = ids.float()
    cs, ss = [], []
    for i in range(ids.shape[-1]):
        c, s = rope_1d(axes_dim[i], pos[:, i])
        cs.append(c)
        ss.append(s)
    return torch.cat(cs, dim=-1), torch.cat(ss, dim=-1)


def apply_rotary_emb(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """embeddings.py:879-914, use_real_unbind_dim=-1 (interleaved pairs)."""
    cos = cos[None, None]
    sin = sin[None, None]
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


# --------------------------------------------------------------------------------------
# norms / attention / blocks
# --------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Tensor, eps: float = 1e-6) -> Tensor:
    """RMSNorm.forward (normalization.py:532-549)."""
    in_dtype = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.to(weight.dtype)
    x = x * weight
    return x if weight is not None else x.to(in_dtype)


def _ln(x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)


def _linear(sd, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _heads(x: Tensor, B: int, H: int, dh: int) -> Tensor:
    return x.view(B, -1, H, dh).transpose(1, 2)


def flux_attention(sd, prefix: str, cfg: FluxConfig, x: Tensor, enc: Optional[Tensor],
                   rope: Tuple[Tensor, Tensor]):
    """FluxAttnProcessor2_0.__call__ (attention_processor.py:1979-2060)."""
    H, dh = cfg.num_attention_heads, cfg.attention_head_dim
    B = x.shape[0] if enc is None else enc.shape[0]
    q = _heads(_linear(sd, prefix + "to_q", x), B, H, dh)
    k = _heads(_linear(sd, prefix + "to_k", x), B, H, dh)
    v = _heads(_linear(sd, prefix + "to_v", x), B, H, dh)
    q = rms_norm(q, sd[prefix + "norm_q.weight"])
    k = rms_norm(k, sd[prefix + "norm_k.weight"])
    if enc is not None:
        eq = _heads(_linear(sd, prefix + "add_q_proj", enc), B, H, dh)
        ek = _heads(_linear(sd, prefix + "add_k_proj", enc), B, H, dh)
        ev = _heads(_linear(sd, prefix + "add_v_proj", enc), B, H, dh)
        eq = rms_norm(eq, sd[prefix + "norm_added_q.weight"])
        ek = rms_norm(ek, sd[prefix + "norm_added_k.weight"])
        q = torch.cat([eq, q], dim=2)
        k = torch.cat([ek, k], dim=2)
        v = torch.cat([ev, v], dim=2)
    q = apply_rotary_emb(q, *rope)
    k = apply_rotary_emb(k, *rope)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, H * dh)
    o = o.to(q.dtype)
    if enc is not None:
        T = enc.shape[1]
        eo, o = o[:, :T], o[:, T:]
        o = _linear(sd, prefix + "to_out.0", o)
        eo = _linear(sd, prefix + "to_add_out", eo)
        return o, eo
    return o


def _feed_forward(sd, prefix: str, x: Tensor) -> Tensor:
    """FeedForward(activation_fn='gelu-approximate') (attention.py:1185-1243; activations.py:82-90)."""
    h = _linear(sd, prefix + "net.0.proj", x)
    h = F.gelu(h, approximate="tanh")
    return _linear(sd, prefix + "net.2", h)


def double_block(sd, i: int, cfg: FluxConfig, x: Tensor, enc: Tensor, temb: Tensor, rope):
    """FluxTransformerBlock.forward (transformer_flux.py:794-841)."""
    p = f"transformer_blocks.{i}."
    emb = _linear(sd, p + "norm1.linear", F.silu(temb))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
    nx = _ln(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    emb_c = _linear(sd, p + "norm1_context.linear", F.silu(temb))
    c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = emb_c.chunk(6, dim=1)
    nc = _ln(enc) * (1 + c_scale_msa[:, None]) + c_shift_msa[:, None]

    attn_x, attn_c = flux_attention(sd, p + "attn.", cfg, nx, nc, rope)

    attn_x = gate_msa.unsqueeze(1) * attn_x
    x = x + attn_x
    nx = _ln(x)
    nx = nx * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    ff = _feed_forward(sd, p + "ff.", nx)
    ff = gate_mlp.unsqueeze(1) * ff
    x = x + ff

    attn_c = c_gate_msa.unsqueeze(1) * attn_c
    enc = enc + attn_c
    nc = _ln(enc)
    nc = nc * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
    ffc = _feed_forward(sd, p + "ff_context.", nc)
    enc = enc + c_gate_mlp.unsqueeze(1) * ffc
    return enc, x


def single_block(sd, i: int, cfg: FluxConfig, h: Tensor, temb: Tensor, rope) -> Tensor:
    """FluxSingleTransformerBlock.forward (transformer_flux.py:715-739)."""
    p = f"single_transformer_blocks.{i}."
    residual = h
    emb = _linear(sd, p + "norm.linear", F.silu(temb))
    shift, scale, gate = emb.chunk(3, dim=1)
    n = _ln(h) * (1 + scale[:, None]) + shift[:, None]
    mlp = F.gelu(_linear(sd, p + "proj_mlp", n), approximate="tanh")
    attn = flux_attention(sd, p + "attn.", cfg, n, None, rope)
    h = torch.cat([attn, mlp], dim=2)
    gate = gate.unsqueeze(1)
    h = gate * _linear(sd, p + "proj_out", h)
    return residual + h


def flux_forward(sd: Dict[str, Tensor], cfg: FluxConfig, hidden_states: Tensor, encoder_hidden_states: Tensor,
                 pooled_projections: Tensor, timestep: Tensor, img_ids: Tensor, txt_ids: Tensor,
                 guidance: Optional[Tensor], taps: Optional[dict] = None) -> Tensor:
    """FluxTransformer2DModel.forward (transformer_flux.py:1028-1212). Returns `sample` [B,S,out]."""
    h = _linear(sd, "x_embedder", hidden_states)
    timestep = timestep.to(h.dtype) * 1000
    if guidance is not None:
        guidance = guidance.to(h.dtype) * 1000
    temb = time_text_embed(sd, cfg, timestep, guidance, pooled_projections)
    enc = _linear(sd, "context_embedder", encoder_hidden_states)
    if txt_ids.ndim == 3:
        txt_ids = txt_ids[0]
    if img_ids.ndim == 3:
        img_ids = img_ids[0]
    ids = torch.cat((txt_ids, img_ids), dim=0)
    rope = flux_pos_embed(ids, cfg.axes_dims_rope)
    if taps is not None:
        taps["temb"] = temb
        taps["rope_cos"], taps["rope_sin"] = rope
        taps["x_embed"] = h
        taps["ctx_embed"] = enc
    for i in range(cfg.num_layers):
        enc, h = double_block(sd, i, cfg, h, enc, temb, rope)
        if taps is not None:
            taps[f"double.{i}.enc"] = enc
            taps[f"double.{i}.x"] = h
    h = torch.cat([enc, h], dim=1)
    for i in range(cfg.num_single_layers):
        h = single_block(sd, i, cfg, h, temb, rope)
        if taps is not None:
            taps[f"single.{i}"] = h
    h = h[:, enc.shape[1]:, ...]
    emb = _linear(sd, "norm_out.linear", F.silu(temb).to(h.dtype))
    scale, shift = torch.chunk(emb, 2, dim=1)
    h = _ln(h) * (1 + scale)[:, None, :] + shift[:, None, :]
    return _linear(sd, "proj_out", h)


# --------------------------------------------------------------------------------------
# scheduler + pipeline glue
# --------------------------------------------------------------------------------------
def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096,
                    base_shift: float = 0.5, max_shift: float = 1.15) -> float:
    """pipeline_flux_fill.py:1248-1258 evaluated with the scheduler config defaults the pipeline reads (:2053-2056)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def euler_set_timesteps(num_inference_steps: int, image_seq_len: int, use_dynamic_shifting: bool = True,
                        shift: float = 1.0, num_train_timesteps: int = 1000) -> Tuple[Tensor, Tensor]:
    """`sigmas=np.linspace(1, 1/n, n)` (pipeline :2049) through set_timesteps (scheduler :184-241).

    Returns (sigmas[n+1] fp32 with trailing 0, timesteps[n] fp32).
    """
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    sig = np.array(sig).astype(np.float32)
    if use_dynamic_shifting:
        mu = calculate_shift(image_seq_len)
        sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)
    else:
        sig = shift * sig / (1 + (shift - 1) * sig)
    sigmas = torch.from_numpy(sig).to(dtype=torch.float32)
    timesteps = sigmas * num_train_timesteps
    sigmas = torch.cat([sigmas, torch.zeros(1)])
    return sigmas, timesteps


def euler_step(model_output: Tensor, sigma: Tensor, sigma_next: Tensor, sample: Tensor) -> Tensor:
    """FlowMatchEulerDiscreteScheduler.step (scheduler :322-330): 0-dim fp32 dt times bf16 v is bf16."""
    sample = sample.to(torch.float32)
    prev = sample + (sigma_next - sigma) * model_output
    return prev.to(model_output.dtype)


def overshoot_set_timesteps(num_inference_steps: int, image_seq_len: int, num_train_timesteps: int = 1000) -> Tuple[Tensor, Tensor]:
    """StochasticRFOvershotDiscreteScheduler.set_timesteps with the pipeline's float64 `sigmas=np.linspace(1, 1/n, n)`
    (scheduling_stochastic_rf_discrete_overshot.py:183-222): unlike the Euler scheduler the shift runs in float64."""
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    mu = calculate_shift(image_seq_len)
    sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)
    sigmas = torch.from_numpy(sig).to(dtype=torch.float32)
    timesteps = sigmas * num_train_timesteps
    return torch.cat([sigmas, torch.zeros(1)]), timesteps


def overshoot_step(model_output: Tensor, sigma: Tensor, sigma_next: Tensor, sample: Tensor, noise: Tensor, c: float = 2.0):
    """StochasticRFOvershotDiscreteScheduler.step, attn_map=None branch, overshot_func(t, dt) = t + dt, c = 2.0 as TextFlux
    configures it (scheduling_stochastic_rf_discrete_overshot.py:300-366; run_inference.py:84-88).  `noise` is the
    randn_tensor draw of :351-355 (fp32, sample's shape).  Returns (prev_sample, predicted_x1)."""
    sample = sample.to(torch.float32)
    t = 1 - sigma
    step_size = sigma - sigma_next
    t_next = min(t + step_size, 1)
    step_size_overshoot = step_size * c
    t_overshoot = min(t_next + step_size_overshoot, 1)
    sample_overshoot = sample + (t_overshoot - t) * (-model_output)
    a = t_next / t_overshoot
    b = ((1 - t_next) ** 2 - (a - t_next) ** 2) ** (0.5)
    prev_sample = sample_overshoot * a + noise * b
    prev_sample = prev_sample.to(model_output.dtype)
    predicted_x1 = sample - sigma * model_output
    return prev_sample, predicted_x1


def pack_latents(latents: Tensor) -> Tensor:
    """pipeline_flux_fill.py:1743-1748."""
    B, C, h, w = latents.shape
    latents = latents.view(B, C, h // 2, 2, w // 2, 2)
    latents = latents.permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(B, (h // 2) * (w // 2), C * 4)


def unpack_latents(latents: Tensor, height: int, width: int, vae_scale_factor: int = 8) -> Tensor:
    """pipeline_flux_fill.py:1752-1765 (height/width in pixels)."""
    B, _, ch = latents.shape
    height = 2 * (int(height) // (vae_scale_factor * 2))
    width = 2 * (int(width) // (vae_scale_factor * 2))
    latents = latents.view(B, height // 2, width // 2, ch // 4, 2, 2)
    latents = latents.permute(0, 3, 1, 4, 2, 5)
    return latents.reshape(B, ch // 4, height, width)


def prepare_latent_image_ids(h2: int, w2: int, dtype=torch.bfloat16) -> Tensor:
    """pipeline_flux_fill.py:1728-1739; h2,w2 = packed token grid (latent h//2, w//2)."""
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h2)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w2)[None, :]
    return ids.reshape(h2 * w2, 3).to(dtype=dtype)


def pack_mask(mask: Tensor, vae_scale_factor: int = 8) -> Tensor:
    """pipeline_flux_fill.py:1563-1580: [B,1,H,W] pixel mask -> [B, S, 256] packed mask channels."""
    B = mask.shape[0]
    H, W = mask.shape[-2:]
    h, w = H // vae_scale_factor, W // vae_scale_factor
    m = mask[:, 0, :, :]
    m = m.view(B, h, vae_scale_factor, w, vae_scale_factor)
    m = m.permute(0, 2, 4, 1, 3)
    m = m.reshape(B, vae_scale_factor * vae_scale_factor, h, w)
    return pack_latents(m)


def normalize_vae_latents(latents: Tensor, shift_factor: float, scaling_factor: float) -> Tensor:
    """pipeline_flux_fill.py:1536: `(masked_image_latents - shift_factor) * scaling_factor`, each op in the tensor dtype."""
    return (latents - shift_factor) * scaling_factor


def denormalize_vae_latents(latents: Tensor, shift_factor: float, scaling_factor: float) -> Tensor:
    """pipeline_flux_fill.py:2127: `(latents / scaling_factor) + shift_factor` (the input of vae.decode)."""
    return (latents / scaling_factor) + shift_factor


def prepare_mask_latents(mask: Tensor, masked_image_latents: Tensor, batch_size: int, num_channels_latents: int,
                         num_images_per_prompt: int, height: int, width: int, dtype, shift_factor: float,
                         scaling_factor: float, vae_scale_factor: int = 8) -> Tuple[Tensor, Tensor]:
    """pipeline_flux_fill.py:1505-1583 in its `masked_image.shape[1] == num_channels_latents` branch (no VAE call):
    normalise, cast, duplicate per prompt, pack the latents; pixel-unshuffle + pack the mask.  height/width in pixels."""
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    mil = normalize_vae_latents(masked_image_latents, shift_factor, scaling_factor).to(dtype=dtype)
    batch_size = batch_size * num_images_per_prompt
    if mask.shape[0] < batch_size:
        if batch_size % mask.shape[0] != 0:
            raise ValueError("The passed mask and the required batch size don't match.")
        mask = mask.repeat(batch_size // mask.shape[0], 1, 1, 1)
    if mil.shape[0] < batch_size:
        if batch_size % mil.shape[0] != 0:
            raise ValueError("The passed images and the required batch size don't match.")
        mil = mil.repeat(batch_size // mil.shape[0], 1, 1, 1)
    assert mil.shape[1] == num_channels_latents and tuple(mil.shape[2:]) == (h, w)
    mil = pack_latents(mil)
    return pack_mask(mask, vae_scale_factor).to(dtype=dtype), mil


def denoise_loop(sd, cfg: FluxConfig, latents: Tensor, cond: Tensor, prompt_embeds: Tensor, pooled: Tensor,
                 txt_ids: Tensor, img_ids: Tensor, guidance_scale: float, num_inference_steps: int,
                 record: Optional[list] = None) -> Tensor:
    """The hot loop of FluxFillPipeline.__call__ (pipeline_flux_fill.py:2049-2119), callbacks omitted."""
    sigmas, timesteps = euler_set_timesteps(num_inference_steps, latents.shape[1])
    guidance = torch.full([1], guidance_scale, dtype=torch.float32).expand(latents.shape[0]) \
        if cfg.guidance_embeds else None
    for i, t in enumerate(timesteps):
        timestep = t.expand(latents.shape[0]).to(latents.dtype)
        noise_pred = flux_forward(sd, cfg, torch.cat((latents, cond), dim=2), prompt_embeds, pooled,
                                  timestep / 1000, img_ids, txt_ids, guidance)
        latents = euler_step(noise_pred, sigmas[i], sigmas[i + 1], latents)
        if record is not None:
            record.append((noise_pred, latents))
    return latents


# --------------------------------------------------------------------------------------
# synthetic inputs and FLOP model (§8a)
# --------------------------------------------------------------------------------------
def synthetic_inputs(cfg: FluxConfig, h2: int, w2: int, T: int, batch: int = 1, seed0: int = 1000,
                     dtype: torch.dtype = torch.bfloat16) -> dict:
    """Per-sample seeded latents / cond / prompt embeds; mask: 0 on the glyph (top) half, 1 on the scene half."""
    S = h2 * w2
    lat, cond, pe, pp = [], [], [], []
    for b in range(batch):
        g = torch.Generator().manual_seed(seed0 + b)
        lat.append(torch.randn(S, cfg.out_channels, generator=g))
        mil = torch.randn(S, cfg.out_channels, generator=g)
        mask = torch.zeros(h2, w2, cfg.in_channels - 2 * cfg.out_channels)
        mask[h2 // 2:, :, :] = 1.0
        cond.append(torch.cat([mil, mask.reshape(S, -1)], dim=1))
        pe.append(torch.randn(T, cfg.joint_attention_dim, generator=g))
        pp.append(torch.randn(cfg.pooled_projection_dim, generator=g))
    return dict(
        latents=torch.stack(lat).to(dtype), cond=torch.stack(cond).to(dtype),
        prompt_embeds=torch.stack(pe).to(dtype), pooled=torch.stack(pp).to(dtype),
        txt_ids=torch.zeros(T, 3).to(dtype), img_ids=prepare_latent_image_ids(h2, w2, dtype),
    )


def flops_per_step(cfg: FluxConfig, S: int, T: int) -> float:
    """Algorithmic FLOPs per sample per denoising step, SURVEY.md §8a (== FlopCounterMode on the reference)."""
    D, N = cfg.inner_dim, S + T
    L = cfg.num_layers + cfg.num_single_layers
    f = L * (24 * N * D * D + 4 * N * N * D)
    f += 2 * S * cfg.in_channels * D + 2 * T * cfg.joint_attention_dim * D + 2 * S * D * cfg.out_channels
    f += cfg.num_layers * 2 * (2 * D * 6 * D) + cfg.num_single_layers * (2 * D * 3 * D) + 2 * D * 2 * D
    n_mlp = 3 if cfg.guidance_embeds else 2
    f += n_mlp * (2 * 256 * D + 2 * D * D) - 2 * 256 * D + 2 * cfg.pooled_projection_dim * D
    return float(f)
