"""CPU/torch restatement of the reference's AutoencoderKL encode / decode.

TEST INFRASTRUCTURE ONLY: imported by tests/, oracle/make_golden_vae.py and nothing under textflux_b200/.
Pinned bit-exactly (torch.equal, fp32 and bf16) to the reference's own modules run in the build container:
tests/golden/vae_*.pt (written by oracle/make_golden_vae.py), checked by tests/test_oracle_golden.py::test_vae_*.

Reference files (D = /root/reference/diffusers/src/diffusers/models):
  D/autoencoders/autoencoder_kl.py:240-324   AutoencoderKL._encode / encode / _decode / decode (no quant convs for FLUX)
  D/autoencoders/vae.py:54-193               Encoder          :195-343  Decoder        :780-803 DiagonalGaussianDistribution
  D/unets/unet_2d_blocks.py                  DownEncoderBlock2D, UpDecoderBlock2D, UNetMidBlock2D (resnet, attention, resnet)
  D/resnet.py:189-375                        ResnetBlock2D (temb = None, output_scale_factor = 1, 1x1 conv_shortcut when Cin != Cout)
  D/downsampling.py:132-149                  Downsample2D: F.pad (0,1,0,1) + 3x3 stride-2 conv
  D/upsampling.py:142-192                    Upsample2D: F.interpolate(scale_factor=2, nearest) + 3x3 conv
  D/attention_processor.py:2790-2881         AttnProcessor2_0 on the mid block: GroupNorm, one head of C channels, residual
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 16
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    mid_block_add_attention: bool = True
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159

    def to_dict(self) -> dict:
        return asdict(self)

    def reference_kwargs(self) -> dict:
        """Arguments of the reference AutoencoderKL for this config (FLUX.1's VAE: no quant / post-quant convolution)."""
        n = len(self.block_out_channels)
        d = self.to_dict()
        d.update(down_block_types=("DownEncoderBlock2D",) * n, up_block_types=("UpDecoderBlock2D",) * n,
                 block_out_channels=tuple(self.block_out_channels), act_fn="silu", use_quant_conv=False, use_post_quant_conv=False,
                 force_upcast=False, sample_size=64)
        return d


FLUX_VAE = VaeConfig()
SMALL_VAE = VaeConfig(block_out_channels=(64, 128), layers_per_block=1)


def state_dict_spec(cfg: VaeConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) of every tensor of AutoencoderKL.state_dict(); kind: w (conv / linear weight), b (bias), g (norm weight)."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(n, o, i, k=3):
        out.append((n + ".weight", (o, i, k, k), "w"))
        out.append((n + ".bias", (o,), "b"))

    def norm(n, c):
        out.append((n + ".weight", (c,), "g"))
        out.append((n + ".bias", (c,), "b"))

    def lin(n, o, i):
        out.append((n + ".weight", (o, i), "w"))
        out.append((n + ".bias", (o,), "b"))

    def resnet(n, i, o):
        norm(n + ".norm1", i); conv(n + ".conv1", o, i); norm(n + ".norm2", o); conv(n + ".conv2", o, o)
        if i != o:
            conv(n + ".conv_shortcut", o, i, 1)

    def mid(n, c):
        resnet(n + ".resnets.0", c, c)
        if cfg.mid_block_add_attention:
            a = n + ".attentions.0"
            norm(a + ".group_norm", c)
            for m in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(a + "." + m, c, c)
        resnet(n + ".resnets.1", c, c)

    ch = cfg.block_out_channels
    conv("encoder.conv_in", ch[0], cfg.in_channels)
    c = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i + 1 < len(ch):
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c)
    mid("encoder.mid_block", c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", 2 * cfg.latent_channels, c)
    conv("decoder.conv_in", ch[-1], cfg.latent_channels)
    c = ch[-1]
    mid("decoder.mid_block", c)
    for i, co in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i + 1 < len(ch):
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c)
    norm("decoder.conv_norm_out", c)
    conv("decoder.conv_out", cfg.out_channels, c)
    return out


def init_state_dict(cfg: VaeConfig, seed: int = 77, dtype=torch.float32, device="cpu") -> Dict[str, Tensor]:
    """Synthetic weights, one seeded generator per tensor: conv / linear weights N(0, 1/fan_in) (activations stay O(1) through the
    ~60 layers), biases N(0, 0.02^2), norm weights 1 + N(0, 0.1^2)."""
    sd = {}
    for idx, (name, shape, kind) in enumerate(state_dict_spec(cfg)):
        g = torch.Generator(device=device).manual_seed(seed * 100003 + idx)
        t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = t * (1.0 / fan_in) ** 0.5
        elif kind == "b":
            t = t * 0.02
        else:
            t = 1.0 + 0.1 * t
        sd[name] = t.to(dtype)
    return sd


# ------------------------------------------------------------------------------------------------------------------ layers
def _conv(sd, n, x, stride=1, padding=1):
    return F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"], stride=stride, padding=padding)


def _gn(sd, n, x, groups):
    return F.group_norm(x, groups, sd[n + ".weight"], sd[n + ".bias"], eps=1e-6)


def resnet_block(sd, n: str, x: Tensor, groups: int) -> Tensor:
    """ResnetBlock2D.forward with temb = None (resnet.py:320-375)."""
    h = F.silu(_gn(sd, n + ".norm1", x, groups))
    h = _conv(sd, n + ".conv1", h)
    h = F.silu(_gn(sd, n + ".norm2", h, groups))
    h = _conv(sd, n + ".conv2", h)
    if (n + ".conv_shortcut.weight") in sd:
        x = _conv(sd, n + ".conv_shortcut", x, padding=0)
    return (x + h) / 1.0


def mid_attention(sd, n: str, x: Tensor, groups: int) -> Tensor:
    """Attention(heads=1, dim_head=C, residual_connection=True, norm_num_groups) through AttnProcessor2_0 (attention_processor.py:2790-2881)."""
    B, C, H, W = x.shape
    res = x
    h = x.view(B, C, H * W).transpose(1, 2)
    h = F.group_norm(h.transpose(1, 2), groups, sd[n + ".group_norm.weight"], sd[n + ".group_norm.bias"], eps=1e-6).transpose(1, 2)
    q = F.linear(h, sd[n + ".to_q.weight"], sd[n + ".to_q.bias"])
    k = F.linear(h, sd[n + ".to_k.weight"], sd[n + ".to_k.bias"])
    v = F.linear(h, sd[n + ".to_v.weight"], sd[n + ".to_v.bias"])
    q, k, v = (t.view(B, -1, 1, C).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, -1, C).to(q.dtype)
    o = F.linear(o, sd[n + ".to_out.0.weight"], sd[n + ".to_out.0.bias"])
    o = o.transpose(-1, -2).reshape(B, C, H, W)
    return (o + res) / 1.0


def mid_block(sd, n: str, x: Tensor, cfg: VaeConfig) -> Tensor:
    x = resnet_block(sd, n + ".resnets.0", x, cfg.norm_num_groups)
    if cfg.mid_block_add_attention:
        x = mid_attention(sd, n + ".attentions.0", x, cfg.norm_num_groups)
    return resnet_block(sd, n + ".resnets.1", x, cfg.norm_num_groups)


@torch.no_grad()
def encode_moments(sd, cfg: VaeConfig, x: Tensor) -> Tensor:
    """AutoencoderKL._encode (autoencoder_kl.py:240-261) -> Encoder.forward (vae.py:140-193): [B,3,H,W] -> [B, 2*latent, H/f, W/f]."""
    g = cfg.norm_num_groups
    h = _conv(sd, "encoder.conv_in", x)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g)
        if i + 1 < n:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)  # Downsample2D(padding=0), downsampling.py:141-147
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = mid_block(sd, "encoder.mid_block", h, cfg)
    h = F.silu(_gn(sd, "encoder.conv_norm_out", h, g))
    return _conv(sd, "encoder.conv_out", h)


@torch.no_grad()
def decode(sd, cfg: VaeConfig, z: Tensor) -> Tensor:
    """AutoencoderKL._decode (autoencoder_kl.py:291-304) -> Decoder.forward (vae.py:284-343): [B, latent, h, w] -> [B,3,h*f,w*f]."""
    g = cfg.norm_num_groups
    h = _conv(sd, "decoder.conv_in", z)
    h = mid_block(sd, "decoder.mid_block", h, cfg)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, g)
        if i + 1 < n:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # Upsample2D, upsampling.py:176-188
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(_gn(sd, "decoder.conv_norm_out", h, g))
    return _conv(sd, "decoder.conv_out", h)


def gaussian_sample(moments: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution(moments).sample() given the N(0,1) draw (vae.py:780-803)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * noise
