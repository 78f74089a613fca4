"""B200AutoencoderKL: drop-in for the reference's AutoencoderKL on `pipe.vae`.

What FluxFillPipeline touches (pipelines/flux/pipeline_flux_fill.py): `vae.encode(x).latent_dist.sample(generator)` through
`retrieve_latents` (:1322-1332, called :1528), `vae.decode(z, return_dict=False)[0]` (:2128), `vae.config.{block_out_channels,
latent_channels, scaling_factor, shift_factor}` (:1393-1400, :1530, :2011, :2127), `vae.dtype`.  Everything on the device runs
through the C ABI (tfx_vae_*, include/textflux_b200.h); there is no CPU or torch fallback.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Callable, Dict, Optional, Union

import torch

from . import _lib
from .engine import FrozenConfig

Tensor = torch.Tensor


def _randn_tensor(shape, generator, device, dtype) -> Tensor:
    """utils/torch_utils.py randn_tensor: CPU generators draw on the CPU (device-independent noise), lists seed per sample."""
    rand_device = device
    if generator is not None:
        g0 = generator[0] if isinstance(generator, list) else generator
        if g0.device.type != device.type and g0.device.type == "cpu":
            rand_device = torch.device("cpu")
        elif g0.device.type != device.type and g0.device.type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {g0.device.type}.")
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        return torch.cat([torch.randn(one, generator=generator[i], device=rand_device, dtype=dtype) for i in range(shape[0])], dim=0).to(device)
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)


class B200DiagonalGaussian:
    """DiagonalGaussianDistribution (models/autoencoders/vae.py:780-841) over moments produced by tfx_vae_encode."""

    def __init__(self, parameters: Tensor, lib):
        self.parameters = parameters  # [B, 2L, h, w] bf16: mean ; logvar
        self._lib = lib
        self.deterministic = False

    @property
    def mean(self) -> Tensor:
        return torch.chunk(self.parameters, 2, dim=1)[0]

    @property
    def logvar(self) -> Tensor:
        return torch.clamp(torch.chunk(self.parameters, 2, dim=1)[1], -30.0, 20.0)

    def mode(self) -> Tensor:
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None) -> Tensor:
        B, L2, h, w = self.parameters.shape
        L = L2 // 2
        noise = _randn_tensor((B, L, h, w), generator, self.parameters.device, self.parameters.dtype).contiguous()
        out = torch.empty_like(noise)
        params = self.parameters.contiguous()  # [B, 2L, h, w] row-major for the kernel (a caller-built tensor may be channels-last)
        with torch.cuda.device(self.parameters.device):
            _lib.check(self._lib.tfx_op_gaussian_sample(params.data_ptr(), noise.data_ptr(), out.data_ptr(), B, L, h * w,
                                                        torch.cuda.current_stream().cuda_stream))
        return out


def pack_vae_weights(get: Callable[[str], Tensor], names, device, dtype=torch.bfloat16) -> Dict[str, Tensor]:
    """Reference AutoencoderKL tensors -> the layouts tfx_vae_set_weight documents: 3x3 convolutions [Cout, 9 * Cin64] with column
    (ky * 3 + kx) * Cin64 + c (input channels zero-padded to a multiple of 64), 1x1 convolutions / Linear [Cout, Cin], vectors [1, C]."""
    P: Dict[str, Tensor] = {}
    for n in names:
        t = get(n).to(device=device, dtype=torch.float32)
        if t.ndim == 4 and t.shape[2] == 3 and t.shape[3] == 3:
            co, ci = t.shape[:2]
            cip = (ci + 63) // 64 * 64
            w = torch.zeros(co, 3, 3, cip, device=device, dtype=torch.float32)
            w[..., :ci] = t.permute(0, 2, 3, 1)
            t = w.reshape(co, 9 * cip)
        elif t.ndim == 4 and t.shape[2] == 1 and t.shape[3] == 1:
            t = t.reshape(t.shape[0], t.shape[1])
        elif t.ndim == 1:
            t = t.reshape(1, -1)
        elif t.ndim != 2:
            raise ValueError(f"textflux_b200: VAE tensor '{n}' has unsupported shape {tuple(t.shape)}")
        P[n] = t.to(dtype).contiguous()
    return P


def vae_reference_names(cfg) -> list:
    """(name, shape) of every tensor AutoencoderKL.state_dict() holds for `cfg` (dict with in_channels, out_channels, latent_channels,
    block_out_channels, layers_per_block, mid_block_add_attention): what a checkpoint must provide, and the shapes synthetic weights take."""
    out = []

    def conv(n, o, i, k=3):
        out.extend([(n + ".weight", (o, i, k, k)), (n + ".bias", (o,))])

    def vec(n, c):
        out.extend([(n + ".weight", (c,)), (n + ".bias", (c,))])

    def resnet(n, i, o):
        vec(n + ".norm1", i); conv(n + ".conv1", o, i); vec(n + ".norm2", o); conv(n + ".conv2", o, o)
        if i != o:
            conv(n + ".conv_shortcut", o, i, 1)

    def mid(n, c):
        resnet(n + ".resnets.0", c, c)
        if cfg.get("mid_block_add_attention", True):
            vec(n + ".attentions.0.group_norm", c)
            for m in ("to_q", "to_k", "to_v", "to_out.0"):
                out.extend([(f"{n}.attentions.0.{m}.weight", (c, c)), (f"{n}.attentions.0.{m}.bias", (c,))])
        resnet(n + ".resnets.1", c, c)

    ch = list(cfg["block_out_channels"])
    conv("encoder.conv_in", ch[0], cfg["in_channels"])
    c = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg["layers_per_block"]):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i + 1 < len(ch):
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c)
    mid("encoder.mid_block", c)
    vec("encoder.conv_norm_out", c)
    conv("encoder.conv_out", 2 * cfg["latent_channels"], c)
    conv("decoder.conv_in", ch[-1], cfg["latent_channels"])
    c = ch[-1]
    mid("decoder.mid_block", c)
    for i, co in enumerate(reversed(ch)):
        for j in range(cfg["layers_per_block"] + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", c, co)
            c = co
        if i + 1 < len(ch):
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c)
    vec("decoder.conv_norm_out", c)
    conv("decoder.conv_out", cfg["out_channels"], c)
    return out


def synthetic_state(names, seed: int, device, dtype=torch.bfloat16) -> Dict[str, Tensor]:
    """Random tensors of the given (name, shape) list, one seeded generator each (bench / smoke: no checkpoints offline): matrices and
    convolutions N(0, 1 / fan_in), 1-D weights of norms 1 + N(0, 0.1^2), other vectors N(0, 0.02^2)."""
    sd = {}
    for idx, (name, shape) in enumerate(names):
        g = torch.Generator(device=device).manual_seed(seed * 100003 + idx)
        t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if len(shape) >= 2:
            fan = 1
            for d in shape[1:]:
                fan *= d
            t = t * (1.0 / fan) ** 0.5 if "embed" not in name and "shared" not in name else t
        elif name.endswith(".weight") and ("norm" in name or "ln" in name):
            t = 1.0 + 0.1 * t
        else:
            t = 0.02 * t
        sd[name] = t.to(dtype)
    return sd


FLUX_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, mid_block_add_attention=True, scaling_factor=0.3611, shift_factor=0.1159,
                       use_quant_conv=False, use_post_quant_conv=False, act_fn="silu")


class B200AutoencoderKL(torch.nn.Module):
    """Drop-in for AutoencoderKL (models/autoencoders/autoencoder_kl.py:35-571) restricted to what FLUX's VAE is: DownEncoderBlock2D /
    UpDecoderBlock2D stacks, SiLU, GroupNorm, one-head mid-block attention, no quant / post-quant convolutions, no tiling / slicing."""

    def __init__(self, config, get: Callable[[str], Tensor], names=None, device: Union[str, torch.device] = "cuda"):
        super().__init__()
        self._lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("textflux_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        cfg = dict(config) if isinstance(config, dict) else {k: getattr(config, k) for k in dir(config) if not k.startswith("_") and not callable(getattr(config, k))}
        for key, want in (("use_quant_conv", False), ("use_post_quant_conv", False), ("act_fn", "silu")):
            if key in cfg and cfg[key] != want:
                raise ValueError(f"textflux_b200: AutoencoderKL config {key}={cfg[key]!r} is not supported (FLUX's VAE has {want!r})")
        for key, want in (("down_block_types", "DownEncoderBlock2D"), ("up_block_types", "UpDecoderBlock2D")):
            if key in cfg and any(t != want for t in cfg[key]):
                raise ValueError(f"textflux_b200: only {want} stacks are supported, got {cfg[key]!r}")
        cfg.setdefault("scaling_factor", 0.18215)
        cfg.setdefault("shift_factor", None)
        cfg.setdefault("mid_block_add_attention", True)
        self.config = FrozenConfig(cfg)
        self._dev = dev
        ch = list(cfg["block_out_channels"])
        if len(ch) > 8:
            raise ValueError("textflux_b200: at most 8 VAE blocks")
        tc = _lib.TfxVaeConfig(cfg["in_channels"], cfg["out_channels"], cfg["latent_channels"], len(ch), (C.c_int32 * 8)(*(ch + [0] * (8 - len(ch)))),
                               cfg["layers_per_block"], cfg["norm_num_groups"], int(bool(cfg["mid_block_add_attention"])))
        h = C.c_void_p()
        _lib.check(self._lib.tfx_vae_create(C.byref(tc), dev.index, C.byref(h)))
        self._h = h
        self.register_buffer("_anchor", torch.zeros(1, dtype=torch.bfloat16, device=dev), persistent=False)
        if names is None:
            raise ValueError("names: the reference state-dict keys to load")
        with torch.cuda.device(dev):
            self._weights = pack_vae_weights(get, names, dev)
        for name, t in self._weights.items():
            _lib.check(self._lib.tfx_vae_set_weight(self._h, name.encode(), t.data_ptr(), t.shape[0], t.shape[1]), self._h, vae=True)
        self._f = 2 ** (len(ch) - 1)

    @classmethod
    def from_reference(cls, module: torch.nn.Module, device="cuda") -> "B200AutoencoderKL":
        """`module`: a loaded reference AutoencoderKL (weights + `.config`); it can be freed afterwards."""
        sd = module.state_dict()
        return cls(module.config, sd.__getitem__, names=list(sd.keys()), device=device)

    @classmethod
    def from_state_dict(cls, config, state_dict: Dict[str, Tensor], device="cuda") -> "B200AutoencoderKL":
        return cls(config, state_dict.__getitem__, names=list(state_dict.keys()), device=device)

    # ---- what DiffusionPipeline reads -------------------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self._dev

    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    def to(self, *args, **kwargs):
        dev, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if dev is not None and (torch.device(dev).type != "cuda" or torch.device(dev).index not in (None, self._dev.index)):
            raise RuntimeError(f"textflux_b200: the VAE engine lives on {self._dev}")
        if dtype is not None and dtype != torch.bfloat16:
            raise RuntimeError("textflux_b200: the VAE engine computes in bf16 only")
        return self

    def counter(self, key: str) -> int:
        v = C.c_int64()
        _lib.check(self._lib.tfx_vae_get_counter(self._h, key.encode(), C.byref(v)), self._h, vae=True)
        return v.value

    def __del__(self):
        h, lib = getattr(self, "_h", None), getattr(self, "_lib", None)
        if h and lib:
            lib.tfx_vae_destroy(h)
            self._h = None

    def enable_slicing(self):
        pass  # every image of a batch already runs through the same launches

    disable_slicing = enable_slicing

    def enable_tiling(self, *a, **k):
        raise NotImplementedError("textflux_b200: tiled VAE decoding is not implemented (180 GB of HBM hold a 2048 x 2048 decode untiled)")

    def disable_tiling(self):
        pass

    # ---- the two calls ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: Tensor, return_dict: bool = True):
        """AutoencoderKL.encode (autoencoder_kl.py:263-289): x [B, 3, H, W] (bf16 or fp32) -> posterior over [B, L, H/f, W/f]."""
        if x.ndim != 4 or x.shape[1] != self.config.in_channels:
            raise ValueError(f"expected an image batch [B, {self.config.in_channels}, H, W], got {tuple(x.shape)}")
        if x.device != self._dev:
            raise ValueError(f"input lives on {x.device}, the engine on {self._dev}")
        if x.dtype not in (torch.bfloat16, torch.float32):
            raise ValueError(f"image dtype {x.dtype} unsupported (bf16 or fp32)")
        B, _, H, W = x.shape
        if H % self._f or W % self._f:
            raise ValueError(f"image {H} x {W} is not a multiple of the VAE scale factor {self._f}")
        x = x.contiguous()
        moments = torch.empty(B, 2 * self.config.latent_channels, H // self._f, W // self._f, device=self._dev, dtype=torch.bfloat16)
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.tfx_vae_encode(self._h, x.data_ptr(), int(x.dtype == torch.float32), B, H, W, moments.data_ptr(),
                                                torch.cuda.current_stream(self._dev).cuda_stream), self._h, vae=True)
        post = B200DiagonalGaussian(moments, self._lib)
        return SimpleNamespace(latent_dist=post) if return_dict else (post,)

    @torch.no_grad()
    def decode(self, z: Tensor, return_dict: bool = True, generator=None):
        """AutoencoderKL.decode (autoencoder_kl.py:306-330): z [B, L, h, w] -> image [B, 3, h*f, w*f] bf16."""
        if z.ndim != 4 or z.shape[1] != self.config.latent_channels:
            raise ValueError(f"expected latents [B, {self.config.latent_channels}, h, w], got {tuple(z.shape)}")
        if z.device != self._dev:
            raise ValueError(f"input lives on {z.device}, the engine on {self._dev}")
        z = z.to(torch.bfloat16).contiguous()
        B, _, h, w = z.shape
        img = torch.empty(B, self.config.out_channels, h * self._f, w * self._f, device=self._dev, dtype=torch.bfloat16)
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.tfx_vae_decode(self._h, z.data_ptr(), B, h, w, img.data_ptr(), torch.cuda.current_stream(self._dev).cuda_stream),
                       self._h, vae=True)
        return SimpleNamespace(sample=img) if return_dict else (img,)

    def forward(self, sample: Tensor, sample_posterior: bool = False, return_dict: bool = True, generator=None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z, return_dict=return_dict)
