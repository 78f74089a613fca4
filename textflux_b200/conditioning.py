"""Conditioning glue either side of the denoising loop, on the GPU through the C ABI.

Mirrors the static helpers of the reference pipeline (same names minus the leading underscore, same argument meaning,
same errors), so that `FluxFillPipeline.prepare_latents / prepare_mask_latents / __call__` can call these instead:

    pack_latents(latents, batch_size, num_channels_latents, height, width)      pipeline_flux_fill.py:1743-1748
    unpack_latents(latents, height, width, vae_scale_factor)                    :1752-1765
    prepare_latent_image_ids(batch_size, height, width, device, dtype)          :1728-1739
    prepare_mask_latents(mask, masked_image_latents, ...)                       :1505-1583 (after the VAE encode)
    pack_conditioning(...)      the three packs written straight into one [B, S, 320] `cond` buffer (what :2046 cats)
    unscale_unpack_latents(...) `_unpack_latents` + `latents / scaling_factor + shift_factor` (:2126-2127), one kernel

Everything here is an index permutation of 2-byte elements (plus the two scalar affine maps with the reference's bf16
rounding points); results are bit-identical to the reference's (tests/test_gpu_conditioning.py).  CUDA tensors only:
there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_cuda(t: Tensor, name: str, dtypes=(torch.bfloat16,)) -> Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (textflux_b200 has no CPU path)")
    if t.dtype not in dtypes:
        raise ValueError(f"{name} must be {' or '.join(str(d) for d in dtypes)}, got {t.dtype}")
    return t.contiguous()


def pack_latents(latents: Tensor, batch_size: int, num_channels_latents: int, height: int, width: int,
                 out: Optional[Tensor] = None, channel_offset: int = 0, shift_factor: Optional[float] = None,
                 scaling_factor: Optional[float] = None) -> Tensor:
    """[B, C, height, width] -> [B, (height//2)*(width//2), C*4] (bf16).  `out`/`channel_offset` write into a wider token
    buffer; shift/scaling_factor apply `(x - shift_factor) * scaling_factor` first, as prepare_mask_latents does."""
    lib = _lib.load()
    latents = _check_cuda(latents, "latents", (torch.bfloat16, torch.float32))
    if tuple(latents.shape) != (batch_size, num_channels_latents, height, width):
        raise ValueError(f"latents shape {tuple(latents.shape)} != {(batch_size, num_channels_latents, height, width)}")
    S = (height // 2) * (width // 2)
    if out is None:
        out = torch.empty(batch_size, S, num_channels_latents * 4, device=latents.device, dtype=torch.bfloat16)
    if out.dtype != torch.bfloat16 or not out.is_cuda or out.dim() != 3 or out.shape[0] != batch_size or out.shape[1] != S or out.stride(2) != 1 \
            or out.stride(0) != S * out.stride(1):
        raise ValueError("out must be a bf16 CUDA tensor [B, S, ld] with contiguous tokens")
    affine = shift_factor is not None or scaling_factor is not None
    with torch.cuda.device(latents.device):
        _lib.check(lib.tfx_op_pack_latents(latents.data_ptr(), int(latents.dtype == torch.float32), out.data_ptr(), out.stride(1),
                                           channel_offset, batch_size, num_channels_latents, height, width, int(affine),
                                           float(shift_factor or 0.0), float(1.0 if scaling_factor is None else scaling_factor),
                                           _stream(latents)))
    return out


def unpack_latents(latents: Tensor, height: int, width: int, vae_scale_factor: int, shift_factor: Optional[float] = None,
                   scaling_factor: Optional[float] = None) -> Tensor:
    """[B, S, C*4] -> [B, C, h, w] with h = 2*(height // (vae_scale_factor*2)) (height/width in pixels, like the reference)."""
    lib = _lib.load()
    latents = _check_cuda(latents, "latents")
    batch_size, num_patches, channels = latents.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    if num_patches != (h // 2) * (w // 2) or channels % 4:
        raise ValueError(f"latents [{batch_size}, {num_patches}, {channels}] do not unpack to {h}x{w}")
    out = torch.empty(batch_size, channels // 4, h, w, device=latents.device, dtype=torch.bfloat16)
    affine = shift_factor is not None or scaling_factor is not None
    with torch.cuda.device(latents.device):
        _lib.check(lib.tfx_op_unpack_latents(latents.data_ptr(), latents.stride(1), out.data_ptr(), batch_size, channels // 4, h, w,
                                             int(affine), float(shift_factor or 0.0),
                                             float(1.0 if scaling_factor is None else scaling_factor), _stream(latents)))
    return out


def unscale_unpack_latents(latents: Tensor, height: int, width: int, vae_scale_factor: int, shift_factor: float,
                           scaling_factor: float) -> Tensor:
    """pipeline_flux_fill.py:2126-2127 in one kernel: unpack, then `latents / scaling_factor + shift_factor` (input of vae.decode)."""
    return unpack_latents(latents, height, width, vae_scale_factor, shift_factor, scaling_factor)


def prepare_latent_image_ids(batch_size: int, height: int, width: int, device, dtype) -> Tensor:
    """(0, i, j) per packed token; height/width are the packed grid (latent h//2, w//2).  Tiny: plain torch on the host."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)


def pack_mask(mask: Tensor, height: int, width: int, vae_scale_factor: int = 8, out: Optional[Tensor] = None,
              channel_offset: int = 0) -> Tensor:
    """mask [B, 1, height*vs, width*vs] (bf16 or fp32) -> [B, S, vs*vs*4] bf16 (height/width = latent size)."""
    lib = _lib.load()
    mask = _check_cuda(mask, "mask", (torch.bfloat16, torch.float32))
    B = mask.shape[0]
    if tuple(mask.shape) != (B, 1, height * vae_scale_factor, width * vae_scale_factor):
        raise ValueError(f"mask shape {tuple(mask.shape)} != {(B, 1, height * vae_scale_factor, width * vae_scale_factor)}")
    S = (height // 2) * (width // 2)
    ch = vae_scale_factor * vae_scale_factor * 4
    if out is None:
        out = torch.empty(B, S, ch, device=mask.device, dtype=torch.bfloat16)
    if out.dtype != torch.bfloat16 or not out.is_cuda or out.dim() != 3 or out.shape[0] != B or out.shape[1] != S or out.stride(2) != 1 \
            or out.stride(0) != S * out.stride(1):
        raise ValueError("out must be a bf16 CUDA tensor [B, S, ld] with contiguous tokens")
    with torch.cuda.device(mask.device):
        _lib.check(lib.tfx_op_pack_mask(mask.data_ptr(), int(mask.dtype == torch.float32), out.data_ptr(), out.stride(1), channel_offset,
                                        B, height, width, vae_scale_factor, _stream(mask)))
    return out


def _repeat_to(t: Tensor, batch_size: int, what: str) -> Tensor:
    if t.shape[0] < batch_size:
        if batch_size % t.shape[0] != 0:
            raise ValueError(f"The passed {what} and the required batch size don't match. {what.capitalize()} are supposed to be "
                             f"duplicated to a total batch size of {batch_size}, but {t.shape[0]} {what} were passed. Make sure the "
                             f"number of {what} that you pass is divisible by the total requested batch size.")
        t = t.repeat(batch_size // t.shape[0], 1, 1, 1)
    return t


def prepare_mask_latents(mask: Tensor, masked_image_latents: Tensor, batch_size: int, num_channels_latents: int,
                         num_images_per_prompt: int, height: int, width: int, dtype, device, shift_factor: float,
                         scaling_factor: float, vae_scale_factor: int = 8) -> Tuple[Tensor, Tensor]:
    """FluxFillPipeline.prepare_mask_latents after the VAE encode (`masked_image.shape[1] == num_channels_latents` branch,
    pipeline_flux_fill.py:1527-1583): normalise, duplicate per prompt, pack both.  height/width in pixels.
    Returns (mask [B, S, 256], masked_image_latents [B, S, 64]) in bf16."""
    if dtype != torch.bfloat16:
        raise ValueError("textflux_b200 computes the hot path in bf16")
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    batch_size = batch_size * num_images_per_prompt
    mask = _repeat_to(mask.to(device), batch_size, "mask")
    mil = _repeat_to(masked_image_latents.to(device), batch_size, "images")
    mil_p = pack_latents(mil, batch_size, num_channels_latents, h, w, shift_factor=shift_factor, scaling_factor=scaling_factor)
    mask_p = pack_mask(mask, h, w, vae_scale_factor)
    return mask_p, mil_p


def pack_conditioning(mask: Tensor, masked_image_latents: Tensor, height: int, width: int, shift_factor: float,
                      scaling_factor: float, vae_scale_factor: int = 8) -> Tensor:
    """The tensor the denoising loop concatenates to the latents every step (pipeline_flux_fill.py:2046, 2085):
    cond[b, s] = [masked_image_latents 0:4C | mask 4C:4C+4*vs*vs], written by two kernels into one buffer.
    height/width = latent size (pixels // vae_scale_factor)."""
    B, C = masked_image_latents.shape[:2]
    S = (height // 2) * (width // 2)
    ch = 4 * C + 4 * vae_scale_factor * vae_scale_factor
    cond = torch.empty(B, S, ch, device=masked_image_latents.device, dtype=torch.bfloat16)
    pack_latents(masked_image_latents, B, C, height, width, out=cond, channel_offset=0, shift_factor=shift_factor,
                 scaling_factor=scaling_factor)
    pack_mask(mask, height, width, vae_scale_factor, out=cond, channel_offset=4 * C)
    return cond
