"""textflux_b200: Blackwell-native (sm_100a) FLUX-Fill denoising engine behind TextFlux's FluxFillPipeline surface.

Public surface (mirrors what the reference pipeline touches):
    B200FluxTransformer           drop-in for FluxTransformer2DModel on `pipe.transformer`
    B200FlowMatchEulerScheduler   drop-in for FlowMatchEulerDiscreteScheduler on `pipe.scheduler`
    B200StochasticRFOvershotScheduler   drop-in for StochasticRFOvershotDiscreteScheduler (TextFlux's "overshoot" sampler)
    attach(pipe)                  swap both into a loaded FluxFillPipeline
    B200AutoencoderKL             drop-in for AutoencoderKL on `pipe.vae` (encode / decode)
    B200T5Encoder, B200CLIPTextEncoder   drop-ins for T5EncoderModel / CLIPTextModel on `pipe.text_encoder_2` / `pipe.text_encoder`
    conditioning                  pack/unpack/mask-pack kernels mirroring FluxFillPipeline._pack_latents & co
    loader                        safetensors / LoRA files straight into the packed weight layout (transformer, VAE, prompt encoders:
                                  load_transformer / load_vae / load_text_encoder / load_components)
"""
from .engine import (B200FlowMatchEulerScheduler, B200FluxTransformer, B200StochasticRFOvershotScheduler,  # noqa: F401
                     FrozenConfig, attach, calculate_shift)
from . import conditioning  # noqa: F401
from .vae import B200AutoencoderKL  # noqa: F401
from .text_encoders import B200CLIPTextEncoder, B200T5Encoder  # noqa: F401
from .packer import fold_lora, lora_modules, pack_weights, packed_layout, reference_names, repack_modules, synthetic_getter  # noqa: F401

__all__ = ["B200FluxTransformer", "B200AutoencoderKL", "B200T5Encoder", "B200CLIPTextEncoder", "B200FlowMatchEulerScheduler", "B200StochasticRFOvershotScheduler", "attach", "calculate_shift", "fold_lora",
           "pack_weights", "packed_layout", "repack_modules", "lora_modules", "reference_names", "synthetic_getter", "FrozenConfig"]
