"""The reference arm: the UNMODIFIED reference (yyyyyxie/textflux's vendored diffusers 0.32.0.dev0) installed under
`baseline/_ref` and driven through its own public API -- `FluxTransformer2DModel.forward`,
`FlowMatchEulerDiscreteScheduler.step`, `FluxFillPipeline.__call__`.  None of the engine's code is on this path.

Install (done by `__graft_entry__.build()` where /root/reference exists; the result is git-ignored but travels to the GPU
box with the repo snapshot):

    cp -r /root/reference/diffusers /tmp/refbuild/diffusers          # the source tree is read-only, the build writes egg-info
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/refbuild/diffusers

Used by `bench.py --impl reference` (CPU, host cores), by bench.py's `gpu_eager_baseline` / `parity` legs (the same
modules on CUDA tensors: what the reference dispatches is aten::addmm + F.scaled_dot_product_attention,
attention_processor.py:2039-2041) and by tests/test_gpu_dropin.py (the unmodified pipeline with the engine attached).
"""
from __future__ import annotations

import os
import sys
from typing import Callable, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "diffusers"))


def import_reference():
    """The reference's `diffusers` package from baseline/_ref (transformers >= 5 needs one shim: the pinned 4.43 still
    exported FLAX_WEIGHTS_NAME, pipelines/pipeline_loading_utils.py:49)."""
    if not available():
        raise RuntimeError(f"reference not installed under {REF_DIR} (see baseline/reference_arm.py)")
    import transformers.utils as tu
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import diffusers
    if not os.path.abspath(diffusers.__file__).startswith(REF_DIR):
        raise RuntimeError(f"`diffusers` resolved to {diffusers.__file__}, not to the reference under {REF_DIR}")
    return diffusers


FLUX_FILL_12B = dict(patch_size=1, in_channels=384, out_channels=64, num_layers=19, num_single_layers=38,
                     attention_head_dim=128, num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768,
                     guidance_embeds=True, axes_dims_rope=(16, 56, 56))


def build_transformer(cfg: dict, get: Callable[[str], "torch.Tensor"], device, dtype):
    """Reference FluxTransformer2DModel with every parameter filled from `get(name)` (names = its own state-dict keys)."""
    import torch
    d = import_reference()
    with torch.device("meta"):
        m = d.FluxTransformer2DModel(**cfg)
    m = m.to_empty(device=device).to(dtype)
    with torch.no_grad():
        for name, p in m.state_dict().items():
            p.copy_(get(name).to(device=device, dtype=dtype))
    return m.eval()


def build_transformer_aliased(cfg: dict, device, dtype, seed: int = 7):
    """The reference model at full depth whose 19 double blocks are ONE block object and whose 38 single blocks are ONE
    block object (same code executed 19 / 38 times, 0.5 GB of weights instead of 23.8 GB): skips a minute of 12B random
    init on the host.  Timing-only -- the output is not a FLUX forward."""
    import torch
    d = import_reference()
    small = dict(cfg, num_layers=1, num_single_layers=1)
    torch.manual_seed(seed)
    m = d.FluxTransformer2DModel(**small).to(device=device, dtype=dtype)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0.0, 0.02)
    m.transformer_blocks = torch.nn.ModuleList([m.transformer_blocks[0]] * cfg["num_layers"])
    m.single_transformer_blocks = torch.nn.ModuleList([m.single_transformer_blocks[0]] * cfg["num_single_layers"])
    return m.eval()


def build_scheduler():
    d = import_reference()
    return d.FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15)


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    import_reference()
    from diffusers.pipelines.flux.pipeline_flux_fill import calculate_shift as cs
    return cs(image_seq_len, base_seq_len, max_seq_len, base_shift, max_shift)


def reference_step_fn(model, scheduler):
    """One iteration of the reference's denoising loop, pipeline_flux_fill.py:2077-2098, verbatim call sequence."""
    import torch

    @torch.no_grad()
    def step(i, latents, cond, prompt_embeds, pooled, guidance, txt_ids, img_ids):
        t = scheduler.timesteps[i]
        timestep = t.expand(latents.shape[0]).to(latents.dtype)
        noise_pred = model(hidden_states=torch.cat((latents, cond), dim=2), timestep=timestep / 1000, guidance=guidance,
                           pooled_projections=pooled, encoder_hidden_states=prompt_embeds, txt_ids=txt_ids,
                           img_ids=img_ids, joint_attention_kwargs=None, return_dict=False)[0]
        return scheduler.step(noise_pred, t, latents, return_dict=False)[0], noise_pred

    return step
