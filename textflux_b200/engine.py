"""Host-side mirror of the reference interface for the denoising hot path.

`B200FluxTransformer` answers the calls FluxFillPipeline makes on `self.transformer`
(pipeline_flux_fill.py:2069 `.config.guidance_embeds`, :2084-2094 `forward(...)`, :1660 `.dtype`;
pipeline_utils.py:478-489 `.device`) and `B200FlowMatchEulerScheduler` those it makes on `self.scheduler`
(:2053-2056 `.config.*`, :1305-1314 `set_timesteps(sigmas=, mu=)`, :2065 `.order`, :2077 `.timesteps`, :2098 `.step`).
Everything below the call is the CUDA library behind include/textflux_b200.h; torch only owns device memory and
the stream.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .packer import pack_weights

Tensor = torch.Tensor


class FrozenConfig(dict):
    """dict with attribute access, like diffusers' FrozenDict (configuration_utils.py:55-62)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is frozen")


def _cfg_dict(cfg) -> dict:
    keys = ("patch_size", "in_channels", "out_channels", "num_layers", "num_single_layers", "attention_head_dim",
            "num_attention_heads", "joint_attention_dim", "pooled_projection_dim", "guidance_embeds", "axes_dims_rope")
    get = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    d = {k: get(k) for k in keys}
    if d["out_channels"] is None:
        d["out_channels"] = d["in_channels"]
    d["axes_dims_rope"] = tuple(d["axes_dims_rope"])
    if d["patch_size"] != 1:
        raise ValueError("textflux_b200: patch_size must be 1 (FLUX packs 2x2 patches in the pipeline)")
    if len(d["axes_dims_rope"]) != 3:
        raise ValueError("textflux_b200: axes_dims_rope must have 3 entries")
    return d


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Transformer2DModelOutput(SimpleNamespace):
    """Mirror of models/modeling_outputs.py Transformer2DModelOutput (attribute `sample`)."""


class B200FluxTransformer(torch.nn.Module):
    """Drop-in for FluxTransformer2DModel (transformer_flux.py:844-1212) on one B200.

    Build it from a loaded reference module (`from_reference`) or from any `get(name) -> tensor` source of the
    reference state dict (`from_getter`), then assign it to `pipe.transformer`.
    """

    def __init__(self, config, get: Callable[[str], Tensor], device: Union[str, torch.device] = "cuda",
                 gemm_cta_group: Optional[int] = None, use_graph: bool = True, gemm_mcast: Optional[int] = None,
                 lora_modules=None):
        """`lora_modules`: reference module names whose weights `get` returns with an adapter already folded in (a
        `fold_lora` getter); recorded so that a later `load_lora_weights` / `unload_lora_weights` restores them."""
        super().__init__()
        self._lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("textflux_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.config = FrozenConfig(_cfg_dict(config))
        self._dev = dev
        c = self.config
        tc = _lib.TfxConfig(c.in_channels, c.out_channels, c.num_layers, c.num_single_layers, c.attention_head_dim,
                            c.num_attention_heads, c.joint_attention_dim, c.pooled_projection_dim,
                            int(bool(c.guidance_embeds)), (C.c_int32 * 3)(*c.axes_dims_rope))
        h = C.c_void_p()
        _lib.check(self._lib.tfx_create(C.byref(tc), dev.index, C.byref(h)))
        self._h = h
        self.register_buffer("_anchor", torch.zeros(1, dtype=torch.bfloat16, device=dev), persistent=False)
        with torch.cuda.device(dev):
            self._weights: Dict[str, Tensor] = pack_weights(SimpleNamespace(**c), get, dev, torch.bfloat16)
        for name, t in self._weights.items():
            _lib.check(self._lib.tfx_set_weight(self._h, name.encode(), t.data_ptr(), t.shape[0], t.shape[1]), self._h)
        _lib.check(self._lib.tfx_finalize_weights(self._h), self._h)
        if gemm_cta_group is not None:
            self.set_option("gemm_cta_group", gemm_cta_group)
        self._shape: Optional[Tuple[int, int, int]] = None
        folded = getattr(get, "lora_modules", None) if lora_modules is None else lora_modules
        self._lora_modules: set = set(folded or ())
        self._side: Dict[str, Tensor] = {}  # side matrices of an unfused adapter ("<packed>.la" / ".lb"), kept alive here
        self.set_option("use_graph", int(use_graph))
        if gemm_mcast is not None:
            self.set_option("gemm_mcast", gemm_mcast)

    # ---- construction helpers -------------------------------------------------------------------------------
    @classmethod
    def from_reference(cls, module: torch.nn.Module, device="cuda", **kw) -> "B200FluxTransformer":
        """`module` is a loaded reference FluxTransformer2DModel (weights + `.config`); it can be freed afterwards.
        A module carrying PEFT adapters (FluxFillPipeline.load_lora_weights -> load_lora_into_transformer,
        loaders/lora_pipeline.py:1745-1820, as run_inference_lora.py:52-65 / demo_beta.py leave it) has its active
        adapters folded into the packed weights: W + sum_a scaling[a] * B_a @ A_a in fp32."""
        sd = module.state_dict()
        peft = _peft_layers(module)
        if not peft and any(".base_layer." in k or ".lora_A." in k for k in sd):
            raise ValueError("textflux_b200: the transformer's state dict has PEFT keys (…base_layer.weight / …lora_A.<adapter>"
                             ".weight) but no PEFT layer objects were found to read `scaling` from; fold the adapter "
                             "with textflux_b200.fold_lora(...) and build the engine from that getter instead")
        if not peft:
            return cls(module.config, sd.__getitem__, device=device, **kw)

        def get(name: str) -> Tensor:
            mod, _, leaf = name.rpartition(".")
            if mod not in peft:
                return sd[name]
            layer = peft[mod]
            base = sd[f"{mod}.base_layer.{leaf}"]
            if leaf != "weight":
                return base
            w = base.to(torch.float32)
            for a in _active_adapters(layer):
                A = sd[f"{mod}.lora_A.{a}.weight"].to(device=w.device, dtype=torch.float32)
                Bm = sd[f"{mod}.lora_B.{a}.weight"].to(device=w.device, dtype=torch.float32)
                w = w + float(layer.scaling[a]) * (Bm @ A)
            return w.to(base.dtype)

        return cls(module.config, get, device=device, lora_modules=set(peft), **kw)

    @classmethod
    def from_state_dict(cls, config, state_dict: Dict[str, Tensor], device="cuda", **kw) -> "B200FluxTransformer":
        return cls(config, state_dict.__getitem__, device=device, **kw)

    # ---- LoRA hot-swap (loaders/lora_pipeline.py: load_lora_weights / unload_lora_weights) ------------------------
    def load_lora_weights(self, get_base: Callable[[str], Tensor], lora: Dict[str, Tensor], scale: float = 1.0,
                          mode: str = "fold", side_threshold: float = 0.05) -> Dict[str, str]:
        """Install an adapter on the live engine, replacing any adapter installed before.  `get_base(name)` returns the BASE
        tensor of the reference state dict (e.g. `loader.Checkpoint(path).getter(device)` or `module.state_dict().__getitem__`).

        mode "fold": W + scale * alpha/r * B A in fp32 -> bf16, written in place into the packed matrices that hold a touched
        module (every device pointer, TMA descriptor and the captured step graph stay valid) -- free at run time.
        mode "side": the reference's own unfused form (run_inference_lora.py:52-65): the base weights stay, each touched packed
        GEMM gains one extension k-block computing `+ bf16(x A^T) (scale * alpha/r * B)^T` in the same accumulator (~3 % step time).
        mode "auto": per packed matrix, measured (`packer.fold_noise`): fold where the bf16 fold keeps the adapter to within
        `side_threshold` of its own norm, side path where the delta is too close to W's bf16 ulp for that.
        Returns {packed name: "fold" | "side"}."""
        from .packer import fold_lora, fold_noise, lora_modules, packed_layout, repack_modules, side_lora
        if mode not in ("fold", "side", "auto"):
            raise ValueError(f"mode must be 'fold', 'side' or 'auto', not {mode!r}")
        cfg = SimpleNamespace(**self.config)
        new_mods = set(lora_modules(lora))
        holders = {name: [r for r in refs if r in new_mods] for name, kind, refs in packed_layout(cfg) if kind == "lin"}
        holders = {n: r for n, r in holders.items() if r}
        if mode == "auto":
            noise = fold_noise(cfg, get_base, lora, scale=scale, device=self._dev)
            plan = {n: ("side" if noise[n] > side_threshold else "fold") for n in holders}
        else:
            plan = {n: mode for n in holders}
        fold_mods = {r for n, refs in holders.items() if plan[n] == "fold" for r in refs}
        with torch.cuda.device(self._dev):
            # modules folded before go back to base unless folded again; modules folded now get W + delta
            restore = (self._lora_modules | fold_mods)
            if restore:
                sub = {k: v for k, v in lora.items() if any(k.startswith(p + m + ".") for m in fold_mods for p in ("transformer.", ""))}
                repack_modules(cfg, fold_lora(get_base, sub, scale=scale) if sub else get_base, self._weights, restore)
            had_side = bool(self._side)
            for name in list(self._side):
                _lib.check(self._lib.tfx_unset_weight(self._h, name.encode()), self._h)
            self._side = side_lora(cfg, lora, self._dev, scale=scale, only=[n for n in holders if plan[n] == "side"])
            for name, t in self._side.items():
                _lib.check(self._lib.tfx_set_weight(self._h, name.encode(), t.data_ptr(), t.shape[0], t.shape[1]), self._h)
            torch.cuda.current_stream(self._dev).synchronize()
            if had_side or self._side:  # a fold-only swap changes values in place: descriptors and the step graph stay
                _lib.check(self._lib.tfx_finalize_weights(self._h), self._h)
        self._lora_modules = fold_mods
        self.set_option("mod_cache_reset", 1)  # modulation vectors cached from the old weights must not be served again
        return plan

    def unload_lora_weights(self, get_base: Callable[[str], Tensor]) -> None:
        """Restore the base weights of every module the current adapter touched and drop its side matrices."""
        from .packer import repack_modules
        with torch.cuda.device(self._dev):
            if self._lora_modules:
                repack_modules(SimpleNamespace(**self.config), get_base, self._weights, self._lora_modules)
            had_side = bool(self._side)
            for name in list(self._side):
                _lib.check(self._lib.tfx_unset_weight(self._h, name.encode()), self._h)
            self._side = {}
            torch.cuda.current_stream(self._dev).synchronize()
            if had_side:
                _lib.check(self._lib.tfx_finalize_weights(self._h), self._h)
        self._lora_modules = set()
        self.set_option("mod_cache_reset", 1)

    # ---- what DiffusionPipeline reads -------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self._dev

    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    def to(self, *args, **kwargs):
        """The engine is bound to its device at construction; `.to()` with the same device/dtype is a no-op."""
        dev, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if dev is not None and torch.device(dev).type != "cuda":
            raise RuntimeError("textflux_b200: the engine cannot leave its CUDA device (no CPU path)")
        if dev is not None and torch.device(dev).index not in (None, self._dev.index):
            raise RuntimeError(f"textflux_b200: engine lives on {self._dev}; build a new one for {dev}")
        if dtype is not None and dtype != torch.bfloat16:
            raise RuntimeError("textflux_b200: the engine computes in bf16 only")
        return self

    def set_option(self, key: str, value: int) -> None:
        _lib.check(self._lib.tfx_set_option(self._h, key.encode(), int(value)), self._h)
        if key in ("gemm_mcast", "mod_cache_slots"):
            self._shape = None  # the library dropped its workspace; re-run tfx_prepare on the next call

    def counter(self, key: str) -> int:
        v = C.c_int64()
        _lib.check(self._lib.tfx_get_counter(self._h, key.encode(), C.byref(v)), self._h)
        return v.value

    def __del__(self):
        h, lib = getattr(self, "_h", None), getattr(self, "_lib", None)
        if h and lib:
            lib.tfx_destroy(h)
            self._h = None

    # ---- helpers ------------------------------------------------------------------------------------------------
    def _prepare(self, B: int, S: int, T: int) -> None:
        if self._shape != (B, S, T):
            _lib.check(self._lib.tfx_prepare(self._h, B, S, T), self._h)
            self._shape = (B, S, T)

    def _bf16(self, t: Tensor, name: str, shape: Tuple[int, ...]) -> Tensor:
        if t is None:
            raise ValueError(f"textflux_b200: `{name}` is required")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"textflux_b200: `{name}` has shape {tuple(t.shape)}, expected {tuple(shape)}")
        if t.device != self._dev:
            raise ValueError(f"textflux_b200: `{name}` is on {t.device}, the engine on {self._dev}")
        return t.to(torch.bfloat16).contiguous()

    def _ids(self, ids: Tensor, name: str, n: int) -> Tensor:
        if ids.ndim == 3:  # deprecated batched ids (transformer_flux.py:1100-1113)
            ids = ids[0]
        out = self._bf16(ids, name, (n, 3))
        if ids.dtype != torch.bfloat16 and not torch.equal(out.to(ids.dtype), ids.to(self._dev)):
            # the engine's RoPE table takes bf16 ids, which is what the pipeline passes (pipeline_flux_fill.py:1728-1739
            # builds them in the latents' dtype); fp32 ids above 256 (> 4096 px) would be rounded
            raise ValueError(f"textflux_b200: `{name}` holds positions that bf16 cannot represent exactly")
        return out

    # ---- FluxTransformer2DModel.forward (transformer_flux.py:1028-1212) -------------------------------------------
    @torch.no_grad()
    def forward(self, hidden_states: Tensor, encoder_hidden_states: Tensor = None, pooled_projections: Tensor = None,
                timestep: Tensor = None, img_ids: Tensor = None, txt_ids: Tensor = None, guidance: Tensor = None,
                joint_attention_kwargs: Optional[dict] = None, controlnet_block_samples=None,
                controlnet_single_block_samples=None, return_dict: bool = True, controlnet_blocks_repeat: bool = False):
        if controlnet_block_samples is not None or controlnet_single_block_samples is not None:
            raise ValueError("textflux_b200: controlnet residuals are not part of the TextFlux path")
        if hidden_states.ndim != 3:
            raise ValueError("textflux_b200: hidden_states must be [B, S, in_channels]")
        c = self.config
        B, S, _ = hidden_states.shape
        T = encoder_hidden_states.shape[1]
        out_dtype = hidden_states.dtype
        hs = self._bf16(hidden_states, "hidden_states", (B, S, c.in_channels))
        enc = self._bf16(encoder_hidden_states, "encoder_hidden_states", (B, T, c.joint_attention_dim))
        pooled = self._bf16(pooled_projections, "pooled_projections", (B, c.pooled_projection_dim))
        t = self._bf16(timestep, "timestep", (B,))
        g = None
        if c.guidance_embeds:
            if guidance is None:
                raise ValueError("textflux_b200: `guidance` is required when config.guidance_embeds is set")
            g = guidance.to(device=self._dev, dtype=torch.float32).expand(B).contiguous()
        img = self._ids(img_ids, "img_ids", S)
        txt = self._ids(txt_ids, "txt_ids", T)
        out = torch.empty(B, S, c.out_channels, dtype=torch.bfloat16, device=self._dev)
        with torch.cuda.device(self._dev):
            self._prepare(B, S, T)
            stream = torch.cuda.current_stream(self._dev).cuda_stream
            _lib.check(self._lib.tfx_forward(self._h, hs.data_ptr(), enc.data_ptr(), pooled.data_ptr(), t.data_ptr(),
                                             _ptr(g), img.data_ptr(), txt.data_ptr(), out.data_ptr(), stream), self._h)
        out = out.to(out_dtype)
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)

    # ---- one fused sampling step: pipeline_flux_fill.py:2082-2098 ---------------------------------------------------
    @torch.no_grad()
    def step(self, latents: Tensor, cond: Tensor, encoder_hidden_states: Tensor, pooled_projections: Tensor,
             timestep: Tensor, guidance: Optional[Tensor], img_ids: Tensor, txt_ids: Tensor, sigma: float,
             sigma_next: float, return_noise_pred: bool = False):
        """latents' = scheduler.step(transformer(cat(latents, cond), ...), t, latents), the Euler update fused into the
        last GEMM's store.  `timestep` is t/1000 in bf16 as the pipeline would pass it."""
        c = self.config
        B, S, _ = latents.shape
        T = encoder_hidden_states.shape[1]
        lat = self._bf16(latents, "latents", (B, S, c.out_channels))
        cnd = self._bf16(cond, "cond", (B, S, c.in_channels - c.out_channels))
        enc = self._bf16(encoder_hidden_states, "encoder_hidden_states", (B, T, c.joint_attention_dim))
        pooled = self._bf16(pooled_projections, "pooled_projections", (B, c.pooled_projection_dim))
        t = self._bf16(timestep, "timestep", (B,))
        g = guidance.to(device=self._dev, dtype=torch.float32).expand(B).contiguous() if c.guidance_embeds else None
        img = self._ids(img_ids, "img_ids", S)
        txt = self._ids(txt_ids, "txt_ids", T)
        out = torch.empty_like(lat)
        pred = torch.empty_like(lat) if return_noise_pred else None
        with torch.cuda.device(self._dev):
            self._prepare(B, S, T)
            stream = torch.cuda.current_stream(self._dev).cuda_stream
            _lib.check(self._lib.tfx_step(self._h, lat.data_ptr(), cnd.data_ptr(), enc.data_ptr(), pooled.data_ptr(),
                                          t.data_ptr(), _ptr(g), img.data_ptr(), txt.data_ptr(), float(sigma),
                                          float(sigma_next), out.data_ptr(), _ptr(pred), stream), self._h)
        return (out, pred) if return_noise_pred else out

    @torch.no_grad()
    def set_schedule(self, timesteps: Tensor, guidance: Optional[Tensor], pooled_projections: Tensor, S: int, T: int) -> None:
        """Precompute temb and every adaLN modulation vector for all steps of a schedule (they depend only on
        (t_i, guidance, pooled)): `timesteps` is [n, B] = t/1000 in bf16, exactly what the pipeline would pass at step i."""
        c = self.config
        n, B = timesteps.shape
        ts = timesteps.to(device=self._dev, dtype=torch.bfloat16).contiguous()
        pooled = self._bf16(pooled_projections, "pooled_projections", (B, c.pooled_projection_dim))
        g = guidance.to(device=self._dev, dtype=torch.float32).expand(B).contiguous() if c.guidance_embeds else None
        with torch.cuda.device(self._dev):
            self._prepare(B, S, T)
            stream = torch.cuda.current_stream(self._dev).cuda_stream
            _lib.check(self._lib.tfx_set_schedule(self._h, ts.data_ptr(), n, _ptr(g), pooled.data_ptr(), stream), self._h)

    @torch.no_grad()
    def step_scheduled(self, i: int, latents: Tensor, cond: Tensor, encoder_hidden_states: Tensor, img_ids: Tensor,
                       txt_ids: Tensor, sigma: float, sigma_next: float, return_noise_pred: bool = False):
        """`step` for step i of the schedule given to `set_schedule` (bit-identical results, no per-step adaLN pass)."""
        c = self.config
        B, S, _ = latents.shape
        T = encoder_hidden_states.shape[1]
        lat = self._bf16(latents, "latents", (B, S, c.out_channels))
        cnd = self._bf16(cond, "cond", (B, S, c.in_channels - c.out_channels))
        enc = self._bf16(encoder_hidden_states, "encoder_hidden_states", (B, T, c.joint_attention_dim))
        img = self._ids(img_ids, "img_ids", S)
        txt = self._ids(txt_ids, "txt_ids", T)
        out = torch.empty_like(lat)
        pred = torch.empty_like(lat) if return_noise_pred else None
        with torch.cuda.device(self._dev):
            if self._shape != (B, S, T):
                raise RuntimeError("textflux_b200: step_scheduled called with a different shape than set_schedule")
            stream = torch.cuda.current_stream(self._dev).cuda_stream
            _lib.check(self._lib.tfx_step_scheduled(self._h, int(i), lat.data_ptr(), cnd.data_ptr(), enc.data_ptr(),
                                                    img.data_ptr(), txt.data_ptr(), float(sigma), float(sigma_next),
                                                    out.data_ptr(), _ptr(pred), stream), self._h)
        return (out, pred) if return_noise_pred else out

    @torch.no_grad()
    def denoise(self, latents: Tensor, cond: Tensor, prompt_embeds: Tensor, pooled_prompt_embeds: Tensor,
                txt_ids: Tensor, img_ids: Tensor, guidance_scale: float, num_inference_steps: int,
                scheduler: Optional["B200FlowMatchEulerScheduler"] = None,
                callback: Optional[Callable[[int, Tensor], None]] = None, precompute_modulation: bool = True,
                callback_on_step_end: Optional[Callable] = None,
                callback_on_step_end_tensor_inputs: Tuple[str, ...] = ("latents",)) -> Tensor:
        """The whole loop of FluxFillPipeline.__call__ (pipeline_flux_fill.py:2049-2119): schedule, then one fused
        step per timestep.  Returns the final packed latents.

        `callback_on_step_end(self, i, t, callback_kwargs) -> dict` follows the pipeline's contract (:2105-2112): it is
        handed the tensors named in `callback_on_step_end_tensor_inputs` ("latents", "prompt_embeds") and whatever it
        returns under those names REPLACES them for the following steps.  Setting `self.interrupt = True` (from the
        callback or another thread) skips the remaining steps like :2078-2079.  `callback(i, latents)` is the
        observe-only short form."""
        sch = scheduler or B200FlowMatchEulerScheduler()
        self.interrupt = False
        B, S = latents.shape[0], latents.shape[1]
        T = prompt_embeds.shape[1]
        mu = calculate_shift(S, sch.config.base_image_seq_len, sch.config.max_image_seq_len, sch.config.base_shift,
                             sch.config.max_shift)
        sch.set_timesteps(sigmas=np.linspace(1.0, 1 / num_inference_steps, num_inference_steps), device=self._dev, mu=mu)
        guidance = torch.full([1], guidance_scale, device=self._dev, dtype=torch.float32).expand(B) \
            if self.config.guidance_embeds else None
        # timestep = t.expand(B).to(latents.dtype); transformer(timestep / 1000)   (:2082,2086)
        ts = (sch.timesteps.to(self._dev)[:, None].expand(-1, B).to(latents.dtype) / 1000).contiguous()
        sig = sch.sigmas_cpu
        if precompute_modulation:
            self.set_schedule(ts, guidance, pooled_prompt_embeds, S, T)
        for i in range(len(sch.timesteps)):
            if self.interrupt:
                continue
            if precompute_modulation:
                latents = self.step_scheduled(i, latents, cond, prompt_embeds, img_ids, txt_ids, sig[i], sig[i + 1])
            else:
                latents = self.step(latents, cond, prompt_embeds, pooled_prompt_embeds, ts[i], guidance, img_ids, txt_ids,
                                    sig[i], sig[i + 1])
            if callback_on_step_end is not None:
                have = {"latents": latents, "prompt_embeds": prompt_embeds}
                unknown = [k for k in callback_on_step_end_tensor_inputs if k not in have]
                if unknown:
                    raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in ['latents', 'prompt_embeds'], "
                                     f"but found {unknown}")
                outs = callback_on_step_end(self, i, sch.timesteps[i], {k: have[k] for k in callback_on_step_end_tensor_inputs})
                latents = outs.pop("latents", latents)
                prompt_embeds = outs.pop("prompt_embeds", prompt_embeds)
            if callback is not None:
                callback(i, latents)
        return latents


# ------------------------------------------------------------------------------------------------------------------
def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16) -> float:
    """pipeline_flux_fill.py:1248-1258 (same default arguments)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


class B200FlowMatchEulerScheduler:
    """Drop-in for FlowMatchEulerDiscreteScheduler (scheduling_flow_match_euler_discrete.py:33-338) on the FLUX path:
    same config keys, same `set_timesteps` arithmetic (numpy float32 sigmas, dynamic time shift), same step-index
    bookkeeping; `step` runs the fused CUDA update instead of four eager kernels."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = True,
                 base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096):
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, shift=shift,
                                   use_dynamic_shifting=use_dynamic_shifting, base_shift=base_shift,
                                   max_shift=max_shift, base_image_seq_len=base_image_seq_len,
                                   max_image_seq_len=max_image_seq_len)
        ts = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        sig = torch.from_numpy(ts).to(torch.float32) / num_train_timesteps
        if not use_dynamic_shifting:
            sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = sig * num_train_timesteps
        self.sigmas = sig.to("cpu")
        self.sigmas_cpu: List[float] = self.sigmas.tolist()
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self._timesteps_cpu = self.timesteps.clone()
        self._step_index: Optional[int] = None
        self._begin_index: Optional[int] = None
        self.num_inference_steps: Optional[int] = None

    @classmethod
    def from_config(cls, config) -> "B200FlowMatchEulerScheduler":
        get = (lambda k, d: config.get(k, d)) if isinstance(config, dict) else (lambda k, d: getattr(config, k, d))
        unsupported = [k for k in ("use_karras_sigmas", "use_exponential_sigmas", "use_beta_sigmas", "invert_sigmas")
                       if get(k, False)]
        if get("shift_terminal", None):
            unsupported.append("shift_terminal")
        if unsupported:
            raise ValueError(f"textflux_b200: scheduler config sets {unsupported}, which the FLUX-Fill path never does and "
                             f"this scheduler does not implement; keep the reference scheduler for that configuration")
        return cls(get("num_train_timesteps", 1000), get("shift", 1.0), get("use_dynamic_shifting", True),
                   get("base_shift", 0.5), get("max_shift", 1.15), get("base_image_seq_len", 256),
                   get("max_image_seq_len", 4096))

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def time_shift(self, mu: float, sigma: float, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps: int = None, device=None, sigmas=None, mu: Optional[float] = None):
        """scheduling_flow_match_euler_discrete.py:184-241."""
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError(" you have a pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        n_train = self.config.num_train_timesteps
        if sigmas is None:
            timesteps = np.linspace(self.sigma_max * n_train, self.sigma_min * n_train, num_inference_steps)
            sigmas = timesteps / n_train
        else:
            sigmas = np.array(sigmas).astype(np.float32)
            num_inference_steps = len(sigmas)
        self.num_inference_steps = num_inference_steps
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        sig = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32)
        ts = sig * n_train
        sig = torch.cat([sig, torch.zeros(1)])
        self._timesteps_cpu = ts.clone()
        self.sigmas_cpu = sig.tolist()
        self.timesteps = ts.to(device=device)
        self.sigmas = sig.to(device=device)
        self._step_index = None
        self._begin_index = None

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        st = self._timesteps_cpu if schedule_timesteps is None else schedule_timesteps.detach().to("cpu")
        t = float(timestep.detach().to("cpu", torch.float32)) if isinstance(timestep, torch.Tensor) else float(timestep)
        idx = (st == t).nonzero()
        pos = 1 if len(idx) > 1 else 0
        return idx[pos].item()

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self._begin_index is None else self._begin_index

    @torch.no_grad()
    def step(self, model_output: Tensor, timestep, sample: Tensor, s_churn: float = 0.0, s_tmin: float = 0.0,
             s_tmax: float = float("inf"), s_noise: float = 1.0, generator=None, return_dict: bool = True):
        """scheduling_flow_match_euler_discrete.py:265-338."""
        if isinstance(timestep, int) or (isinstance(timestep, torch.Tensor)
                                         and timestep.dtype in (torch.int32, torch.int64)):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                             " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                             " one of the `scheduler.timesteps` as a timestep.")
        if self._step_index is None:
            self._init_step_index(timestep)
        if model_output.device.type != "cuda":
            raise RuntimeError("textflux_b200: scheduler.step runs on CUDA tensors only (no CPU path)")
        if model_output.dtype != torch.bfloat16 or model_output.shape != sample.shape:
            raise ValueError("textflux_b200: scheduler.step expects bf16 model_output with sample's shape")
        if sample.dtype != torch.bfloat16:
            # the reference upcasts `sample` to fp32 and casts the result back to model_output.dtype (:322-330); with an fp32
            # sample that is a different rounding than this bf16 kernel implements
            raise ValueError("textflux_b200: scheduler.step expects a bf16 sample (the FLUX-Fill pipeline's latents dtype)")
        sigma, sigma_next = self.sigmas_cpu[self._step_index], self.sigmas_cpu[self._step_index + 1]
        v = model_output.contiguous()
        x = sample.to(torch.bfloat16).contiguous()
        out = torch.empty_like(v)
        lib = _lib.load()
        with torch.cuda.device(v.device):
            stream = torch.cuda.current_stream(v.device).cuda_stream
            _lib.check(lib.tfx_euler_step(v.data_ptr(), x.data_ptr(), out.data_ptr(), v.numel(), float(sigma),
                                          float(sigma_next), stream))
        self._step_index += 1
        if not return_dict:
            return (out,)
        return SimpleNamespace(prev_sample=out)

    def __len__(self):
        return self.config.num_train_timesteps


class B200StochasticRFOvershotScheduler(B200FlowMatchEulerScheduler):
    """Drop-in for StochasticRFOvershotDiscreteScheduler (scheduling_stochastic_rf_discrete_overshot.py), TextFlux's
    default "overshoot"/AMO sampler (demo.py:15, run_inference.py:79-91), in the configuration TextFlux uses: attn_map =
    None and any `overshot_func(t, dt)` evaluated on the host.  Differences from the Euler scheduler it mirrors exactly:
    `set_timesteps` shifts the pipeline's float64 sigmas in float64 (:203-211), `step` re-noises
    (x_o * a + eps * b, eps from torch's generator, :351-357) and returns (prev_sample, predicted_x1)."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.c = 2.0
        self.overshot_func = lambda t, dt: t + dt
        self.attn_map = None

    def set_c(self, c: float):
        self.c = c

    def set_overshot_func(self, overshot_func):
        self.overshot_func = overshot_func

    def set_attn_map(self, attn_map):
        if attn_map is not None:
            raise ValueError("textflux_b200: per-token attn_map overshoot is not part of the TextFlux path")

    def set_timesteps(self, num_inference_steps: int = None, device=None, sigmas=None, mu: Optional[float] = None):
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError(" you have a pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        n_train = self.config.num_train_timesteps
        if sigmas is None:
            self.num_inference_steps = num_inference_steps
            timesteps = np.linspace(self.sigma_max * n_train, self.sigma_min * n_train, num_inference_steps)
            sigmas = timesteps / n_train
        sigmas = np.asarray(sigmas)  # no float32 cast here, unlike the Euler scheduler: the shift runs in float64
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        sig = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32)
        ts = sig * n_train
        sig = torch.cat([sig, torch.zeros(1)])
        self._timesteps_cpu = ts.clone()
        self.sigmas_cpu = sig.tolist()
        self.timesteps = ts.to(device=device)
        self.sigmas = sig.to(device=device)
        self._step_index = None
        self._begin_index = None

    def _scalars(self, i: int):
        """(t_o - t, a, b, sigma) in the reference's fp32 0-dim tensor arithmetic (:322-350)."""
        f = np.float32
        sigma, sigma_next = f(self.sigmas_cpu[i]), f(self.sigmas_cpu[i + 1])
        t = f(1) - sigma
        step = sigma - sigma_next
        tn = t + step
        t_next = tn if tn <= 1 else 1
        so = f(step * f(self.c))
        to_ = self.overshot_func(t_next, so)
        to_ = f(to_) if not isinstance(to_, (int,)) else to_
        t_o = to_ if to_ <= 1 else 1
        coef = f(f(t_o) - t)
        a = f(f(t_next) / f(t_o))
        with np.errstate(invalid="ignore"):
            r = f(f(f(1) - f(t_next)) * f(f(1) - f(t_next))) - f(f(a - f(t_next)) * f(a - f(t_next)))
            b = f(np.sqrt(f(r)))
        return float(coef), float(a), float(b), float(sigma)

    @torch.no_grad()
    def step(self, model_output: Tensor, timestep, sample: Tensor, s_churn: float = 0.0, s_tmin: float = 0.0,
             s_tmax: float = float("inf"), s_noise: float = 1.0, generator=None, return_dict: bool = True):
        if isinstance(timestep, int) or (isinstance(timestep, torch.Tensor)
                                         and timestep.dtype in (torch.int32, torch.int64)):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                             " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                             " one of the `scheduler.timesteps` as a timestep.")
        if self._step_index is None:
            self._init_step_index(timestep)
        if model_output.device.type != "cuda":
            raise RuntimeError("textflux_b200: scheduler.step runs on CUDA tensors only (no CPU path)")
        if model_output.dtype != torch.bfloat16 or model_output.shape != sample.shape:
            raise ValueError("textflux_b200: scheduler.step expects bf16 model_output with sample's shape")
        coef, a, b, sigma = self._scalars(self._step_index)
        v = model_output.contiguous()
        x = sample.to(torch.bfloat16).contiguous()
        # randn_tensor(sample.shape, generator, device=sample.device, dtype=float32) (utils/torch_utils.py:38-83): a CPU
        # generator draws on the CPU and the result is moved, so seeds reproduce across devices
        if generator is not None and generator.device.type == "cpu":
            eps = torch.randn(x.shape, generator=generator, dtype=torch.float32).to(x.device)
        else:
            eps = torch.randn(x.shape, generator=generator, device=x.device, dtype=torch.float32)
        prev = torch.empty_like(v)
        x1 = torch.empty(v.shape, dtype=torch.float32, device=v.device)
        lib = _lib.load()
        with torch.cuda.device(v.device):
            stream = torch.cuda.current_stream(v.device).cuda_stream
            _lib.check(lib.tfx_overshoot_step(v.data_ptr(), x.data_ptr(), eps.data_ptr(), prev.data_ptr(), x1.data_ptr(),
                                              v.numel(), coef, a, b, sigma, stream))
        self._step_index += 1
        if not return_dict:
            return (prev, x1)
        return SimpleNamespace(prev_sample=prev, predicted_x1=x1)


def _peft_layers(module: torch.nn.Module) -> Dict[str, torch.nn.Module]:
    """{module name: PEFT LoRA layer} for every wrapped Linear of `module` (peft.tuners.lora.layer.Linear keeps the
    original as `.base_layer` and the adapters in `.lora_A` / `.lora_B` ModuleDicts with `.scaling[adapter]`)."""
    out = {}
    for name, m in module.named_modules():
        if hasattr(m, "base_layer") and hasattr(m, "lora_A") and hasattr(m, "lora_B") and hasattr(m, "scaling"):
            out[name] = m
    return out


def _active_adapters(layer) -> List[str]:
    if getattr(layer, "merged", False) or getattr(layer, "disable_adapters", False):
        return []
    act = getattr(layer, "active_adapters", None)
    if act is None:
        act = getattr(layer, "active_adapter", None)
    if act is None:
        act = list(layer.lora_A.keys())
    if isinstance(act, str):
        act = [act]
    return [a for a in act if a in layer.lora_A]


def attach(pipe, vae="auto", text_encoders="auto", **engine_kw):
    """Swap the engine into a loaded reference FluxFillPipeline in place (zero edits to reference files):
    `pipe.transformer` -> B200FluxTransformer built from the loaded weights, `pipe.scheduler` -> fused scheduler, `pipe.vae` ->
    B200AutoencoderKL and `pipe.text_encoder` / `pipe.text_encoder_2` -> B200CLIPTextEncoder / B200T5Encoder (`True`: required;
    "auto": when the module's config is one the engine implements, else the reference module stays; False: never)."""
    eng = B200FluxTransformer.from_reference(pipe.transformer, device=pipe.transformer.device, **engine_kw)
    if type(pipe.scheduler).__name__ == "StochasticRFOvershotDiscreteScheduler":
        sch = B200StochasticRFOvershotScheduler.from_config(pipe.scheduler.config)
        sch.set_c(getattr(pipe.scheduler, "c", 2.0))
        sch.set_overshot_func(getattr(pipe.scheduler, "overshot_func", lambda t, dt: t + dt))
    else:
        sch = B200FlowMatchEulerScheduler.from_config(pipe.scheduler.config)
    if vae and getattr(pipe, "vae", None) is not None and type(pipe.vae).__name__ == "AutoencoderKL":
        from .vae import B200AutoencoderKL
        try:
            pipe.vae = B200AutoencoderKL.from_reference(pipe.vae, device=eng.device)
        except ValueError:
            if vae is True:
                raise
    if text_encoders:
        from .text_encoders import B200CLIPTextEncoder, B200T5Encoder
        for attr, ref_cls, cls in (("text_encoder", "CLIPTextModel", B200CLIPTextEncoder), ("text_encoder_2", "T5EncoderModel", B200T5Encoder)):
            mod = getattr(pipe, attr, None)
            if mod is None or type(mod).__name__ != ref_cls:
                continue
            try:
                setattr(pipe, attr, cls.from_reference(mod, device=eng.device))
            except ValueError:
                if text_encoders is True:
                    raise
    pipe.transformer = eng
    pipe.scheduler = sch
    return pipe
