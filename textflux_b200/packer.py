"""Reference state-dict -> engine weight layout.

The engine consumes the tensors of FluxTransformer2DModel.state_dict() (names in SURVEY.md Appendix A) re-grouped so
that each fused kernel streams one matrix:

    d{i}.qkv_x / qkv_c   [3D, D]   to_q;to_k;to_v  /  add_q_proj;add_k_proj;add_v_proj   (attention_processor.py:237-260)
    d{i}.out_x / out_c   [D, D]    to_out.0 / to_add_out
    d{i}.ff1_*, ff2_*    [4D, D], [D, 4D]   ff.net.0.proj, ff.net.2 / ff_context.*      (attention.py:1218-1232)
    s{j}.qkvmlp          [7D, D]   to_q;to_k;to_v;proj_mlp                               (transformer_flux.py:694,702-713)
    s{j}.out             [D, 5D]   proj_out (input order attn | mlp, transformer_flux.py:732)
    mod                  [(12L + 3Ls + 2) D, D]   every adaLN linear of the model, in block order:
                         per double block norm1.linear (6D) then norm1_context.linear (6D); per single block
                         norm.linear (3D); norm_out.linear (2D)                           (normalization.py:148,187,353)
    t_embed / g_embed / p_embed .l1/.l2     time_text_embed MLPs                          (embeddings.py:1318-1339)

Row re-ordering only: every value is copied bit-exactly (tests/test_host_cpu.py::test_packer_is_a_bit_exact_row_regrouping).  LoRA adapters are folded at pack
time, W <- W + (alpha/r) * B @ A in fp32 (loaders/lora_pipeline.py:1618-1743 file format), see fold_lora().
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

Tensor = torch.Tensor


def reference_names(cfg) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every tensor FluxTransformer2DModel.state_dict() holds for `cfg`."""
    D, dh = cfg.num_attention_heads * cfg.attention_head_dim, cfg.attention_head_dim
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def lin(n, o, i):
        out.append((n + ".weight", (o, i)))
        out.append((n + ".bias", (o,)))

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
        for n in ("to_q", "to_k", "to_v", "add_k_proj", "add_v_proj", "add_q_proj", "to_out.0", "to_add_out"):
            lin(p + "attn." + n, D, D)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            out.append((p + f"attn.{n}.weight", (dh,)))
        lin(p + "ff.net.0.proj", 4 * D, D)
        lin(p + "ff.net.2", D, 4 * D)
        lin(p + "ff_context.net.0.proj", 4 * D, D)
        lin(p + "ff_context.net.2", D, 4 * D)
    for j in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{j}."
        lin(p + "norm.linear", 3 * D, D)
        lin(p + "proj_mlp", 4 * D, D)
        lin(p + "proj_out", D, 5 * D)
        for n in ("to_q", "to_k", "to_v"):
            lin(p + "attn." + n, D, D)
        for n in ("norm_q", "norm_k"):
            out.append((p + f"attn.{n}.weight", (dh,)))
    lin("norm_out.linear", 2 * D, D)
    lin("proj_out", cfg.out_channels, D)
    return out


def _cat_lin(get, names: Iterable[str], device, dtype) -> Tuple[Tensor, Tensor]:
    names = list(names)
    w = torch.cat([get(n + ".weight").to(device=device, dtype=dtype) for n in names], dim=0).contiguous()
    b = torch.cat([get(n + ".bias").to(device=device, dtype=dtype) for n in names], dim=0).reshape(1, -1).contiguous()
    return w, b


def packed_layout(cfg) -> List[Tuple[str, str, List[str]]]:
    """The packed layout as data: (packed name, kind, reference module/tensor names) in pack order.
    kind "lin": `<packed>.w` / `<packed>.b` = row-concatenation of the modules' weights / biases;
    kind "vec": `<packed>` = one reference tensor reshaped [1, n]."""
    L: List[Tuple[str, str, List[str]]] = []

    def put(name, names):
        L.append((name, "lin", list(names)))

    def vec(name, ref):
        L.append((name, "vec", [ref]))

    put("x_embedder", ["x_embedder"])
    put("context_embedder", ["context_embedder"])
    put("t_embed.l1", ["time_text_embed.timestep_embedder.linear_1"])
    put("t_embed.l2", ["time_text_embed.timestep_embedder.linear_2"])
    if cfg.guidance_embeds:
        put("g_embed.l1", ["time_text_embed.guidance_embedder.linear_1"])
        put("g_embed.l2", ["time_text_embed.guidance_embedder.linear_2"])
    put("p_embed.l1", ["time_text_embed.text_embedder.linear_1"])
    put("p_embed.l2", ["time_text_embed.text_embedder.linear_2"])
    put("proj_out", ["proj_out"])
    mod_names: List[str] = []
    for i in range(cfg.num_layers):
        r, d = f"transformer_blocks.{i}.", f"d{i}."
        put(d + "qkv_x", [r + "attn.to_q", r + "attn.to_k", r + "attn.to_v"])
        put(d + "qkv_c", [r + "attn.add_q_proj", r + "attn.add_k_proj", r + "attn.add_v_proj"])
        put(d + "out_x", [r + "attn.to_out.0"])
        put(d + "out_c", [r + "attn.to_add_out"])
        put(d + "ff1_x", [r + "ff.net.0.proj"])
        put(d + "ff2_x", [r + "ff.net.2"])
        put(d + "ff1_c", [r + "ff_context.net.0.proj"])
        put(d + "ff2_c", [r + "ff_context.net.2"])
        vec(d + "rms_q_x", r + "attn.norm_q.weight")
        vec(d + "rms_k_x", r + "attn.norm_k.weight")
        vec(d + "rms_q_c", r + "attn.norm_added_q.weight")
        vec(d + "rms_k_c", r + "attn.norm_added_k.weight")
        mod_names += [r + "norm1.linear", r + "norm1_context.linear"]
    for j in range(cfg.num_single_layers):
        r, s = f"single_transformer_blocks.{j}.", f"s{j}."
        put(s + "qkvmlp", [r + "attn.to_q", r + "attn.to_k", r + "attn.to_v", r + "proj_mlp"])
        put(s + "out", [r + "proj_out"])
        vec(s + "rms_q", r + "attn.norm_q.weight")
        vec(s + "rms_k", r + "attn.norm_k.weight")
        mod_names.append(r + "norm.linear")
    mod_names.append("norm_out.linear")
    put("mod", mod_names)
    return L


def pack_weights(cfg, get: Callable[[str], Tensor], device, dtype=torch.bfloat16) -> Dict[str, Tensor]:
    """Build the packed layout on `device`.  `get(name)` returns the reference tensor (any device); it is called once
    per tensor, block by block, so a 12B model never needs a second full copy in memory."""
    P: Dict[str, Tensor] = {}
    for name, kind, refs in packed_layout(cfg):
        if kind == "lin":
            P[name + ".w"], P[name + ".b"] = _cat_lin(get, refs, device, dtype)
        else:
            P[name] = get(refs[0]).to(device=device, dtype=dtype).reshape(1, -1).contiguous()
    return P


def repack_modules(cfg, get: Callable[[str], Tensor], P: Dict[str, Tensor], modules: Iterable[str]) -> List[str]:
    """Rewrite IN PLACE the packed matrices that contain any of `modules` (reference module names such as
    `transformer_blocks.3.attn.to_q`) from `get`; every other packed tensor, and every device pointer, stays as it is.
    This is the hot-swap path of LoRA adapters: only the rows of a touched module change (its fused neighbours are
    re-copied bit-exactly).  Returns the packed names rewritten."""
    modules = set(modules)
    done = []
    for name, kind, refs in packed_layout(cfg):
        if kind != "lin" or not modules.intersection(refs):
            continue
        row = 0
        w = P[name + ".w"]
        for ref in refs:
            n = get(ref + ".weight")
            if ref in modules:
                w[row:row + n.shape[0]].copy_(n.to(device=w.device, dtype=w.dtype))
            row += n.shape[0]
        done.append(name)
    return done


def lora_modules(lora: Dict[str, Tensor], prefix: str = "transformer.") -> List[str]:
    """Reference module names an adapter state dict touches."""
    out = []
    for k in lora:
        if k.endswith(".lora_A.weight"):
            m = k[: -len(".lora_A.weight")]
            out.append(m[len(prefix):] if m.startswith(prefix) else m)
    return out


def fold_lora(get: Callable[[str], Tensor], lora: Dict[str, Tensor], scale: float = 1.0,
              prefix: str = "transformer.") -> Callable[[str], Tensor]:
    """Wrap `get` so that `<module>.weight` returns W + scale*(alpha/r) * lora_B @ lora_A (fp32 math, then the
    caller's cast).  Keys follow the diffusers/PEFT file format: `transformer.<module>.lora_A.weight [r, in]`,
    `.lora_B.weight [out, r]`, optional `.alpha` (loaders/lora_pipeline.py:1618-1743; train_lora.py:527-532 uses
    alpha = r, i.e. factor 1)."""
    mods = {}
    for k in lora:
        if k.endswith(".lora_A.weight"):
            m = k[: -len(".lora_A.weight")]
            mods[m[len(prefix):] if m.startswith(prefix) else m] = m

    def wrapped(name: str) -> Tensor:
        w = get(name)
        if name.endswith(".weight") and name[: -len(".weight")] in mods:
            m = mods[name[: -len(".weight")]]
            A = lora[m + ".lora_A.weight"].to(device=w.device, dtype=torch.float32)
            Bm = lora[m + ".lora_B.weight"].to(device=w.device, dtype=torch.float32)
            r = A.shape[0]
            alpha = float(lora[m + ".alpha"]) if (m + ".alpha") in lora else float(r)
            w = (w.to(torch.float32) + (scale * alpha / r) * (Bm @ A)).to(w.dtype)
        return w

    wrapped.lora_modules = set(mods)  # B200FluxTransformer records them so a later swap / unload restores these modules
    return wrapped


SIDE_RANK = 64  # columns of the GEMM's extension k-block: the ranks of the modules packed into one matrix must fit


def _lora_index(lora: Dict[str, Tensor], prefix: str) -> Dict[str, str]:
    mods = {}
    for k in lora:
        if k.endswith(".lora_A.weight"):
            m = k[: -len(".lora_A.weight")]
            mods[m[len(prefix):] if m.startswith(prefix) else m] = m
    return mods


def _lora_factor(lora: Dict[str, Tensor], m: str, scale: float) -> float:
    r = lora[m + ".lora_A.weight"].shape[0]
    alpha = float(lora[m + ".alpha"]) if (m + ".alpha") in lora else float(r)
    return scale * alpha / r


def side_lora(cfg, lora: Dict[str, Tensor], device, scale: float = 1.0, prefix: str = "transformer.",
              only: Optional[Iterable[str]] = None, dtype=torch.bfloat16) -> Dict[str, Tensor]:
    """The adapter in the layout of the engine's UNFUSED path (include/textflux_b200.h, tfx_set_weight): for every packed
    matrix `<p>.w` [N, K] that holds an adapted module, `<p>.la` [64, K] = the modules' lora_A stacked (rank rows each, zero
    padded to 64) and `<p>.lb` [N, 64] = scale * alpha/r * lora_B of each module in that module's rows and rank columns.
    The packed GEMM then computes x W^T + bf16(x la^T) lb^T -- PEFT's `base(x) + scaling * lora_B(lora_A(x))`
    (run_inference_lora.py:52-65 keeps adapters unfused) -- and the base weight stays untouched.  `only`: packed names to build."""
    mods = _lora_index(lora, prefix)
    shapes = dict(reference_names(cfg))
    only = None if only is None else set(only)
    out: Dict[str, Tensor] = {}
    for name, kind, refs in packed_layout(cfg):
        if kind != "lin" or not any(r in mods for r in refs) or (only is not None and name not in only):
            continue
        N = sum(shapes[r + ".weight"][0] for r in refs)
        K = shapes[refs[0] + ".weight"][1]
        la = torch.zeros(SIDE_RANK, K, device=device, dtype=dtype)
        lb = torch.zeros(N, SIDE_RANK, device=device, dtype=dtype)
        row = col = 0
        for r in refs:
            o = shapes[r + ".weight"][0]
            if r in mods:
                A, Bm = lora[mods[r] + ".lora_A.weight"], lora[mods[r] + ".lora_B.weight"]
                rk = A.shape[0]
                if col + rk > SIDE_RANK:
                    raise ValueError(f"textflux_b200: the adapters packed into '{name}' have total rank {col + rk} > {SIDE_RANK}; "
                                     f"use the folded path for it")
                la[col:col + rk].copy_(A.to(device=device, dtype=dtype))
                lb[row:row + o, col:col + rk].copy_((Bm.to(device=device, dtype=torch.float32) * _lora_factor(lora, mods[r], scale)).to(dtype))
                col += rk
            row += o
        out[name + ".la"], out[name + ".lb"] = la, lb
    return out


def fold_noise(cfg, get: Callable[[str], Tensor], lora: Dict[str, Tensor], scale: float = 1.0, prefix: str = "transformer.",
               device=None) -> Dict[str, float]:
    """How much of an adapter a bf16 fold loses, measured: per packed matrix, the largest
    || (bf16(W + d) - W) - d ||_F / || d ||_F over its adapted modules, d = scale * alpha/r * B A in fp32.  Near 0 the fold realises
    the adapter; it grows as d shrinks towards W's bf16 ulp, where the unfused side path should be used."""
    mods = _lora_index(lora, prefix)
    out: Dict[str, float] = {}
    for name, kind, refs in packed_layout(cfg):
        if kind != "lin":
            continue
        worst = None
        for r in refs:
            if r not in mods:
                continue
            w = get(r + ".weight")
            dev = device or w.device
            w32 = w.to(device=dev, dtype=torch.float32)
            A = lora[mods[r] + ".lora_A.weight"].to(device=dev, dtype=torch.float32)
            Bm = lora[mods[r] + ".lora_B.weight"].to(device=dev, dtype=torch.float32)
            d = _lora_factor(lora, mods[r], scale) * (Bm @ A)
            folded = (w32 + d).to(torch.bfloat16).to(torch.float32)
            e = float(((folded - w32) - d).norm() / d.norm().clamp_min(1e-30))
            worst = e if worst is None else max(worst, e)
        if worst is not None:
            out[name] = worst
    return out


def synthetic_getter(cfg, seed: int, device, dtype=torch.bfloat16, w_std=0.02, b_std=0.02, rms_std=0.1):
    """Random weights of the reference's shapes generated on `device`, one tensor at a time (bench / smoke use: there
    is no network for real checkpoints).  N(0, 0.02^2) linears, RMSNorm weights 1 + N(0, 0.1^2)."""
    shapes = dict(reference_names(cfg))
    order = {n: i for i, (n, _) in enumerate(reference_names(cfg))}

    def get(name: str) -> Tensor:
        g = torch.Generator(device=device).manual_seed(seed * 100003 + order[name])
        t = torch.randn(shapes[name], generator=g, device=device, dtype=torch.float32)
        if name.endswith(".bias"):
            t = t * b_std
        elif ".norm_" in name:
            t = 1.0 + t * rms_std
        else:
            t = t * w_std
        return t.to(dtype)

    return get
