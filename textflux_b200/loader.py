"""Checkpoint files straight into the packed weight layout.

What the reference does for this: `FluxTransformer2DModel.from_pretrained(dir, torch_dtype=bf16)` reads
`diffusion_pytorch_model.safetensors` or the shards named by `diffusion_pytorch_model.safetensors.index.json`
(utils/constants.py:29-35) into an nn.Module, and `pipe.load_lora_weights(dir)` reads `pytorch_lora_weights.safetensors`
(loaders/lora_pipeline.py:1618-1743: keys `transformer.<module>.lora_A.weight [r, in]`, `.lora_B.weight [out, r]`,
optional `.alpha`) and keeps the adapters as separate PEFT layers at inference (run_inference_lora.py:52-65).

Here no nn.Module is ever built: a safetensors file is an 8-byte little-endian header length, a JSON header
{name: {dtype, shape, data_offsets}} and raw little-endian tensor bytes; `SafetensorsFile` memory-maps it and hands out
one tensor at a time, `pack_weights` asks for them block by block and concatenates them on the device, so the host never
holds more than one tensor and the device holds exactly the packed 23.8 GB.  LoRA adapters are folded at load
(`W + scale * alpha/r * B A` in fp32, packer.fold_lora).  Host-side logic only; the arithmetic stays in torch ops on the
target device.
"""
from __future__ import annotations

import json
import mmap
import os
import struct
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

from .packer import fold_lora, reference_names

CONFIG_NAME = "config.json"                                                   # utils/constants.py:29
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"              # :34
SAFE_WEIGHTS_INDEX_NAME = "diffusion_pytorch_model.safetensors.index.json"    # :35
LORA_WEIGHT_NAME_SAFE = "pytorch_lora_weights.safetensors"                    # loaders/lora_base.py
TRANSFORMERS_WEIGHTS_NAME = "model.safetensors"                               # transformers' names for the two text encoders
TRANSFORMERS_INDEX_NAME = "model.safetensors.index.json"

_DTYPES = {"BF16": torch.bfloat16, "F16": torch.float16, "F32": torch.float32, "F64": torch.float64, "I64": torch.int64,
           "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool}
_NAMES = {v: k for k, v in _DTYPES.items()}


class SafetensorsFile:
    """Read-only view of one .safetensors file (memory-mapped; tensors are materialised one at a time)."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        if size < 8:
            raise ValueError(f"{path}: not a safetensors file (shorter than its 8-byte header length)")
        (n,) = struct.unpack("<Q", self._f.read(8))
        if n > size - 8 or n > 100 * 1024 * 1024:
            raise ValueError(f"{path}: header length {n} is inconsistent with the file size {size}")
        header = json.loads(self._f.read(n).decode("utf-8"))
        self.metadata: Dict[str, str] = header.pop("__metadata__", {}) or {}
        self._data0 = 8 + n
        self._entries: Dict[str, Tuple[torch.dtype, Tuple[int, ...], int, int]] = {}
        for name, e in header.items():
            if e["dtype"] not in _DTYPES:
                raise ValueError(f"{path}: tensor {name} has unsupported dtype {e['dtype']}")
            b, en = e["data_offsets"]
            dt = _DTYPES[e["dtype"]]
            numel = 1
            for s in e["shape"]:
                numel *= int(s)
            if en - b != numel * torch.empty((), dtype=dt).element_size() or self._data0 + en > size or b < 0:
                raise ValueError(f"{path}: tensor {name} has inconsistent offsets {e['data_offsets']} for shape {e['shape']}")
            self._entries[name] = (dt, tuple(int(s) for s in e["shape"]), b, en)
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size > 0 else None

    def keys(self) -> List[str]:
        return list(self._entries)

    def __contains__(self, name: str) -> bool:
        return name in self._entries

    def shape(self, name: str) -> Tuple[int, ...]:
        return self._entries[name][1]

    def dtype(self, name: str) -> torch.dtype:
        return self._entries[name][0]

    def get(self, name: str, device="cpu", dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """A fresh tensor holding `name` on `device` (the file bytes are copied exactly; `dtype` casts afterwards)."""
        if name not in self._entries:
            raise KeyError(f"{self.path}: no tensor named {name}")
        dt, shape, b, e = self._entries[name]
        if e == b:
            t = torch.empty(shape, dtype=dt)
        else:
            # bytearray copy: torch.frombuffer needs a writable buffer, the mapping is read-only
            t = torch.frombuffer(bytearray(self._mm[self._data0 + b:self._data0 + e]), dtype=dt).reshape(shape)
        t = t.to(device)
        return t if dtype is None else t.to(dtype)

    def close(self):
        if self._mm is not None:
            self._mm.close()
            self._mm = None
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def save_safetensors(tensors: Dict[str, torch.Tensor], path: str, metadata: Optional[Dict[str, str]] = None) -> None:
    """Minimal writer of the same format (tests, and exporting a folded / packed checkpoint)."""
    header, off = {}, 0
    if metadata:
        header["__metadata__"] = metadata
    blobs = []
    for name in sorted(tensors):
        t = tensors[name].detach().contiguous().cpu()
        if t.dtype not in _NAMES:
            raise ValueError(f"{name}: dtype {t.dtype} cannot be stored")
        raw = t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    h = json.dumps(header, separators=(",", ":")).encode("utf-8")
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for raw in blobs:
            f.write(raw)


class Checkpoint:
    """A checkpoint directory (single file or shards + index, under diffusers' or transformers' file names) or a single
    .safetensors file."""

    def __init__(self, path: str):
        self.files: Dict[str, SafetensorsFile] = {}
        self.where: Dict[str, str] = {}
        if os.path.isdir(path):
            for index_name, single_name in ((SAFE_WEIGHTS_INDEX_NAME, SAFETENSORS_WEIGHTS_NAME), (TRANSFORMERS_INDEX_NAME, TRANSFORMERS_WEIGHTS_NAME)):
                index, single = os.path.join(path, index_name), os.path.join(path, single_name)
                if os.path.exists(index):
                    with open(index) as f:
                        wm = json.load(f)["weight_map"]
                    for name, fn in wm.items():
                        self.where[name] = os.path.join(path, fn)
                    break
                if os.path.exists(single):
                    self._add(single)
                    break
            else:
                raise FileNotFoundError(f"{path}: no {SAFE_WEIGHTS_INDEX_NAME}, {SAFETENSORS_WEIGHTS_NAME}, {TRANSFORMERS_INDEX_NAME} "
                                        f"or {TRANSFORMERS_WEIGHTS_NAME} found")
            cfg = os.path.join(path, CONFIG_NAME)
            self.config = json.load(open(cfg)) if os.path.exists(cfg) else None
        else:
            self._add(path)
            self.config = None

    def _file(self, fn: str) -> SafetensorsFile:
        if fn not in self.files:
            self.files[fn] = SafetensorsFile(fn)
        return self.files[fn]

    def _add(self, fn: str):
        for name in self._file(fn).keys():
            self.where[name] = fn

    def keys(self) -> List[str]:
        return list(self.where)

    def getter(self, device="cpu") -> Callable[[str], torch.Tensor]:
        def get(name: str) -> torch.Tensor:
            if name not in self.where:
                raise KeyError(f"checkpoint has no tensor named {name}")
            return self._file(self.where[name]).get(name, device)
        return get

    def check(self, cfg) -> None:
        """Every tensor FluxTransformer2DModel needs is present with the right shape (what load_state_dict(strict=True)
        reports as missing / size-mismatch keys)."""
        missing, bad = [], []
        for name, shape in reference_names(cfg):
            if name not in self.where:
                missing.append(name)
            elif tuple(self._file(self.where[name]).shape(name)) != tuple(shape):
                bad.append(f"{name}: {self._file(self.where[name]).shape(name)} != {tuple(shape)}")
        if missing or bad:
            raise ValueError(f"checkpoint does not match the config: {len(missing)} missing (first: {missing[:3]}), "
                             f"{len(bad)} shape mismatches (first: {bad[:3]})")

    def close(self):
        for f in self.files.values():
            f.close()
        self.files.clear()


def load_lora_file(path: str) -> Dict[str, torch.Tensor]:
    """`pytorch_lora_weights.safetensors` (or a directory holding it) -> {key: tensor} on the host (adapters are MBs)."""
    if os.path.isdir(path):
        path = os.path.join(path, LORA_WEIGHT_NAME_SAFE)
    with SafetensorsFile(path) as f:
        out = {k: f.get(k) for k in f.keys()}
    bad = [k for k in out if not (k.endswith(".lora_A.weight") or k.endswith(".lora_B.weight") or k.endswith(".alpha"))]
    if bad:
        raise ValueError(f"{path}: keys outside the diffusers/PEFT LoRA format (…lora_A.weight / …lora_B.weight / ….alpha): {bad[:3]}")
    for k in out:
        if k.endswith(".lora_A.weight") and (k[: -len("lora_A.weight")] + "lora_B.weight") not in out:
            raise ValueError(f"{path}: {k} has no matching lora_B")
    return out


def load_transformer(path: str, device="cuda", lora: Optional[str] = None, lora_scale: float = 1.0, config=None, **engine_kw):
    """B200FluxTransformer from a checkpoint directory / file, optionally with a LoRA file folded in at load."""
    from .engine import B200FluxTransformer, FrozenConfig, _cfg_dict
    ck = Checkpoint(path)
    cfg = config if config is not None else ck.config
    if cfg is None:
        raise ValueError(f"{path}: no {CONFIG_NAME} next to the weights; pass config=")
    cfg = FrozenConfig(_cfg_dict(cfg))  # same normalisation as the engine (out_channels None -> in_channels, ...)
    ck.check(cfg)
    get = ck.getter(device)
    if lora is not None:
        get = fold_lora(get, load_lora_file(lora), scale=lora_scale)
    try:
        return B200FluxTransformer(cfg, get, device=device, **engine_kw)
    finally:
        ck.close()


def _component(path: str, config):
    ck = Checkpoint(path)
    cfg = config if config is not None else ck.config
    if cfg is None:
        ck.close()
        raise ValueError(f"{path}: no {CONFIG_NAME} next to the weights; pass config=")
    return ck, cfg


def load_vae(path: str, device="cuda", config=None):
    """B200AutoencoderKL from `<pipeline>/vae` (AutoencoderKL.from_pretrained's files: config.json +
    diffusion_pytorch_model.safetensors)."""
    from .vae import B200AutoencoderKL
    ck, cfg = _component(path, config)
    try:
        return B200AutoencoderKL(cfg, ck.getter(device), names=ck.keys(), device=device)
    finally:
        ck.close()


def load_text_encoder(path: str, device="cuda", config=None, **kw):
    """B200CLIPTextEncoder / B200T5Encoder from `<pipeline>/text_encoder` / `<pipeline>/text_encoder_2` (transformers' files:
    config.json + model.safetensors or its shards); the class follows config.json's `model_type` / `architectures`."""
    from .text_encoders import B200CLIPTextEncoder, B200T5Encoder
    ck, cfg = _component(path, config)
    try:
        kind = str(cfg.get("model_type", "")) + " ".join(cfg.get("architectures") or [])
        if "t5" in kind.lower():
            return B200T5Encoder(cfg, ck.getter(device), device=device, **kw)
        if "clip" in kind.lower():
            return B200CLIPTextEncoder(cfg, ck.getter(device), device=device, **kw)
        raise ValueError(f"{path}: config.json names neither a T5 nor a CLIP text encoder ({kind!r})")
    finally:
        ck.close()


def load_components(root: str, device="cuda", lora: Optional[str] = None, lora_scale: float = 1.0, **engine_kw) -> Dict[str, object]:
    """Every device stage of a FluxFillPipeline directory (model_index.json's layout: transformer/, vae/, text_encoder/,
    text_encoder_2/, scheduler/scheduler_config.json) as engines; sub-directories that are absent are skipped.  The
    tokenizers stay with the caller (host-side string work, out of scope)."""
    out: Dict[str, object] = {}
    sub = lambda n: os.path.join(root, n)
    if os.path.isdir(sub("transformer")):
        out["transformer"] = load_transformer(sub("transformer"), device, lora=lora, lora_scale=lora_scale, **engine_kw)
    if os.path.isdir(sub("vae")):
        out["vae"] = load_vae(sub("vae"), device)
    for n in ("text_encoder", "text_encoder_2"):
        if os.path.isdir(sub(n)):
            out[n] = load_text_encoder(sub(n), device)
    sc = os.path.join(sub("scheduler"), "scheduler_config.json")
    if os.path.exists(sc):
        from .engine import B200FlowMatchEulerScheduler, B200StochasticRFOvershotScheduler
        with open(sc) as f:
            raw = json.load(f)
        cls = B200StochasticRFOvershotScheduler if "Overshot" in raw.get("_class_name", "") else B200FlowMatchEulerScheduler
        out["scheduler"] = cls.from_config(raw)  # refuses the sigma options the FLUX path never sets
    return out
