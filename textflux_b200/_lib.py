"""ctypes binding of include/textflux_b200.h.  There is no CPU or PyTorch fallback: if the CUDA library is missing
or fails to load, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH


class TfxVaeConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("latent_channels", C.c_int32), ("num_blocks", C.c_int32),
                ("block_out_channels", C.c_int32 * 8), ("layers_per_block", C.c_int32), ("norm_num_groups", C.c_int32),
                ("mid_block_add_attention", C.c_int32)]


class TfxTextEncConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("vocab_size", C.c_int32), ("d_model", C.c_int32), ("d_kv", C.c_int32), ("num_heads", C.c_int32),
                ("num_layers", C.c_int32), ("d_ff", C.c_int32), ("max_positions", C.c_int32), ("rel_buckets", C.c_int32),
                ("rel_max_distance", C.c_int32), ("eps", C.c_float)]


class TfxConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("num_layers", C.c_int32),
                ("num_single_layers", C.c_int32), ("attention_head_dim", C.c_int32),
                ("num_attention_heads", C.c_int32), ("joint_attention_dim", C.c_int32),
                ("pooled_projection_dim", C.c_int32), ("guidance_embeds", C.c_int32),
                ("axes_dims_rope", C.c_int32 * 3)]


_P, _I32, _I64, _U32, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float

#: every symbol include/textflux_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tfx_create": (C.c_int, [C.POINTER(TfxConfig), _I32, C.POINTER(_P)]),
    "tfx_destroy": (None, [_P]),
    "tfx_last_error": (C.c_char_p, [_P]),
    "tfx_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "tfx_get_counter": (C.c_int, [_P, C.c_char_p, C.POINTER(_I64)]),
    "tfx_set_weight": (C.c_int, [_P, C.c_char_p, _P, _I64, _I64]),
    "tfx_unset_weight": (C.c_int, [_P, C.c_char_p]),
    "tfx_finalize_weights": (C.c_int, [_P]),
    "tfx_prepare": (C.c_int, [_P, _I32, _I32, _I32]),
    "tfx_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tfx_euler_step": (C.c_int, [_P, _P, _P, _I64, _F, _F, _P]),
    "tfx_overshoot_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _P]),
    "tfx_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P]),
    "tfx_set_schedule": (C.c_int, [_P, _P, _I32, _P, _P, _P]),
    "tfx_step_scheduled": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P]),
    "tfx_op_linear": (C.c_int, [_P, _I64, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _P]),
    "tfx_op_linear_lora": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P]),
    "tfx_op_linear_qkv": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "tfx_op_linear_euler": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "tfx_op_attention": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "tfx_op_ln_modulate": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _I64, _I64, _I64, _P]),
    "tfx_op_gemv": (C.c_int, [_P, _I32, _I32, _P, _P, _I64, _P, _I32, _P]),
    "tfx_op_rope_table": (C.c_int, [_P, _P, _I32, _I32, C.POINTER(_I32), _P, _P]),
    "tfx_op_timestep_embed": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "tfx_debug_set_attention_trace": (C.c_int, [_P]),
    "tfx_debug_set_attention_cta_trace": (C.c_int, [_P]),
    "tfx_op_pack_latents": (C.c_int, [_P, _I32, _P, _I64, _I64, _I32, _I32, _I32, _I32, _I32, C.c_float, C.c_float, _P]),
    "tfx_op_unpack_latents": (C.c_int, [_P, _I64, _P, _I32, _I32, _I32, _I32, _I32, C.c_float, C.c_float, _P]),
    "tfx_op_pack_mask": (C.c_int, [_P, _I32, _P, _I64, _I64, _I32, _I32, _I32, _I32, _P]),
    "tfx_vae_create": (C.c_int, [C.POINTER(TfxVaeConfig), _I32, C.POINTER(_P)]),
    "tfx_vae_destroy": (None, [_P]),
    "tfx_vae_last_error": (C.c_char_p, [_P]),
    "tfx_vae_get_counter": (C.c_int, [_P, C.c_char_p, C.POINTER(_I64)]),
    "tfx_vae_set_weight": (C.c_int, [_P, C.c_char_p, _P, _I64, _I64]),
    "tfx_vae_encode": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "tfx_vae_decode": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "tfx_op_gaussian_sample": (C.c_int, [_P, _P, _P, _I32, _I32, _I64, _P]),
    "tfx_textenc_create": (C.c_int, [C.POINTER(TfxTextEncConfig), _I32, C.POINTER(_P)]),
    "tfx_textenc_destroy": (None, [_P]),
    "tfx_textenc_last_error": (C.c_char_p, [_P]),
    "tfx_textenc_get_counter": (C.c_int, [_P, C.c_char_p, C.POINTER(_I64)]),
    "tfx_textenc_set_weight": (C.c_int, [_P, C.c_char_p, _P, _I64, _I64]),
    "tfx_textenc_encode": (C.c_int, [_P, _P, _I32, _I32, _P, _P, _P, _P, _P]),
    "tfx_op_umma_probe": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _U32, _U32, _U32, _P]),
}

_lib = None


def load(path: str | None = None):
    """Load libtextflux_b200.so (built in-tree by textflux_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("TEXTFLUX_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(f"textflux_b200: CUDA library not found at {path}; run `python -c 'import __graft_entry__ as g; "
                           f"g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TfxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"textflux_b200 error {code}: {msg}")
        self.code = code


def check(code: int, handle=None, vae: bool = False, textenc: bool = False):
    if code != 0:
        lib = load()
        msg = lib.tfx_vae_last_error(handle) if vae else lib.tfx_textenc_last_error(handle) if textenc else lib.tfx_last_error(handle)
        text = msg.decode() if msg else "unknown error"
        if code == 1:
            raise ValueError(f"textflux_b200: {text}")  # the reference raises ValueError on bad inputs
        raise TfxError(code, text)
