"""Import the REAL reference (vendored diffusers of yyyyyxie/textflux) read-only, build container only.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; everything that must
run there uses oracle/flux_oracle.py (pinned to this reference by tests/golden/, see make_golden.py).
"""
from __future__ import annotations

import os
import sys

REF_SRC = os.environ.get("TEXTFLUX_REFERENCE_SRC", "/root/reference/diffusers/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "diffusers"))


def import_reference():
    """Returns the `diffusers` module of the reference tree (transformers>=5 needs one shim)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_SRC}")
    import transformers.utils as tu
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import diffusers  # noqa
    assert diffusers.__file__.startswith(REF_SRC), diffusers.__file__
    return diffusers


def build_reference_transformer(cfg, state_dict, dtype):
    """Reference FluxTransformer2DModel carrying `state_dict` (names are the reference's own)."""
    import torch
    d = import_reference()
    with torch.device("meta"):
        m = d.FluxTransformer2DModel(**cfg.to_dict())
    m = m.to_empty(device="cpu").to(dtype)
    missing, unexpected = m.load_state_dict({k: v.to(dtype) for k, v in state_dict.items()}, strict=True)
    assert not missing and not unexpected
    return m.eval()


def build_reference_scheduler():
    d = import_reference()
    return d.FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15)
