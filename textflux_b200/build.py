"""In-tree nvcc build of libtextflux_b200.so for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtextflux_b200.so")
SOURCES = ["tfx_api.cu"]
HEADERS = ["ptx.cuh", "gemm.cuh", "attention_common.cuh", "attention3.cuh", "attention4.cuh", "attention5.cuh", "conditioning.cuh",
           "pointwise.cuh", "probe.cuh", "vae.cuh", "vae_host.inl", "textenc.cuh", "textenc_host.inl"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; textflux_b200 needs the CUDA 12.9+ toolkit to build for sm_100a")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "textflux_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd[1:1] = os.environ.get("TFX_NVCC_FLAGS", "").split()  # experiments: e.g. -DTFX_ATTN8 (tools/experiments)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
