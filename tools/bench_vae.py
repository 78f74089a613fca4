"""AutoencoderKL encode / decode at the benchmark canvases: the engine (tfx_vae_*) against the reference's CUDA-eager ops (the oracle
restatement on CUDA tensors: F.conv2d -> cuDNN, F.group_norm, SDPA) on the same box.  CUDA events, median of `iters`.
Usage: python tools/bench_vae.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vae_oracle as vo  # noqa: E402  (the CUDA-eager baseline leg; never on the engine's path)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def conv_flops(cfg, H, W, decode):
    """Multiply-add = 2 FLOPs over every convolution / linear / attention matmul of one image (algorithmic: unpadded channels)."""
    f = 0
    for name, shape, kind in vo.state_dict_spec(cfg):
        if kind != "w" or not name.endswith(".weight") or name.startswith("encoder.") == decode:
            continue
        n = len(cfg.block_out_channels)
        # resolution of the layer's OUTPUT
        if name.startswith("encoder."):
            if ".down_blocks." in name:
                i = int(name.split(".")[2])
                s = 2 ** i * (2 if "downsamplers" in name else 1)
            elif "conv_in" in name:
                s = 1
            else:
                s = 2 ** (n - 1)
        else:
            if ".up_blocks." in name:
                i = int(name.split(".")[2])
                s = 2 ** (n - 1 - i) // (2 if "upsamplers" in name else 1)
            elif "conv_out" in name:
                s = 1
            else:
                s = 2 ** (n - 1)
        px = (H // s) * (W // s)
        k = 1
        for d in shape[1:]:
            k *= d
        f += 2 * px * shape[0] * k
    if cfg.mid_block_add_attention:
        N, C = (H // 2 ** (len(cfg.block_out_channels) - 1)) * (W // 2 ** (len(cfg.block_out_channels) - 1)), cfg.block_out_channels[-1]
        f += 4 * N * N * C
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--sizes", default="1024x1152,1024x2048,512x1024")
    args = ap.parse_args()
    from textflux_b200.vae import B200AutoencoderKL
    cfg = vo.FLUX_VAE
    sd32 = vo.init_state_dict(cfg, seed=31, device="cuda")
    sd16 = {k: v.to(torch.bfloat16) for k, v in sd32.items()}
    vae = B200AutoencoderKL.from_state_dict(cfg.reference_kwargs(), sd32, device="cuda:0")
    res = []
    for size in args.sizes.split(","):
        H, W = (int(v) for v in size.split("x"))
        g = torch.Generator(device="cuda").manual_seed(H + W)
        image = (torch.rand(1, 3, H, W, generator=g, device="cuda") * 2 - 1).to(torch.bfloat16)
        z = torch.randn(1, 16, H // 8, W // 8, generator=g, device="cuda").to(torch.bfloat16)
        row = {"H": H, "W": W}
        for what, ours, ref, fl in (("encode", lambda: vae.encode(image), lambda: vo.encode_moments(sd16, cfg, image), conv_flops(cfg, H, W, False)),
                                    ("decode", lambda: vae.decode(z), lambda: vo.decode(sd16, cfg, z), conv_flops(cfg, H, W, True))):
            l0 = vae.counter("launches")
            ours()
            row[f"{what}_launches"] = vae.counter("launches") - l0
            ms = timeit(ours)
            row[f"{what}_ms"] = ms
            row[f"{what}_tflops"] = fl / ms / 1e9
            try:
                ms_ref = timeit(ref, iters=3, warm=1)
                row[f"{what}_cuda_eager_ms"] = ms_ref
            except torch.cuda.OutOfMemoryError:
                row[f"{what}_cuda_eager_ms"] = None
            row[f"{what}_tflop"] = fl / 1e12
        print(row, flush=True)
        res.append(row)
        torch.cuda.empty_cache()
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
