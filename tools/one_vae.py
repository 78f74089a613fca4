"""One vae.decode (and one vae.encode) of the headline canvas between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=1024)
    ap.add_argument("--W", type=int, default=1152)
    ap.add_argument("--what", default="decode", choices=["decode", "encode", "t5"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    if args.what == "t5":
        from textflux_b200 import B200T5Encoder
        from textflux_b200.text_encoders import T5_XXL_CONFIG, t5_reference_names
        from textflux_b200.vae import synthetic_state
        enc = B200T5Encoder(T5_XXL_CONFIG, synthetic_state(t5_reference_names(T5_XXL_CONFIG), 7, dev).__getitem__, device=dev)
        ids = torch.randint(2, 32128, (1, 512), device=dev)
        fn = lambda: enc(ids)
    else:
        from textflux_b200 import B200AutoencoderKL
        from textflux_b200.vae import FLUX_VAE_CONFIG, synthetic_state, vae_reference_names
        vae = B200AutoencoderKL.from_state_dict(FLUX_VAE_CONFIG, synthetic_state(vae_reference_names(FLUX_VAE_CONFIG), 31, dev), device=dev)
        g = torch.Generator(device=dev).manual_seed(5)
        image = (torch.rand(1, 3, args.H, args.W, generator=g, device=dev) * 2 - 1).to(torch.bfloat16)
        z = torch.randn(1, 16, args.H // 8, args.W // 8, generator=g, device=dev).to(torch.bfloat16)
        fn = (lambda: vae.decode(z)) if args.what == "decode" else (lambda: vae.encode(image))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", args.what)


if __name__ == "__main__":
    main()
