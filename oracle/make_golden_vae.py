"""Generate tests/golden/vae_{small,flux}.pt by running the REAL reference AutoencoderKL (read-only import) in the build container.

    python -m oracle.make_golden_vae

TEST INFRASTRUCTURE ONLY.  Weights are regenerated from oracle.vae_oracle.init_state_dict (seeded); the fixtures hold inputs and the
reference's outputs (fp32 and bf16): moments of `encode`, `latent_dist.sample(generator)`, `decode`."""
from __future__ import annotations

import os

import torch

from . import vae_oracle as vo
from .ref_loader import import_reference

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference_vae(cfg: vo.VaeConfig, sd, dtype):
    d = import_reference()
    m = d.AutoencoderKL(**cfg.reference_kwargs())
    missing, unexpected = m.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()}, strict=True)
    assert not missing and not unexpected
    return m.to(dtype).eval()


@torch.no_grad()
def case(cfg: vo.VaeConfig, seed: int, B: int, H: int, W: int):
    sd32 = vo.init_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    image = torch.rand(B, cfg.in_channels, H, W, generator=g) * 2 - 1
    f = 2 ** (len(cfg.block_out_channels) - 1)
    z = torch.randn(B, cfg.latent_channels, H // f, W // f, generator=g)
    out = {"config": cfg.to_dict(), "seed": seed, "image": image, "z": z}
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        m = build_reference_vae(cfg, sd32, dtype)
        post = m.encode(image.to(dtype)).latent_dist
        out[f"moments_{tag}"] = post.parameters.clone()
        out[f"sample_{tag}"] = post.sample(generator=torch.Generator().manual_seed(5)).clone()
        out[f"mode_{tag}"] = post.mode().clone()
        out[f"decoded_{tag}"] = m.decode(z.to(dtype), return_dict=False)[0].clone()
        # the oracle must restate the reference exactly
        sd = {k: v.to(dtype) for k, v in sd32.items()}
        assert torch.equal(vo.encode_moments(sd, cfg, image.to(dtype)), out[f"moments_{tag}"]), (tag, "encode")
        assert torch.equal(vo.decode(sd, cfg, z.to(dtype)), out[f"decoded_{tag}"]), (tag, "decode")
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save(case(vo.SMALL_VAE, 77, 2, 32, 48), os.path.join(GOLDEN, "vae_small.pt"))
    torch.save(case(vo.FLUX_VAE, 78, 1, 64, 96), os.path.join(GOLDEN, "vae_flux.pt"))
    for n in ("vae_small.pt", "vae_flux.pt"):
        print(n, os.path.getsize(os.path.join(GOLDEN, n)))


if __name__ == "__main__":
    main()
