"""GPU parity (bit-exact) of the conditioning glue kernels against the oracle and the reference's golden outputs
(SURVEY.md §8f rank 2): FluxFillPipeline._pack_latents / _unpack_latents / prepare_mask_latents / the :2127 de-normalisation."""
import pytest
import torch

from oracle import flux_oracle as fo  # checker only

pytestmark = pytest.mark.gpu

SF, SC, VS = 0.1159, 0.3611, 8


def test_conditioning_matches_reference_golden(golden):
    from textflux_b200 import conditioning as cd
    d = golden("conditioning.pt")
    for c in d["cases"]:
        B, h, w, n_img = c["B"], c["h"], c["w"], c["num_images_per_prompt"]
        H, W = h * VS, w * VS
        sf, sc = d["shift_factor"], d["scaling_factor"]
        mil = c["masked_image_latents"].cuda()
        mp, lp = cd.prepare_mask_latents(c["mask"].cuda(), mil, B, 16, n_img, H, W, torch.bfloat16, "cuda", sf, sc, VS)
        assert torch.equal(mp.cpu(), c["mask_packed"])
        packed = cd.pack_latents(c["latents"].cuda(), B, 16, h, w)
        assert torch.equal(packed.cpu(), c["latents_packed"])
        assert torch.equal(cd.unpack_latents(packed, H, W, VS).cpu(), c["latents"])
        # The two scalar affine maps: the golden values were produced by the reference on the CPU, where eager add/sub
        # round the python scalar to bf16 first; CUDA eager (what the pipeline runs in production, and what the kernel
        # reproduces) keeps it in fp32.  So: bit-exact against the same reference expressions evaluated by CUDA eager,
        # and within two bf16 ulps (two rounded ops) plus the bf16 rounding of the shift scalar of the CPU golden.
        want_lp = fo.pack_latents(((mil - sf) * sc).repeat(n_img, 1, 1, 1))
        assert torch.equal(lp, want_lp)
        dec = cd.unscale_unpack_latents(packed, H, W, VS, sf, sc)
        assert torch.equal(dec, (fo.unpack_latents(packed, H, W, VS) / sc) + sf)
        for got, gold in ((lp.cpu(), c["masked_image_latents_packed"]), (dec.cpu(), c["decode_in"])):
            ulp = torch.maximum(gold.float().abs(), torch.tensor(2.0 ** -126)).log2().floor().exp2() * 2.0 ** -7
            assert ((got.float() - gold.float()).abs() <= 2 * ulp + 2.5e-4).all()  # |bf16(0.1159) - 0.1159| = 1.8e-4
        assert torch.equal(cd.prepare_latent_image_ids(B, h // 2, w // 2, "cuda", torch.bfloat16).cpu(), c["img_ids"])


# latent sizes of every BASELINE.json canvas plus ragged small ones
@pytest.mark.parametrize("B,h,w", [(1, 2, 2), (3, 6, 10), (1, 128, 64), (2, 144, 128), (1, 128, 128), (1, 256, 128), (1, 256, 192)])
def test_pack_unpack_mask_bit_exact_vs_oracle(B, h, w):
    from textflux_b200 import conditioning as cd
    g = torch.Generator().manual_seed(B * 100 + h + w)
    lat = torch.randn(B, 16, h, w, generator=g).to(torch.bfloat16)
    mil32 = torch.randn(B, 16, h, w, generator=g) * 2
    mask = (torch.rand(B, 1, h * VS, w * VS, generator=g) > 0.3).float()
    p = cd.pack_latents(lat.cuda(), B, 16, h, w)
    assert torch.equal(p.cpu(), fo.pack_latents(lat))
    assert torch.equal(cd.unpack_latents(p, h * VS, w * VS, VS).cpu(), lat)                      # round trip
    # `latents / scaling_factor + shift_factor` (:2127): the reference expression on CUDA eager is the checker (see above)
    assert torch.equal(cd.unscale_unpack_latents(p, h * VS, w * VS, VS, SF, SC), (fo.unpack_latents(p, h * VS, w * VS) / SC) + SF)
    for mil in (mil32, mil32.to(torch.bfloat16)):                                                  # fp32 and bf16 VAE outputs
        want_mask, want_mil = fo.prepare_mask_latents(mask, mil, B, 16, 1, h * VS, w * VS, torch.bfloat16, SF, SC, VS)
        cond = cd.pack_conditioning(mask.cuda(), mil.cuda(), h, w, SF, SC, VS)
        assert cond.shape == (B, (h // 2) * (w // 2), 320)
        assert torch.equal(cond[..., 64:].cpu(), want_mask)
        if mil.dtype == torch.float32:  # fp32 scalars on both devices: the CPU oracle is bit-exact
            assert torch.equal(cond[..., :64].cpu(), want_mil)
        assert torch.equal(cond[..., :64], fo.pack_latents(((mil.cuda() - SF) * SC).to(torch.bfloat16)))
    # one-hot pixels land in exactly one packed channel (Appendix B: channel (py*8+px)*4 + di*2 + dj of token (i, j))
    one = torch.zeros(1, 1, h * VS, w * VS)
    y, x = (h * VS) // 2 + 3, (w * VS) // 2 + 5
    one[0, 0, y, x] = 1
    pm = cd.pack_mask(one.cuda(), h, w, VS).cpu()
    i, di, py = (y // VS) // 2, (y // VS) % 2, y % VS
    j, dj, px = (x // VS) // 2, (x // VS) % 2, x % VS
    assert pm.sum() == 1 and pm[0, i * (w // 2) + j, (py * VS + px) * 4 + di * 2 + dj] == 1


def test_conditioning_errors_like_the_reference():
    from textflux_b200 import conditioning as cd
    lat = torch.zeros(1, 16, 4, 4, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        cd.pack_latents(lat, 1, 16, 4, 6)                       # shape mismatch
    with pytest.raises(ValueError):
        cd.pack_latents(lat.cpu(), 1, 16, 4, 4)                 # no CPU path
    with pytest.raises(ValueError):                             # 3 masks cannot be duplicated to batch 4 (pipeline_flux_fill.py:1541-1547)
        cd.prepare_mask_latents(torch.zeros(3, 1, 32, 32, device="cuda"), lat.expand(4, -1, -1, -1), 4, 16, 1, 32, 32, torch.bfloat16,
                                "cuda", SF, SC)
    odd = torch.zeros(1, 16, 3, 4, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        cd.pack_latents(odd, 1, 16, 3, 4)                       # odd latent height cannot be 2x2-packed
