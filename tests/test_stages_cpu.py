"""CPU: host-side logic of the once-per-image stages -- weight layouts and name tables, no GPU needed."""
import pytest
import torch
import torch.nn.functional as F

from oracle import textenc_oracle as to
from oracle import vae_oracle as vo


def test_packed_conv_weight_is_the_implicit_gemm_operand():
    """pack_vae_weights lays a 3x3 weight out as [Cout, 9 * Cin64] with column (ky * 3 + kx) * Cin64 + c (tfx_vae_set_weight): a GEMM of
    the 9 shifted NHWC views against it must equal F.conv2d, which is what the convolution mode of the GEMM kernel computes."""
    from textflux_b200.vae import pack_vae_weights
    g = torch.Generator().manual_seed(0)
    for ci, co in ((3, 8), (16, 8), (64, 16), (96, 8)):
        w = torch.randn(co, ci, 3, 3, generator=g)
        b = torch.randn(co, generator=g)
        x = torch.randn(2, ci, 6, 7, generator=g)
        P = pack_vae_weights({"c.weight": w, "c.bias": b}.__getitem__, ["c.weight", "c.bias"], "cpu", dtype=torch.float32)
        cip = (ci + 63) // 64 * 64
        assert P["c.weight"].shape == (co, 9 * cip) and P["c.bias"].shape == (1, co)
        xn = torch.zeros(2, 6 + 2, 7 + 2, cip)
        xn[:, 1:-1, 1:-1, :ci] = x.permute(0, 2, 3, 1)  # NHWC, channels padded to 64, spatial zero padding = TMA out-of-bounds fill
        cols = torch.cat([xn[:, ky:ky + 6, kx:kx + 7, :] for ky in range(3) for kx in range(3)], dim=-1)  # [B, H, W, 9 * Cin64]
        out = (cols.reshape(-1, 9 * cip) @ P["c.weight"].T + P["c.bias"]).reshape(2, 6, 7, co).permute(0, 3, 1, 2)
        assert torch.allclose(out, F.conv2d(x, w, b, padding=1), atol=1e-4)
    # 1x1 convolutions and vectors
    P = pack_vae_weights({"s.weight": torch.ones(4, 6, 1, 1), "n.weight": torch.ones(5)}.__getitem__, ["s.weight", "n.weight"], "cpu")
    assert P["s.weight"].shape == (4, 6) and P["n.weight"].shape == (1, 5)
    with pytest.raises(ValueError):
        pack_vae_weights({"bad": torch.ones(2, 2, 5, 5)}.__getitem__, ["bad"], "cpu")


def test_stride2_parity_view_reads_the_pixels_downsample2d_reads():
    """Downsample2D pads (0,1,0,1) and strides 2 (downsampling.py:141-147): output (y, x), tap (ky, kx) reads input (2y + ky, 2x + kx).
    The kernel addresses the image as [C, 2, W/2, 2, H/2]: parity (ky & 1, kx & 1), pair (y + (ky >> 1), x + (kx >> 1))."""
    H, W = 6, 8
    img = torch.arange(H * W).reshape(H, W)
    view = img.reshape(H // 2, 2, W // 2, 2)  # [y pair, py, x pair, px]
    for ky in range(3):
        for kx in range(3):
            for y in range(H // 2):
                for x in range(W // 2):
                    yy, xx = y + (ky >> 1), x + (kx >> 1)
                    got = int(view[yy, ky & 1, xx, kx & 1]) if yy < H // 2 and xx < W // 2 else None  # None = out of bounds = zero
                    iy, ix = 2 * y + ky, 2 * x + kx
                    want = int(img[iy, ix]) if iy < H and ix < W else None
                    assert got == want


def test_product_name_tables_equal_the_oracle_specs():
    from textflux_b200.text_encoders import CLIP_L_CONFIG, T5_XXL_CONFIG, clip_reference_names, t5_reference_names
    from textflux_b200.vae import FLUX_VAE_CONFIG, vae_reference_names
    assert sorted(vae_reference_names(FLUX_VAE_CONFIG)) == sorted((n, s) for n, s, _ in vo.state_dict_spec(vo.FLUX_VAE))
    assert sorted(t5_reference_names(T5_XXL_CONFIG)) == sorted((n, s) for n, s, _ in to.t5_spec(to.T5_XXL))
    assert sorted(clip_reference_names(CLIP_L_CONFIG)) == sorted((n, s) for n, s, _ in to.clip_spec(to.CLIP_L))
    # FLUX's VAE: 83.8 M parameters; T5-XXL encoder 4.76 B (incl. the shared embedding); CLIP-L text tower 123 M
    count = lambda names: sum(torch.Size(s).numel() for _, s in names)
    assert abs(count(vae_reference_names(FLUX_VAE_CONFIG)) - 83.8e6) < 0.2e6
    assert abs(count(t5_reference_names(T5_XXL_CONFIG)) - 4.762e9) < 0.01e9
    assert abs(count(clip_reference_names(CLIP_L_CONFIG)) - 123.06e6) < 0.1e6


def test_mirrors_refuse_cpu_devices_without_touching_the_library_state():
    """No CPU path anywhere: the mirrors raise before any weight is packed."""
    from textflux_b200 import B200AutoencoderKL, B200CLIPTextEncoder, B200T5Encoder
    from textflux_b200.text_encoders import CLIP_L_CONFIG, T5_XXL_CONFIG
    from textflux_b200.vae import FLUX_VAE_CONFIG
    with pytest.raises(RuntimeError):
        B200AutoencoderKL(FLUX_VAE_CONFIG, {}.__getitem__, names=[], device="cpu")
    with pytest.raises(RuntimeError):
        B200T5Encoder(T5_XXL_CONFIG, {}.__getitem__, device="cpu")
    with pytest.raises(RuntimeError):
        B200CLIPTextEncoder(CLIP_L_CONFIG, {}.__getitem__, device="cpu")
