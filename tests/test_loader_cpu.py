"""CPU: safetensors / LoRA loader -- file format against the `safetensors` library, sharded
checkpoints through the packer bit-exactly, LoRA fold against the manual W + (alpha/r) B A, error behaviour."""
import json
import os

import pytest
import torch

from oracle import flux_oracle as fo
from textflux_b200 import fold_lora, pack_weights
from textflux_b200 import loader as ld


def _tiny_sd():
    return fo.init_state_dict(fo.TINY, seed=77, dtype=torch.bfloat16)


def test_reader_and_writer_agree_with_the_safetensors_library(tmp_path):
    sd = _tiny_sd()
    sd["extra.f32"] = torch.randn(3, 5)
    sd["extra.empty"] = torch.empty(0, 4, dtype=torch.float16)
    mine = str(tmp_path / "mine.safetensors")
    ld.save_safetensors(sd, mine, metadata={"format": "pt"})
    with ld.SafetensorsFile(mine) as f:
        assert sorted(f.keys()) == sorted(sd) and f.metadata == {"format": "pt"}
        for k, v in sd.items():
            got = f.get(k)
            assert got.dtype == v.dtype and tuple(got.shape) == tuple(v.shape) and torch.equal(got, v)
    st = pytest.importorskip("safetensors.torch")
    theirs = str(tmp_path / "theirs.safetensors")
    st.save_file({k: v.contiguous() for k, v in sd.items()}, theirs)
    with ld.SafetensorsFile(theirs) as f:                      # our reader on the library's file
        for k, v in sd.items():
            assert torch.equal(f.get(k), v)
    back = st.load_file(mine)                                   # the library's reader on our file
    for k, v in sd.items():
        assert torch.equal(back[k], v)


def test_sharded_checkpoint_packs_bit_exactly(tmp_path):
    cfg, sd = fo.TINY, _tiny_sd()
    names = sorted(sd)
    shards = [names[0::3], names[1::3], names[2::3]]
    wm = {}
    for i, part in enumerate(shards):
        fn = f"diffusion_pytorch_model-{i + 1:05d}-of-00003.safetensors"
        ld.save_safetensors({k: sd[k] for k in part}, str(tmp_path / fn))
        wm.update({k: fn for k in part})
    json.dump({"metadata": {}, "weight_map": wm}, open(tmp_path / ld.SAFE_WEIGHTS_INDEX_NAME, "w"))
    json.dump(dict(cfg.to_dict(), _class_name="FluxTransformer2DModel"), open(tmp_path / ld.CONFIG_NAME, "w"))
    ck = ld.Checkpoint(str(tmp_path))
    assert ck.config["num_layers"] == cfg.num_layers
    ck.check(cfg)
    got = pack_weights(cfg, ck.getter("cpu"), "cpu")
    want = pack_weights(cfg, sd.__getitem__, "cpu")
    assert got.keys() == want.keys()
    for k in want:
        assert torch.equal(got[k], want[k]), k
    ck.close()
    # a missing tensor is reported like load_state_dict(strict=True) would
    os.remove(tmp_path / "diffusion_pytorch_model-00001-of-00003.safetensors")
    ld.save_safetensors({k: sd[k] for k in shards[0][1:]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00003.safetensors"))
    wm.pop(shards[0][0])
    json.dump({"metadata": {}, "weight_map": wm}, open(tmp_path / ld.SAFE_WEIGHTS_INDEX_NAME, "w"))
    with pytest.raises(ValueError, match="missing"):
        ld.Checkpoint(str(tmp_path)).check(cfg)


def test_lora_file_folds_like_the_manual_formula(tmp_path):
    cfg, sd = fo.TINY, _tiny_sd()
    D = cfg.inner_dim
    g = torch.Generator().manual_seed(5)
    r = 4
    lora = {}
    mods = ["transformer_blocks.0.attn.to_q", "transformer_blocks.1.ff.net.0.proj", "single_transformer_blocks.0.proj_mlp",
            "single_transformer_blocks.1.proj_out"]
    for m in mods:
        o, i = sd[m + ".weight"].shape
        lora[f"transformer.{m}.lora_A.weight"] = (torch.randn(r, i, generator=g) * 0.05).to(torch.bfloat16)
        lora[f"transformer.{m}.lora_B.weight"] = (torch.randn(o, r, generator=g) * 0.05).to(torch.bfloat16)
    lora[f"transformer.{mods[1]}.alpha"] = torch.tensor(8.0)
    (tmp_path / "lora").mkdir()
    ld.save_safetensors(lora, str(tmp_path / "lora" / ld.LORA_WEIGHT_NAME_SAFE))
    got = ld.load_lora_file(str(tmp_path / "lora"))
    assert got.keys() == lora.keys()
    get = fold_lora(sd.__getitem__, got, scale=0.5)
    for m in mods:
        A, Bm = lora[f"transformer.{m}.lora_A.weight"].float(), lora[f"transformer.{m}.lora_B.weight"].float()
        alpha = 8.0 if m == mods[1] else float(r)
        want = (sd[m + ".weight"].float() + 0.5 * alpha / r * (Bm @ A)).to(torch.bfloat16)
        assert torch.equal(get(m + ".weight"), want)
    assert torch.equal(get("transformer_blocks.0.attn.to_k.weight"), sd["transformer_blocks.0.attn.to_k.weight"])  # untouched
    assert D == sd["transformer_blocks.0.attn.to_q.weight"].shape[0]
    bad = dict(lora)
    bad.pop(f"transformer.{mods[0]}.lora_B.weight")
    ld.save_safetensors(bad, str(tmp_path / "bad.safetensors"))
    with pytest.raises(ValueError, match="lora_B"):
        ld.load_lora_file(str(tmp_path / "bad.safetensors"))


def test_corrupt_files_are_rejected(tmp_path):
    p = tmp_path / "short.safetensors"
    p.write_bytes(b"\x01\x02")
    with pytest.raises(ValueError):
        ld.SafetensorsFile(str(p))
    p = tmp_path / "lying.safetensors"
    h = json.dumps({"w": {"dtype": "BF16", "shape": [4, 4], "data_offsets": [0, 8]}}).encode()
    p.write_bytes(len(h).to_bytes(8, "little") + h + b"\0" * 8)
    with pytest.raises(ValueError, match="inconsistent"):
        ld.SafetensorsFile(str(p))
    with pytest.raises(FileNotFoundError):
        ld.Checkpoint(str(tmp_path))


def test_lora_hot_swap_repacks_only_touched_matrices_bit_exactly():
    """repack_modules(base + adapter) on an already packed model == pack_weights(fold) from scratch; swapping to a second
    adapter and unloading restore exactly what a fresh pack would hold (hot-swap without re-packing)."""
    from textflux_b200 import lora_modules, repack_modules
    cfg, sd = fo.TINY, _tiny_sd()
    g = torch.Generator().manual_seed(11)

    def adapter(mods, r=4):
        out = {}
        for m in mods:
            o, i = sd[m + ".weight"].shape
            out[f"transformer.{m}.lora_A.weight"] = (torch.randn(r, i, generator=g) * 0.05).to(torch.bfloat16)
            out[f"transformer.{m}.lora_B.weight"] = (torch.randn(o, r, generator=g) * 0.05).to(torch.bfloat16)
        return out

    l1 = adapter(["transformer_blocks.0.attn.to_k", "single_transformer_blocks.1.proj_mlp", "transformer_blocks.1.ff.net.2"])
    l2 = adapter(["transformer_blocks.0.attn.to_q", "single_transformer_blocks.0.attn.to_v"])
    base = pack_weights(cfg, sd.__getitem__, "cpu")
    P = {k: v.clone() for k, v in base.items()}
    ptrs = {k: v.data_ptr() for k, v in P.items()}
    touched = repack_modules(cfg, fold_lora(sd.__getitem__, l1, scale=0.8), P, lora_modules(l1))
    assert sorted(touched) == ["d0.qkv_x", "d1.ff2_x", "s1.qkvmlp"]
    want = pack_weights(cfg, fold_lora(sd.__getitem__, l1, scale=0.8), "cpu")
    assert all(torch.equal(P[k], want[k]) for k in want) and {k: v.data_ptr() for k, v in P.items()} == ptrs
    # swap: modules of the old adapter go back to base, the new ones are folded
    repack_modules(cfg, fold_lora(sd.__getitem__, l2), P, set(lora_modules(l1)) | set(lora_modules(l2)))
    want = pack_weights(cfg, fold_lora(sd.__getitem__, l2), "cpu")
    assert all(torch.equal(P[k], want[k]) for k in want)
    repack_modules(cfg, sd.__getitem__, P, lora_modules(l2))
    assert all(torch.equal(P[k], base[k]) for k in base)


def test_checkpoint_resolves_transformers_file_names(tmp_path):
    """Text-encoder directories use transformers' names (model.safetensors / model.safetensors.index.json)."""
    import json
    from textflux_b200 import loader as ld
    g = torch.Generator().manual_seed(3)
    sd = {f"encoder.block.{i}.w": torch.randn(4, 8, generator=g).to(torch.bfloat16) for i in range(4)}
    one = tmp_path / "one"; one.mkdir()
    ld.save_safetensors(sd, str(one / ld.TRANSFORMERS_WEIGHTS_NAME))
    json.dump({"model_type": "t5"}, open(one / ld.CONFIG_NAME, "w"))
    ck = ld.Checkpoint(str(one))
    assert sorted(ck.keys()) == sorted(sd) and ck.config == {"model_type": "t5"}
    assert all(torch.equal(ck.getter()(k), v) for k, v in sd.items())
    ck.close()
    two = tmp_path / "two"; two.mkdir()
    names = sorted(sd)
    wm = {}
    for i, part in enumerate([names[:1], names[1:]]):
        fn = f"model-{i + 1:05d}-of-00002.safetensors"
        ld.save_safetensors({k: sd[k] for k in part}, str(two / fn))
        wm.update({k: fn for k in part})
    json.dump({"metadata": {}, "weight_map": wm}, open(two / ld.TRANSFORMERS_INDEX_NAME, "w"))
    ck = ld.Checkpoint(str(two))
    assert all(torch.equal(ck.getter()(k), v) for k, v in sd.items())
    ck.close()
    empty = tmp_path / "empty"; empty.mkdir()
    with pytest.raises(FileNotFoundError):
        ld.Checkpoint(str(empty))
    with pytest.raises(ValueError, match="config"):
        ld.load_text_encoder(str(two))  # weights but no config.json: refused before any device work


def test_load_components_scheduler_only_directory(tmp_path):
    """A pipeline directory without model sub-directories yields just the scheduler, built through from_config (which refuses the
    sigma options the FLUX path never sets) -- no device is touched."""
    import json
    from textflux_b200 import B200FlowMatchEulerScheduler, B200StochasticRFOvershotScheduler
    from textflux_b200 import loader as ld
    (tmp_path / "scheduler").mkdir()
    conf = {"_class_name": "FlowMatchEulerDiscreteScheduler", "_diffusers_version": "0.32.0", "shift": 3.0, "use_dynamic_shifting": True,
            "base_shift": 0.5, "max_shift": 1.15, "base_image_seq_len": 256, "max_image_seq_len": 4096, "num_train_timesteps": 1000,
            "a_key_from_a_newer_diffusers": None}
    json.dump(conf, open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    parts = ld.load_components(str(tmp_path), device="cpu")
    assert set(parts) == {"scheduler"} and type(parts["scheduler"]) is B200FlowMatchEulerScheduler
    assert parts["scheduler"].config["shift"] == 3.0 and parts["scheduler"].config["max_image_seq_len"] == 4096
    json.dump(dict(conf, _class_name="StochasticRFOvershotDiscreteScheduler"), open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    assert type(ld.load_components(str(tmp_path))["scheduler"]) is B200StochasticRFOvershotScheduler
    json.dump(dict(conf, use_karras_sigmas=True), open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    with pytest.raises(ValueError, match="use_karras_sigmas"):
        ld.load_components(str(tmp_path))
