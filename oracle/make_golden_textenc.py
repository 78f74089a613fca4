"""Generate tests/golden/textenc_{t5,clip}.pt by running the INSTALLED transformers T5EncoderModel / CLIPTextModel (the third-party
implementation the reference calls, pipeline_flux_fill.py:1438,1483) in the build container, eager attention.

    python -m oracle.make_golden_textenc

TEST INFRASTRUCTURE ONLY.  Weights are regenerated from oracle.textenc_oracle.init_state_dict (seeded); the fixtures hold the token
ids and the library's outputs in fp32 and bf16, and the generator asserts the oracle restates them bit-exactly."""
from __future__ import annotations

import os

import torch

from . import textenc_oracle as to

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@torch.no_grad()
def t5_case(cfg: to.T5Cfg, seed: int, B: int, T: int):
    import transformers
    from transformers import T5Config, T5EncoderModel
    hf = T5Config(vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.num_layers, num_heads=cfg.num_heads,
                  relative_attention_num_buckets=cfg.relative_attention_num_buckets, relative_attention_max_distance=cfg.relative_attention_max_distance,
                  layer_norm_epsilon=cfg.layer_norm_epsilon, feed_forward_proj="gated-gelu", dropout_rate=0.0, tie_word_embeddings=False,
                  is_encoder_decoder=False, use_cache=False)
    sd32 = to.init_state_dict(to.t5_spec(cfg), seed)
    g = torch.Generator().manual_seed(seed + 1)
    ids = torch.randint(2, cfg.vocab_size, (B, T), generator=g)
    ids[:, T - T // 3:] = 0  # padding="max_length": pad id 0 tail, attended like any token (no attention mask is passed)
    ids[:, T - T // 3 - 1] = 1  # </s>
    out = {"config": cfg.to_dict(), "seed": seed, "input_ids": ids, "transformers": transformers.__version__}
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        m = T5EncoderModel(hf)
        full = dict(sd32)
        full["encoder.embed_tokens.weight"] = sd32["shared.weight"]
        missing, unexpected = m.load_state_dict(full, strict=True)
        assert not missing and not unexpected
        m = m.to(dtype).eval()
        y = m(ids, output_hidden_states=False)[0]
        out[f"last_hidden_{tag}"] = y.clone()
        sd = {k: v.to(dtype) for k, v in sd32.items()}
        mine = to.t5_encode(sd, cfg, ids)
        assert torch.equal(mine, y), (tag, (mine.float() - y.float()).abs().max())
    return out


@torch.no_grad()
def clip_case(cfg: to.ClipCfg, seed: int, B: int, T: int):
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    hf = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                        max_position_embeddings=cfg.max_position_embeddings, hidden_act="quick_gelu", layer_norm_eps=cfg.layer_norm_eps,
                        attention_dropout=0.0, eos_token_id=cfg.eos_token_id, bos_token_id=0, pad_token_id=1, projection_dim=cfg.hidden_size)
    hf._attn_implementation = "eager"
    sd32 = to.init_state_dict(to.clip_spec(cfg), seed)
    g = torch.Generator().manual_seed(seed + 1)
    ids = torch.randint(3, cfg.vocab_size - 2, (B, T), generator=g)
    ids[:, 0] = cfg.vocab_size - 2  # <|startoftext|>
    for b in range(B):
        e = T // 2 + 5 * b
        ids[b, e] = cfg.vocab_size - 1  # <|endoftext|>: the largest id, which the argmax pooling rule finds
        ids[b, e + 1:] = cfg.vocab_size - 1 if b % 2 else 3  # CLIP pads with EOS (openai) -- or another token: argmax still takes the first EOS
    out = {"config": cfg.to_dict(), "seed": seed, "input_ids": ids, "transformers": transformers.__version__}
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        m = CLIPTextModel(hf)
        missing, unexpected = m.load_state_dict(sd32, strict=True)
        assert not missing and not unexpected
        m = m.to(dtype).eval()
        y = m(ids, output_hidden_states=False)
        out[f"last_hidden_{tag}"], out[f"pooled_{tag}"] = y.last_hidden_state.clone(), y.pooler_output.clone()
        sd = {k: v.to(dtype) for k, v in sd32.items()}
        lh, po = to.clip_encode(sd, cfg, ids)
        assert torch.equal(lh, y.last_hidden_state), (tag, (lh.float() - y.last_hidden_state.float()).abs().max())
        assert torch.equal(po, y.pooler_output), tag
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save(t5_case(to.T5_TINY, 91, 2, 48), os.path.join(GOLDEN, "textenc_t5.pt"))
    torch.save(clip_case(to.CLIP_TINY, 92, 2, 77), os.path.join(GOLDEN, "textenc_clip.pt"))
    for n in ("textenc_t5.pt", "textenc_clip.pt"):
        print(n, os.path.getsize(os.path.join(GOLDEN, n)))


if __name__ == "__main__":
    main()
