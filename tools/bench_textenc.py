"""Prompt encoders at production size: T5-XXL encoder (24 layers, d_model 4096, 4.76 B parameters, 512 tokens) and CLIP-L text tower
(12 layers, 77 tokens), engine (tfx_textenc_*) against the reference's CUDA-eager ops (the oracle restatement on CUDA tensors =
what transformers' eager modules dispatch) on the same box.  CUDA events, median.  Usage: python tools/bench_textenc.py [--json out]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import textenc_oracle as to  # noqa: E402  (baseline leg only)


def timeit(fn, iters=7, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--t5-layers", type=int, default=24)
    args = ap.parse_args()
    from textflux_b200 import B200CLIPTextEncoder, B200T5Encoder
    res = {}
    # ---- T5-XXL
    cfg = to.T5Cfg(num_layers=args.t5_layers)
    sd = to.init_state_dict(to.t5_spec(cfg), 7, dtype=torch.bfloat16, device="cuda")
    enc = B200T5Encoder(dict(cfg.to_dict(), feed_forward_proj="gated-gelu"), sd.__getitem__, device="cuda:0")
    ids = torch.randint(2, cfg.vocab_size, (1, 512), device="cuda")
    ids[:, 60:] = 0
    inner = cfg.num_heads * cfg.d_kv
    params = sum(v.numel() for k, v in sd.items() if k != "shared.weight")
    flops = 2 * 512 * (4 * inner * cfg.d_model + 3 * cfg.d_ff * cfg.d_model) * cfg.num_layers + 4 * 512 * 512 * inner * cfg.num_layers
    l0 = enc.counter("launches")
    enc(ids)
    n_launch = enc.counter("launches") - l0
    ms = timeit(lambda: enc(ids))
    ms_ref = timeit(lambda: to.t5_encode(sd, cfg, ids), iters=3, warm=1)
    res["t5_xxl"] = {"layers": cfg.num_layers, "tokens": 512, "params_b": params / 1e9, "weight_gb": params * 2 / 1e9, "tflop": flops / 1e12,
                     "engine_ms": ms, "engine_tflops": flops / ms / 1e9, "weight_stream_gbps": params * 2 / ms / 1e6, "launches": n_launch,
                     "cuda_eager_ms": ms_ref}
    print(res["t5_xxl"], flush=True)
    del enc, sd
    torch.cuda.empty_cache()
    # ---- CLIP-L
    ccfg = to.CLIP_L
    csd = to.init_state_dict(to.clip_spec(ccfg), 8, dtype=torch.bfloat16, device="cuda")
    cenc = B200CLIPTextEncoder(dict(ccfg.to_dict(), hidden_act="quick_gelu"), csd.__getitem__, device="cuda:0", cache=False)
    cids = torch.randint(3, ccfg.vocab_size - 2, (1, 77), device="cuda")
    cids[:, 30:] = ccfg.vocab_size - 1
    l0 = cenc.counter("launches")
    cenc(cids)
    n_launch = cenc.counter("launches") - l0
    ms = timeit(lambda: cenc(cids))
    ms_ref = timeit(lambda: to.clip_encode(csd, ccfg, cids), iters=3, warm=1)
    cached = B200CLIPTextEncoder(dict(ccfg.to_dict(), hidden_act="quick_gelu"), csd.__getitem__, device="cuda:0", cache=True)
    cached(cids)
    ms_hit = timeit(lambda: cached(cids))
    res["clip_l"] = {"tokens": 77, "engine_ms": ms, "launches": n_launch, "cuda_eager_ms": ms_ref, "engine_cached_prompt_ms": ms_hit}
    print(res["clip_l"], flush=True)
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
