"""Kernel micro-benchmarks through the C ABI (CUDA events, L2-cold via a flush buffer): GEMM shapes of the FLUX blocks
and joint attention, in TFLOP/s.  Usage: python tools/bench_kernels.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textflux_b200 import _lib  # noqa: E402


def timeit(fn, iters=10, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(1024 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > L2, and long enough on the GPU to hide the host-side launch cost of the timed call
    res = []
    shapes = [(2560, 9216, 3072, "qkv (double, per stream-pair)"), (2560, 3072, 3072, "out-proj"),
              (2560, 12288, 3072, "ff up"), (2560, 3072, 12288, "ff down"), (2560, 21504, 3072, "single qkv+mlp"),
              (2560, 3072, 15360, "single proj_out"), (5120, 21504, 3072, "single qkv+mlp cfg3"),
              (8192, 8192, 8192, "square 8192")]
    if args.quick:
        shapes = shapes[:4]
    if args.only and args.only != "gemm":
        shapes = []
    for M, N, K, name in shapes:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
        b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        row = {"kernel": "gemm", "name": name, "M": M, "N": N, "K": K}
        for cg in (1, 2, 22, 24):
            def f():
                _lib.check(lib.tfx_op_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, 0, None, None, cg, st))
            ms = timeit(f, flush=flush)
            row[f"cg{cg}_ms"] = ms
            row[f"cg{cg}_tflops"] = 2.0 * M * N * K / ms / 1e9
        gate = torch.randn(N, device="cuda").to(torch.bfloat16)
        for bn in (256, 224, 192):
            os.environ["TFX_OP_LINEAR_BLOCK_N"] = str(bn)
            def f():
                _lib.check(lib.tfx_op_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, 0, None, None, 2, st))
            ms = timeit(f, flush=flush)
            row[f"cg2_bn{bn}_tflops"] = 2.0 * M * N * K / ms / 1e9
        os.environ.pop("TFX_OP_LINEAR_BLOCK_N")
        for mode, mname in ((1, "gelu"), (2, "gate_res")):
            def f():
                _lib.check(lib.tfx_op_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, mode,
                                             gate.data_ptr(), out.data_ptr(), 2, st))
            ms = timeit(f, flush=flush)
            row[f"cg2_{mname}_tflops"] = 2.0 * M * N * K / ms / 1e9
        ms = timeit(lambda: torch.nn.functional.linear(A, W, b), flush=flush)
        row["cublas_ms"] = ms
        row["cublas_tflops"] = 2.0 * M * N * K / ms / 1e9
        print(row, flush=True)
        res.append(row)
        del A, W, out
    # tile order (GemmParams::m_band) on the wide-K GEMMs, and the cost of the unfused-LoRA side path
    band_shapes = [(5120, 3072, 15360, "single proj_out cfg3"), (5120, 3072, 12288, "ff down cfg3"), (8704, 3072, 15360, "single proj_out cfg5"),
                   (2560, 3072, 15360, "single proj_out cfg2"), (5120, 9216, 3072, "qkv cfg3"), (5120, 12288, 3072, "ff up cfg3")]
    for M, N, K, name in (band_shapes if args.only in ("", "band") else []):
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
        b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        la = (torch.randn(64, K, device="cuda") * 0.02).to(torch.bfloat16)
        lb = (torch.randn(N, 64, device="cuda") * 0.02).to(torch.bfloat16)
        tt = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
        row = {"kernel": "gemm_band", "name": name, "M": M, "N": N, "K": K}
        for band in (0, 2, 3, 4, 5, 6, 8, 10):
            def f():
                _lib.check(lib.tfx_op_linear_lora(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), None, None, None, out.data_ptr(), N, M, N, K, 0,
                                                  None, None, 2, band, st))
            ms = timeit(f, flush=flush)
            row[f"band{band}_tflops"] = 2.0 * M * N * K / ms / 1e9
        def f():
            _lib.check(lib.tfx_op_linear_lora(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), la.data_ptr(), lb.data_ptr(), tt.data_ptr(),
                                              out.data_ptr(), N, M, N, K, 0, None, None, 2, 0, st))
        ms = timeit(f, flush=flush)
        row["side_lora_tflops"] = 2.0 * M * N * K / ms / 1e9
        row["side_lora_ms"] = ms
        print(row, flush=True)
        res.append(row)
        del A, W, out
    for (T, S) in ([] if args.only not in ("", "attention") else [(512, 2048), (512, 4096), (512, 4608), (512, 8192), (512, 12288)]):
        H, dh, N = 24, 128, T + S
        q = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
        k = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
        v = torch.randn(1, H, N, dh, device="cuda").to(torch.bfloat16)
        out = torch.empty(N, H * dh, device="cuda", dtype=torch.bfloat16)
        row = {"kernel": "attention", "N": N, "H": H, "dh": dh}
        fl = 4.0 * N * N * H * dh
        for qt in [26, 29, 27, 7, 37] + [int(x) for x in os.environ.get("TFX_EXTRA_QT", "").split(",") if x]:
            def f():
                _lib.check(lib.tfx_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), H * dh, 1, H, T, S, dh, qt, st))
            ms = timeit(f, flush=flush)
            row[f"qt{qt}_ms"] = ms
            row[f"qt{qt}_tflops"] = fl / ms / 1e9
        ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), flush=flush)
        row["sdpa_ms"] = ms
        row["sdpa_tflops"] = fl / ms / 1e9
        print(row, flush=True)
        res.append(row)
    for rows, D in ([] if args.only not in ("", "ln") else [(2560, 3072), (5120, 3072)]):
        x = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
        y = torch.empty_like(x)
        mod = torch.randn(1, 3 * D, device="cuda").to(torch.bfloat16)
        def f():
            _lib.check(lib.tfx_op_ln_modulate(x.data_ptr(), y.data_ptr(), rows, D, rows, mod.data_ptr(), 3 * D, 0, D, st))
        ms = timeit(f, flush=flush)
        row = {"kernel": "ln_modulate", "rows": rows, "D": D, "us": ms * 1e3, "GBps": 2 * rows * D * 2 / ms / 1e6}
        print(row, flush=True)
        res.append(row)
    if args.only in ("", "gemv"):
        Nn, K = 344 * 3072, 3072
        W = (torch.randn(Nn, K, device="cuda") * 0.02).to(torch.bfloat16)
        b = torch.zeros(Nn, device="cuda", dtype=torch.bfloat16)
        xx = torch.randn(1, K, device="cuda").to(torch.bfloat16)
        oo = torch.empty(1, Nn, device="cuda", dtype=torch.bfloat16)
        def f():
            _lib.check(lib.tfx_op_gemv(xx.data_ptr(), 1, K, W.data_ptr(), b.data_ptr(), Nn, oo.data_ptr(), 1, st))
        ms = timeit(f)
        row = {"kernel": "gemv_mod", "N": Nn, "K": K, "us": ms * 1e3, "GBps": Nn * K * 2 / ms / 1e6}
        print(row, flush=True)
        res.append(row)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
