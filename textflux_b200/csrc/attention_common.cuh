// Shared pieces of the joint [text;image] flash-attention kernels (attention3.cuh, attention4.cuh) for sm_100a.
//
// They replace F.scaled_dot_product_attention(q, k, v) of FluxAttnProcessor2_0 (attention_processor.py:2039-2043,
// SURVEY.md section 2d K2): q,k,v [B,H,N,dh] bf16 (already RMS-normed and RoPE'd by the QKV GEMM epilogue), scale dh^-0.5,
// no mask, output written token-major [rows, H*dh] (the transpose+reshape of :2042 is folded into the store, and for
// single-stream blocks the store lands directly in the [attn | mlp] concat buffer, transformer_flux.py:732).
//
// Per 128-row KV tile j and 128-row query tile q:
//     S_q  = Q_q K_j^T      tcgen05.mma SS  (A = Q smem K-major, B = K smem K-major)  -> TMEM fp32 [128 x 128]
//     P_q  = exp2(S_q*c - m*c)   softmax warpgroup q: one thread per row, running max/sum in registers,
//                                P written back to TMEM as packed bf16 over the S columns it has consumed
//     O_q += P_q V_j        tcgen05.mma TS  (A = P from TMEM, B = V smem MN-major)    -> TMEM fp32 [128 x dh]
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace tfx {

struct AttnParams {
  int B, H, N, T, S;  // N = T + S joint tokens per sample
  float scale_log2;   // dh^-0.5 * log2(e)
  __nv_bfloat16* out;
  long long ld_out;   // row stride of `out` in elements
  long long* trace = nullptr;  // clock64 stamps of CTA (0,0,0) (attention3 trace builds only; see tools/attn_trace.py)
  long long* cta_trace = nullptr;  // [num CTAs][8] globaltimer ns per CTA (trace builds; tools/attn_timeline.py): 0 entry,
                                   // 1 set-up done, 2 Q + first K landed, 3 first scores seen, 4 last P handed over,
                                   // 5 last PV retired, 6 output stored, 7 SM id
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A fraction of the softmax exponentials is computed on the FMA pipes (Cody-Waite split + degree-3 minimax
// polynomial, max relative error 7.6e-5, far below the bf16 rounding of P) to take load off the 16/clk/SM MUFU.EX2
// unit.  A scalar version of this cost more issue slots than it freed XU cycles (measured -9 %); the packed fp32x2
// version below gains 4-8 % at 2 pairs in 8 (profiles/r1i_kernels.json).
// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): two softmax columns per instruction
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for a pair on the packed pipes; x <= ~8 here (scores minus the lazily updated row maximum)
__device__ __forceinline__ void ex2_emu2(f32x2 x, float& p0, float& p1) {
  float x0, x1;
  unpack2(x, x0, x1);
  x = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const f32x2 kMagic = pack2(12582912.0f, 12582912.0f), kNegMagic = pack2(-12582912.0f, -12582912.0f);
  const f32x2 xr = add2(x, kMagic);
  const f32x2 n = add2(xr, kNegMagic);
  const f32x2 f = fma2(n, pack2(-1.0f, -1.0f), x);
  f32x2 pf = fma2(f, pack2(0.05520550534f, 0.05520550534f), pack2(0.24261397123f, 0.24261397123f));
  pf = fma2(pf, f, pack2(0.69325476885f, 0.69325476885f));
  pf = fma2(pf, f, pack2(0.99992769957f, 0.99992769957f));
  float r0, r1, q0, q1;
  unpack2(xr, r0, r1);
  unpack2(pf, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}
// kEmu of every 8 column PAIRS go through ex2_emu2 (0 = all on MUFU)
template <int kEmu>
__device__ __forceinline__ bool emu_pair(int pair) {
  const int r = pair & 7;
  return (kEmu >= 1 && r == 6) || (kEmu >= 2 && r == 2) || (kEmu >= 3 && r == 4) || (kEmu >= 4 && r == 0);
}

}  // namespace tfx
