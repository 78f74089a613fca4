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

// ---- one KV tile of one query row: the softmax step shared by the three schedules -------------------------------------------------
// State per row: m = reference maximum the probabilities are expressed against (lazily updated: it only moves when the true maximum
// ran away from it by more than 2^8), l = running sum of probabilities against m, pend = true maximum of the last tile seen.
//
// kAttnLag = false (the build default): the tile's row maximum is found first, then the reference is updated, then the exponentials run.
// kAttnLag = true (-DTFX_ATTN_LAG=1, an experiment): the reference is updated from the PREVIOUS tile's maximum (known before the scores arrive), the exponentials start at
//   once and this tile's maximum is found alongside them (FMNMX on the ALU pipe under the MUFU-bound stream) -- the ~330-cycle
//   maximum leaves the chain S-ready -> P-ready that bounds the kernel (profiles/r1j_attn_trace.md).  Probabilities may then exceed 1
//   by the growth of the maximum inside one tile; they are exact in fp32 / bf16 up to 2^kDanger, and a tile whose maximum jumps by
//   more than that (never observed: RMS-normed q, k bound the scores) is redone on the exact path from the scores still in TMEM.
//   The first tile of a row always takes the exact path.
// `handover(half, pk)` stores 32 packed bf16 pairs (64 probabilities) over the S columns and signals the issuer.
constexpr float kAttnRescaleThreshold = 8.0f;  // log2 units
constexpr float kAttnDanger = 30.0f;           // log2 units

template <int kHeadDim>
__device__ __forceinline__ void attn_rescale_o(uint32_t t_o, float alpha) {
#pragma unroll 1
  for (int cch = 0; cch < kHeadDim / 32; ++cch) {
    uint32_t v[32];
    tmem_ld32(t_o + cch * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
    tmem_st32(t_o + cch * 32, v);
  }
  tmem_st_wait();
}

template <int kEmu>
__device__ __forceinline__ void attn_exp_half(const uint32_t (&s0)[32], const uint32_t (&s1)[32], f32x2 c2, f32x2 nmc2, f32x2& sum2, uint32_t (&pk)[32]) {
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const uint32_t(&s)[32] = cc ? s1 : s0;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const f32x2 x2 = fma2(pack2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nmc2);
      float p0, p1;
      if (kEmu > 0 && emu_pair<kEmu>(i >> 1)) {
        ex2_emu2(x2, p0, p1);
      } else {
        float x0, x1;
        unpack2(x2, x0, x1);
        p0 = ex2(x0);
        p1 = ex2(x1);
      }
      sum2 = add2(sum2, pack2(p0, p1));
      pk[cc * 16 + (i >> 1)] = pack_bf16(p0, p1);
    }
  }
}

__device__ __forceinline__ float attn_row_max(const uint32_t (&sr)[4][32]) {
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
    mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
    mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
    mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

__device__ __forceinline__ void attn_load_scores(uint32_t t_s, int valid, uint32_t (&sr)[4][32]) {
  tmem_ld32(t_s + 0, sr[0]);
  tmem_ld32(t_s + 32, sr[1]);
  tmem_ld32(t_s + 64, sr[2]);
  tmem_ld32(t_s + 96, sr[3]);
  tmem_ld_wait();
  if (valid < 128) {  // ragged last tile: keys past N score -inf -> probability 0
#pragma unroll
    for (int cch = 0; cch < 4; ++cch)
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (cch * 32 + i >= valid) sr[cch][i] = 0xff800000u;
  }
}

// `first`: first tile of this row's accumulation (O holds nothing yet).  `after_max()` is a hook for trace stamps.
#ifndef TFX_ATTN_LAG
#define TFX_ATTN_LAG 0  // measured (profiles/r2m_attn_lag.md): correct, 1-8 % SLOWER than the exact-maximum path -> not compiled in
#endif
constexpr bool kAttnLag = TFX_ATTN_LAG != 0;

// kAttnNoMax (-DTFX_ATTN_NOMAX=1): after a row's first tile NO maximum is computed.  The reference m stays where it is and the
//   probabilities are evaluated against it directly; what the maximum protected against -- probabilities running away from the
//   reference -- is detected on the half-row sums that are computed anyway (sum of 64 probabilities > 2^16, or inf / NaN), before the
//   half is handed over.  A flagged first half falls back to the exact path (nothing handed over yet, scores in registers); a flagged
//   second half waits for the PV MMAs of the first half (`wait_h0`, a commit the issuer makes per tile), moves the reference, rescales
//   O and l and re-evaluates its 64 columns.  Saves the 98 FMNMX per warp and tile and takes the maximum off the S-ready -> P path.
#ifndef TFX_ATTN_NOMAX
#define TFX_ATTN_NOMAX 0
#endif
constexpr bool kAttnNoMax = TFX_ATTN_NOMAX != 0;
constexpr float kAttnSumLimit = 65536.0f;

__device__ __forceinline__ float attn_half_max(const uint32_t (&a)[32], const uint32_t (&b)[32]) {
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    mx0 = fmaxf(mx0, __uint_as_float(a[i]));
    mx1 = fmaxf(mx1, __uint_as_float(b[i]));
  }
  return fmaxf(mx0, mx1);
}

template <int kHeadDim, int kEmu, bool kSplit, typename Handover, typename Stamp, typename WaitH0>
__device__ __forceinline__ void attn_softmax_tile(uint32_t t_s, uint32_t t_o, float c, int valid, bool first, float& m, float& l,
                                                  float& pend, Handover&& handover, Stamp&& stamp, WaitH0&& wait_h0) {
  const f32x2 c2 = pack2(c, c);
  if (kAttnNoMax && kSplit && !kAttnLag) {  // whole-P hand-over keeps the exact path: its first half is not consumed early
    uint32_t sr[4][32];
    attn_load_scores(t_s, valid, sr);
    stamp(1);
    uint32_t pk[32];
    if (!first) {
      const float mc = m * c;
      const f32x2 nmc2 = pack2(-mc, -mc);
      f32x2 sum2 = pack2(0.f, 0.f);
      attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
      float a0, a1;
      unpack2(sum2, a0, a1);
      const float s_a = a0 + a1;
      stamp(2);
      if (!__any_sync(0xffffffffu, !(s_a <= kAttnSumLimit))) {
        handover(0, pk);
        stamp(3);
        sum2 = pack2(0.f, 0.f);
        attn_exp_half<kEmu>(sr[2], sr[3], c2, nmc2, sum2, pk);
        unpack2(sum2, a0, a1);
        float s_b = a0 + a1;
        const bool over = !(s_b <= kAttnSumLimit);
        if (__any_sync(0xffffffffu, over)) {
          // second half ran away after the first was handed over: O holds PV(0..j-1) + PV over keys 0..63 of this tile once wait_h0 returns
          const float m_new = over ? fmaxf(attn_half_max(sr[2], sr[3]), m) : m;
          const float alpha = ex2((m - m_new) * c);
          wait_h0();
          attn_rescale_o<kHeadDim>(t_o, alpha);
          l = (l + s_a) * alpha;
          m = m_new;
          const float mc1 = m * c;
          sum2 = pack2(0.f, 0.f);
          attn_exp_half<kEmu>(sr[2], sr[3], c2, pack2(-mc1, -mc1), sum2, pk);
          unpack2(sum2, a0, a1);
          l += a0 + a1;
        } else {
          l += s_a + s_b;
        }
        handover(1, pk);
        stamp(4);
        pend = m;
        return;
      }
      // first half ran away: nothing handed over yet -> exact path on the scores in registers
    }
    const float mx = attn_row_max(sr);
    const bool need = (mx - m) * c > kAttnRescaleThreshold;  // true on the first tile (m = -inf)
    const float m_new = need ? mx : m;
    const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
    if (!first && __any_sync(0xffffffffu, need)) attn_rescale_o<kHeadDim>(t_o, alpha);
    stamp(2);
    const float mc = m_new * c;
    const f32x2 nmc2 = pack2(-mc, -mc);
    f32x2 sum2 = pack2(0.f, 0.f);
    attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
    handover(0, pk);
    stamp(3);
    attn_exp_half<kEmu>(sr[2], sr[3], c2, nmc2, sum2, pk);
    handover(1, pk);
    stamp(4);
    float sum0, sum1;
    unpack2(sum2, sum0, sum1);
    l = l * alpha + (sum0 + sum1);
    m = m_new;
    pend = mx;
    return;
  }
  if (kAttnLag && !first) {
    // ---- fast path: reference from the previous tile's maximum; this tile's maximum rides along with the exponentials
    const bool need = (pend - m) * c > kAttnRescaleThreshold;
    const float m_new = need ? pend : m;
    if (__any_sync(0xffffffffu, need)) {
      const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
      attn_rescale_o<kHeadDim>(t_o, alpha);  // O holds PV(0..j-1): retired, s_full(j) flipped behind PV(j-1)
      l *= alpha;
      m = m_new;
    }
    uint32_t sr[4][32];
    attn_load_scores(t_s, valid, sr);
    stamp(1);
    const float mc = m * c;
    const f32x2 nmc2 = pack2(-mc, -mc);
    f32x2 sum2 = pack2(0.f, 0.f);
    uint32_t pk[32];
    attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
    const float mx = attn_row_max(sr);
    stamp(2);
    if (!__any_sync(0xffffffffu, (mx - m) * c > kAttnDanger)) {
      handover(0, pk);
      stamp(3);
      attn_exp_half<kEmu>(sr[2], sr[3], c2, nmc2, sum2, pk);
      handover(1, pk);
      stamp(4);
      float sum0, sum1;
      unpack2(sum2, sum0, sum1);
      l += sum0 + sum1;
      pend = mx;
      return;
    }
    // a jump of more than 2^kAttnDanger inside one tile: nothing was handed over yet, S is intact in TMEM -> exact path below
  }
  uint32_t sr[4][32];
  attn_load_scores(t_s, valid, sr);
  stamp(1);
  const float mx = attn_row_max(sr);
  // lazy rescaling: the reference point m only moves when the true max ran away from it
  const bool need = (mx - m) * c > kAttnRescaleThreshold;  // true on the first tile (m = -inf)
  const float m_new = need ? mx : m;
  const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
  if (!first && __any_sync(0xffffffffu, need)) attn_rescale_o<kHeadDim>(t_o, alpha);
  stamp(2);
  const float mc = m_new * c;
  const f32x2 nmc2 = pack2(-mc, -mc);
  f32x2 sum2 = pack2(0.f, 0.f);
  uint32_t pk[32];
  attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
  handover(0, pk);
  stamp(3);
  attn_exp_half<kEmu>(sr[2], sr[3], c2, nmc2, sum2, pk);
  handover(1, pk);
  stamp(4);
  float sum0, sum1;
  unpack2(sum2, sum0, sum1);
  l = l * alpha + (sum0 + sum1);
  m = m_new;
  pend = mx;
}

}  // namespace tfx
