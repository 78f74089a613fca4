// Joint [text;image] non-causal flash attention on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces F.scaled_dot_product_attention(q, k, v) of FluxAttnProcessor2_0 (attention_processor.py:2039-2043,
// SURVEY.md §2d K2): q,k,v [B,H,N,dh] bf16 (already RMS-normed and RoPE'd by the QKV GEMM epilogue), scale dh^-0.5,
// no mask, output written token-major [rows, H*dh] (the transpose+reshape of :2042 is folded into the store, and for
// single-stream blocks the store lands directly in the [attn | mlp] concat buffer, transformer_flux.py:732).
//
// One CTA handles kQTiles x 128 query rows of one (batch, head).  Per 128-row KV tile j and query tile q:
//     S_q  = Q_q K_j^T      tcgen05.mma SS  (A = Q smem K-major, B = K smem K-major)  -> TMEM fp32 [128 x 128]
//     P_q  = exp2(S_q*c - m*c)   softmax warpgroup q: one thread per row, running max/sum in registers,
//                                P written back to TMEM as packed bf16 over the S columns it has consumed
//     O_q += P_q V_j        tcgen05.mma TS  (A = P from TMEM, B = V smem MN-major)    -> TMEM fp32 [128 x dh]
// The single MMA-issuing thread interleaves the two query tiles so that softmax of one overlaps the MMAs of the other.
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
};

constexpr int kAttnKvStages = 2;

template <int kHeadDim, int kQTiles>
struct AttnCfg {
  static constexpr int kHalves = kHeadDim / 64;             // 64-element (128 B) swizzle atoms along dh
  static constexpr int kTileBytes = 128 * kHeadDim * 2;     // one 128-row Q/K/V tile
  static constexpr int kThreads = 64 + 128 * kQTiles;       // TMA warp + MMA warp + softmax warpgroups
  static constexpr int kSmemBytes = (kQTiles + 2 * kAttnKvStages) * kTileBytes + 1024 + 256;
  static constexpr int kTmemCols = 512;
  static constexpr int kSCol = 0;                           // S_q at kSCol + q*128 (P aliases its first 64 columns)
  static constexpr int kOCol = 256;                         // O_q at kOCol + q*128
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

template <int kHeadDim, int kQTiles, int kEmu>
__global__ void __launch_bounds__(AttnCfg<kHeadDim, kQTiles>::kThreads, 1)  // 10 warps = 3 on one SMSP -> 168 regs/thread
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<kHeadDim, kQTiles>;
  constexpr int kHalves = Cfg::kHalves;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                      // [kQTiles][kHalves][128][64]
  uint8_t* sK = sQ + kQTiles * Cfg::kTileBytes;            // [stages][kHalves][128][64]
  uint8_t* sV = sK + kAttnKvStages * Cfg::kTileBytes;      // [stages][kHalves][128 kv][64 dh]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttnKvStages * Cfg::kTileBytes);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = q_full + 1;                // [stages]
  uint64_t* v_full = k_full + kAttnKvStages;    // [stages]
  uint64_t* kv_empty = v_full + kAttnKvStages;  // [stages]
  uint64_t* s_full = kv_empty + kAttnKvStages;  // [kQTiles]
  uint64_t* p_full = s_full + kQTiles;          // [kQTiles]
  uint64_t* pv_done = p_full + kQTiles;         // [kQTiles]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + kQTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (128 * kQTiles);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + head;
  const int n_kv = (p.N + 127) / 128;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAttnKvStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < kQTiles; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_base_ptr, Cfg::kTmemCols);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_arrive_expect_tx(q_full, kQTiles * Cfg::kTileBytes);
    for (int q = 0; q < kQTiles; ++q)
      for (int h = 0; h < kHalves; ++h)
        tma_load_3d(&tmQ, q_full, sQ + q * Cfg::kTileBytes + h * kHalfBytes, h * 64, q0 + q * 128, bh, kEvictFirst);
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&k_full[stage], Cfg::kTileBytes);
      for (int h = 0; h < kHalves; ++h)
        tma_load_3d(&tmK, &k_full[stage], sK + stage * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
      mbar_arrive_expect_tx(&v_full[stage], Cfg::kTileBytes);
      for (int h = 0; h < kHalves; ++h)
        tma_load_3d(&tmV, &v_full[stage], sV + stage * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
      if (++stage == kAttnKvStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (dh contiguous)
    auto issue_qk = [&](int q, int stage) {
      const uint32_t a0 = smem_u32(sQ + q * Cfg::kTileBytes);
      const uint32_t b0 = smem_u32(sK + stage * Cfg::kTileBytes);
#pragma unroll
      for (int kk = 0; kk < kHeadDim / 16; ++kk) {
        const uint32_t off = uint32_t((kk / 4) * kHalfBytes + (kk % 4) * 32);
        umma_ss<1>(tmem_base + uint32_t(Cfg::kSCol + q * 128), make_smem_desc(a0 + off, 16, 1024, kLayoutSW128),
                   make_smem_desc(b0 + off, 16, 1024, kLayoutSW128), idesc_qk, kk != 0);
      }
      umma_commit(&s_full[q]);
    };
    auto issue_pv = [&](int q, int stage, int j) {
      const uint32_t v0 = smem_u32(sV + stage * Cfg::kTileBytes);
#pragma unroll
      for (int kk = 0; kk < 128 / 16; ++kk) {
        // 16 kv rows (= 2 KiB) per step; dh halves are kHalfBytes apart (leading-dimension byte offset)
        umma_ts(tmem_base + uint32_t(Cfg::kOCol + q * 128), tmem_base + uint32_t(Cfg::kSCol + q * 128 + kk * 8),
                make_smem_desc(v0 + kk * 2048, kHalfBytes, 1024, kLayoutSW128), idesc_pv, (j | kk) != 0);
      }
      umma_commit(&pv_done[q]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    for (int q = 0; q < kQTiles; ++q) issue_qk(q, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_kv; ++j) {
      const int nstage = (stage + 1 == kAttnKvStages) ? 0 : stage + 1;
      const uint32_t nphase = (stage + 1 == kAttnKvStages) ? phase ^ 1 : phase;
      mbar_wait(&v_full[stage], phase);
      for (int q = 0; q < kQTiles; ++q) {
        mbar_wait(&p_full[q], j & 1);
        tc_fence_after();
        issue_pv(q, stage, j);
        if (q == kQTiles - 1) umma_commit(&kv_empty[stage]);
        if (j + 1 < n_kv) {
          if (q == 0) { mbar_wait(&k_full[nstage], nphase); tc_fence_after(); }
          issue_qk(q, nstage);
        }
      }
      stage = nstage;
      phase = nphase;
    }
  } else if (warp >= 2) {
    // ===================== softmax warpgroups: one thread per query row =====================
    const int q = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int pos = q0 + q * 128 + row_in_tile;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128);
    const float c = p.scale_log2;
    float m = -INFINITY, l = 0.f;
    const float kRescaleThreshold = 8.0f;  // log2 units: keep a stale row max until it is off by more than 2^8
    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.N - j * 128;  // >= 128 on every tile but possibly the last
      mbar_wait(&s_full[q], j & 1);
      tc_fence_after();
      // the whole 128-column score row of this thread in registers: one TMEM pass
      uint32_t sr[4][32];
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
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
        mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
        mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // lazy rescaling: the reference point m only moves when the true max ran away from it
      const bool need = (mx - m) * c > kRescaleThreshold;  // true on the first tile (m = -inf)
      const float m_new = need ? mx : m;
      const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
      const float mc = m_new * c;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[2][32];
      if (kEmu > 0 && valid >= 128) {
        const f32x2 c2 = pack2(c, c), nmc2 = pack2(-mc, -mc);
        f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
        for (int cch = 0; cch < 4; ++cch) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const f32x2 x2 = fma2(pack2(__uint_as_float(sr[cch][i]), __uint_as_float(sr[cch][i + 1])), c2, nmc2);
            float p0, p1;
            if (emu_pair<kEmu>(i >> 1)) {
              ex2_emu2(x2, p0, p1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            sum2 = add2(sum2, pack2(p0, p1));
            pk[cch >> 1][(cch & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
          }
        }
        unpack2(sum2, sum0, sum1);
      } else {
#pragma unroll
        for (int cch = 0; cch < 4; ++cch) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(sr[cch][i]), c, -mc));
            const float p1 = ex2(fmaf(__uint_as_float(sr[cch][i + 1]), c, -mc));
            sum0 += p0;
            sum1 += p1;
            pk[cch >> 1][(cch & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
          }
        }
      }
      tmem_st32(t_s + 0, pk[0]);  // P (bf16 pairs) over the S columns just consumed
      tmem_st32(t_s + 32, pk[1]);
      l = l * alpha + (sum0 + sum1);
      m = m_new;
      tmem_st_wait();
      if (j > 0) {
        mbar_wait(&pv_done[q], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[q]);
    }
    // ---- finalize: O / l -> bf16, token-major store
    mbar_wait(&pv_done[q], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim;
#pragma unroll 1
    for (int cch = 0; cch < kHeadDim / 32; ++cch) {
      uint32_t v[32];
      tmem_ld32(t_o + cch * 32, v);
      tmem_ld_wait();
      if (row_ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + cch * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[8 * i + 0]) * inv_l, __uint_as_float(v[8 * i + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(v[8 * i + 2]) * inv_l, __uint_as_float(v[8 * i + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(v[8 * i + 4]) * inv_l, __uint_as_float(v[8 * i + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(v[8 * i + 6]) * inv_l, __uint_as_float(v[8 * i + 7]) * inv_l);
          d4[i] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// Version 2 of the schedule ("QK-ahead"), templated on query tiles per CTA (kQT) and keys per tile (kKV).  In the kernel above a query tile's chain is strictly
//   softmax(j) -> PV(j), QK(j+1) -> softmax(j+1): softmax warps and the tensor pipe wait on each other (ncu: 31 % of
// warp samples sit in the wait for S, tensor pipe 46 % busy, profiles/r1d).  Here KV tiles are 64 keys, so the
// 512 TMEM columns hold TWO score buffers per query tile (2 q-tiles x 2 x 64) next to the two O accumulators
// (2 x 128), and the issuer runs QK two tiles ahead: S(j+1) is already in TMEM when softmax(j) finishes, the softmax
// warpgroups run back to back and the MMAs (PV(j), QK(j+2)) execute in their shadow.
template <int kHeadDim, int kQT, int kKV>
struct Attn2Cfg {
  static constexpr int kHalves = kHeadDim / 64;
  static constexpr int kQBytes = 128 * kHeadDim * 2;   // one 128-row query tile
  static constexpr int kKvRows = kKV;
  static constexpr int kKvBytes = kKvRows * kHeadDim * 2;  // one K (or V) tile
  // K runs two tiles ahead of V (QK(j+2) is issued next to PV(j)), so the rings are separate: 4 K stages, and as many
  // V stages as the 227 KB allow
  static constexpr int kKStages = 4;
  static constexpr int kVStages = (kQT * kQBytes + 8 * kKvBytes <= 200 * 1024) ? 4 : 2;
  static constexpr int kThreads = 64 + 128 * kQT;
  static constexpr int kSmemBytes = kQT * kQBytes + (kKStages + kVStages) * kKvBytes + 1024 + 256;
  static constexpr int kOCol = kQT * 2 * kKV;  // S[q][b] at (q*2+b)*kKV, O_q at kOCol + q*128
  static_assert(kQT * 2 * kKV + kQT * 128 <= 512, "TMEM budget");
};

template <int kHeadDim, int kQT, int kKV>
__global__ void __launch_bounds__(Attn2Cfg<kHeadDim, kQT, kKV>::kThreads, 1)
attention2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn2Cfg<kHeadDim, kQT, kKV>;
  constexpr int kHalves = Cfg::kHalves;
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages;
  constexpr int kQHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKvHalfBytes = kKV * 128;  // kKV rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                           // [kQT][kHalves][128][64]
  uint8_t* sK = sQ + kQT * Cfg::kQBytes;        // [stages][kHalves][kKV][64]
  uint8_t* sV = sK + kKS * Cfg::kKvBytes;       // [stages][kHalves][kKV kv][64 dh]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVS * Cfg::kKvBytes);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = q_full + 1;           // [kKS]
  uint64_t* k_empty = k_full + kKS;        // [kKS]
  uint64_t* v_full = k_empty + kKS;        // [kVS]
  uint64_t* v_empty = v_full + kVS;        // [kVS]
  uint64_t* s_full = v_empty + kVS;        // [2 q][2 buffers]
  uint64_t* p_full = s_full + 4;           // [2]
  uint64_t* pv_done = p_full + 2;          // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (128 * kQT);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + head;
  const int n_kv = (p.N + kKV - 1) / kKV;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < kVS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_base_ptr, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_arrive_expect_tx(q_full, kQT * Cfg::kQBytes);
    for (int q = 0; q < kQT; ++q)
      for (int h = 0; h < kHalves; ++h)
        tma_load_3d(&tmQ, q_full, sQ + q * Cfg::kQBytes + h * kQHalfBytes, h * 64, q0 + q * 128, bh, kEvictFirst);
    // K tiles run two ahead of V tiles, like their consumers
    for (int t = 0; t < n_kv + 2; ++t) {
      if (t < n_kv) {
        const int st = t % kKS;
        mbar_wait(&k_empty[st], ((t / kKS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::kKvBytes);
        for (int h = 0; h < kHalves; ++h)
          tma_load_3d(&tmK, &k_full[st], sK + st * Cfg::kKvBytes + h * kKvHalfBytes, h * 64, t * kKV, bh, kEvictLast);
      }
      if (t >= 2) {
        const int j = t - 2, st = j % kVS;
        mbar_wait(&v_empty[st], ((j / kVS) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Cfg::kKvBytes);
        for (int h = 0; h < kHalves; ++h)
          tma_load_3d(&tmV, &v_full[st], sV + st * Cfg::kKvBytes + h * kKvHalfBytes, h * 64, j * kKV, bh, kEvictLast);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, kKV, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, kHeadDim, 0, 1);
    auto issue_qk = [&](int q, int j) {
      const int stage = j % kKS;
      const uint32_t a0 = smem_u32(sQ + q * Cfg::kQBytes);
      const uint32_t b0 = smem_u32(sK + stage * Cfg::kKvBytes);
      const uint32_t d = tmem_base + uint32_t((q * 2 + (j & 1)) * kKV);
#pragma unroll
      for (int kk = 0; kk < kHeadDim / 16; ++kk) {
        umma_ss<1>(d, make_smem_desc(a0 + uint32_t((kk / 4) * kQHalfBytes + (kk % 4) * 32), 16, 1024, kLayoutSW128),
                   make_smem_desc(b0 + uint32_t((kk / 4) * kKvHalfBytes + (kk % 4) * 32), 16, 1024, kLayoutSW128), idesc_qk, kk != 0);
      }
      umma_commit(&s_full[q * 2 + (j & 1)]);
    };
    auto issue_pv = [&](int q, int j) {
      const int stage = j % kVS;
      const uint32_t v0 = smem_u32(sV + stage * Cfg::kKvBytes);
      const uint32_t pcol = tmem_base + uint32_t((q * 2 + (j & 1)) * kKV);
#pragma unroll
      for (int kk = 0; kk < kKV / 16; ++kk) {
        umma_ts(tmem_base + uint32_t(Cfg::kOCol + q * 128), pcol + uint32_t(kk * 8),
                make_smem_desc(v0 + kk * 2048, kKvHalfBytes, 1024, kLayoutSW128), idesc_pv, (j | kk) != 0);
      }
      umma_commit(&pv_done[q]);
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < 2 && j < n_kv; ++j) {  // run two KV tiles ahead
      mbar_wait(&k_full[j % kKS], 0);
      tc_fence_after();
      for (int q = 0; q < kQT; ++q) issue_qk(q, j);
      umma_commit(&k_empty[j % kKS]);
    }
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&v_full[j % kVS], (j / kVS) & 1);
      for (int q = 0; q < kQT; ++q) {
        mbar_wait(&p_full[q], j & 1);
        tc_fence_after();
        issue_pv(q, j);
        if (j + 2 < n_kv) {
          if (q == 0) { mbar_wait(&k_full[(j + 2) % kKS], ((j + 2) / kKS) & 1); tc_fence_after(); }
          issue_qk(q, j + 2);  // overwrites S[q][j&1] = P(q,j): ordered behind PV(q,j) by the in-order tensor pipe
          if (q == kQT - 1) umma_commit(&k_empty[(j + 2) % kKS]);
        }
      }
      umma_commit(&v_empty[j % kVS]);
    }
  } else if (warp >= 2) {
    // ===================== softmax warpgroups: one thread per query row =====================
    const int q = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int pos = q0 + q * 128 + quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128);
    const float c = p.scale_log2;
    const float kRescaleThreshold = 8.0f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.N - j * kKV;
      const uint32_t t_s = t_lane + uint32_t((q * 2 + (j & 1)) * kKV);
      mbar_wait(&s_full[q * 2 + (j & 1)], (j >> 1) & 1);
      tc_fence_after();
      constexpr int kCh = kKV / 32;
      uint32_t sr[kCh][32];
#pragma unroll
      for (int cch = 0; cch < kCh; ++cch) tmem_ld32(t_s + cch * 32, sr[cch]);
      tmem_ld_wait();
      if (valid < kKV) {
#pragma unroll
        for (int cch = 0; cch < kCh; ++cch)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cch * 32 + i >= valid) sr[cch][i] = 0xff800000u;
      }
      float mxa[kCh];
#pragma unroll
      for (int cch = 0; cch < kCh; ++cch) mxa[cch] = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int cch = 0; cch < kCh; ++cch) mxa[cch] = fmaxf(mxa[cch], __uint_as_float(sr[cch][i]));
      float mx = mxa[0];
#pragma unroll
      for (int cch = 1; cch < kCh; ++cch) mx = fmaxf(mx, mxa[cch]);
      const bool need = (mx - m) * c > kRescaleThreshold;
      const float m_new = need ? mx : m;
      const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
      const float mc = m_new * c;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[kCh / 2][32];
#pragma unroll
      for (int cch = 0; cch < kCh; ++cch) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(sr[cch][i]), c, -mc));
          const float p1 = ex2(fmaf(__uint_as_float(sr[cch][i + 1]), c, -mc));
          sum0 += p0;
          sum1 += p1;
          pk[cch >> 1][(cch & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
        }
      }
#pragma unroll
      for (int h2 = 0; h2 < kCh / 2; ++h2) tmem_st32(t_s + h2 * 32, pk[h2]);
      l = l * alpha + (sum0 + sum1);
      m = m_new;
      tmem_st_wait();
      if (j > 0) {
        // Wait for PV(q, j-1) every iteration: (1) O may only be rescaled once it has landed (PV(q, j) cannot start
        // before p_full below); (2) it keeps this warpgroup at most one phase ahead of the issuer on p_full -- with
        // S double-buffered the softmax could otherwise complete two phases before the issuer looks, and an mbarrier
        // parity wait cannot tell phase k from phase k+2.
        mbar_wait(&pv_done[q], (j - 1) & 1);
        tc_fence_after();
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[q]);
    }
    mbar_wait(&pv_done[q], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim;
#pragma unroll 1
    for (int cch = 0; cch < kHeadDim / 32; ++cch) {
      uint32_t v[32];
      tmem_ld32(t_o + cch * 32, v);
      tmem_ld_wait();
      if (row_ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + cch * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[8 * i + 0]) * inv_l, __uint_as_float(v[8 * i + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(v[8 * i + 2]) * inv_l, __uint_as_float(v[8 * i + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(v[8 * i + 4]) * inv_l, __uint_as_float(v[8 * i + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(v[8 * i + 6]) * inv_l, __uint_as_float(v[8 * i + 7]) * inv_l);
          d4[i] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
