// Joint [text;image] flash attention, schedule 8 (experiment): schedule 3's CTA (one per (head, 256 query rows), two 128-row query
// tiles, S / P / O in TMEM, 12 warps) with the two softmax warpgroups TIME-MULTIPLEXED over the two query tiles instead of owning one
// each: a row is handled by a pair of threads -- warpgroup 0 takes score columns 0..63 of whichever tile is ready, warpgroup 1
// columns 64..127 -- so both warps of a scheduler work on the SAME tile and its S-ready -> P-ready latency is the XU time of the
// tile (~780 cycles per scheduler) instead of one warp's ~1330 cycles of dependent MUFU / FFMA2 chains; the tiles then take turns.
// P is handed over in two steps of (32 + 32) keys: the first 32 columns of BOTH halves, then the rest (PV MMAs kk = 0,1,4,5, then
// 2,3,6,7), so only 256 cycles of PV follow the last exponential.  Pair exchange of half-row maxima as in schedule 6.
#pragma once
#include <cuda.h>

#include "../../textflux_b200/csrc/attention3.cuh"

namespace tfx {

template <int kHeadDim>
struct Attn8Cfg {
  static constexpr int kThreads = 384;  // wg0: TMA, MMA, TMEM alloc, spare; wg1: score columns 0..63 of both tiles; wg2: columns 64..127
  static constexpr int kXchBytes = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] fp32: half-row maxima (and, at the end, sums)
  static constexpr int kSmemBytes = Attn3Cfg<kHeadDim>::kSmemBytes - Attn3Cfg<kHeadDim>::kXchBytes + kXchBytes;
  static constexpr int kRegsSmall = Attn3Cfg<kHeadDim>::kRegsSmall, kRegsLarge = Attn3Cfg<kHeadDim>::kRegsLarge;
};

// 64-thread named barrier of the two warps that share the rows of a lane quadrant: ids 1..4
__device__ __forceinline__ void pair8_bar_sync(int quad) { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); }

// 32 score columns -> 16 packed bf16 pairs (exponentials split between MUFU and the FMA pipe as in attn_exp_half)
template <int kEmu>
__device__ __forceinline__ void attn_exp_quarter(const uint32_t (&s)[32], f32x2 c2, f32x2 nmc2, f32x2& sum2, uint32_t (&pk)[16]) {
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
    pk[i >> 1] = pack_bf16(p0, p1);
  }
}

template <int kHeadDim, int kEmu>
__global__ void __launch_bounds__(Attn8Cfg<kHeadDim>::kThreads, 1)
attention8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn3Cfg<kHeadDim>;
  using Cfg8 = Attn8Cfg<kHeadDim>;
  constexpr int kHalves = Cfg::kHalves;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2][kHalves][128][64]
  uint8_t* sK = sQ + 2 * Cfg::kTileBytes;          // [kKS][kHalves][128][64]
  uint8_t* sV = sK + kKS * Cfg::kTileBytes;        // [kVS][kHalves][128 kv][64 dh]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVS * Cfg::kTileBytes);
  uint64_t* q_full = bars;               // [1]
  uint64_t* k_full = q_full + 1;         // [kKS]
  uint64_t* k_empty = k_full + kKS;      // [kKS]
  uint64_t* v_full = k_empty + kKS;      // [kVS]
  uint64_t* v_empty = v_full + kVS;      // [kVS]
  uint64_t* s_full = v_empty + kVS;      // [2]
  uint64_t* p_full = s_full + 2;         // [2 q][2 halves]
  uint64_t* pv_done = p_full + 4;        // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 parity][2 q][2 half][128 rows]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
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
    for (int i = 0; i < kKS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kVS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 8);      // first 32 columns of both halves: the eight softmax warps
      mbar_init(&p_full[2 * i + 1], 8);  // ... the other 32 columns of both halves
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_base_ptr, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  pdl_wait();

  if (warp < 4) {
    setmaxnreg_dec<Cfg8::kRegsSmall>();
    if (warp == 0) {
      // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
      const bool leader = elect_one();
      if (leader) {
        mbar_arrive_expect_tx(q_full, 2 * Cfg::kTileBytes);
        for (int q = 0; q < 2; ++q)
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmQ, q_full, sQ + q * Cfg::kTileBytes + h * kHalfBytes, h * 64, q0 + q * 128, bh, kEvictFirst);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % kKS, vs = j % kVS;
        mbar_wait(&k_empty[ks], ((j / kKS) & 1) ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&k_full[ks], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmK, &k_full[ks], sK + ks * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        mbar_wait(&v_empty[vs], ((j / kVS) & 1) ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&v_full[vs], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmV, &v_full[vs], sV + vs * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===================== MMA issuer: warp-uniform control flow, one elected lane issues (as schedule 3, split P) ============
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (dh contiguous)
      const bool leader = elect_one();
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
      const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSW128);
      const uint64_t dV = make_smem_desc(smem_u32(sV), kHalfBytes, 1024, kLayoutSW128);
      constexpr uint32_t kTile16 = Cfg::kTileBytes / 16;
      auto issue_qk = [&](int q, int stage) {
        const uint64_t a = dQ + uint64_t(q * kTile16), bb = dK + uint64_t(stage * kTile16);
        const uint32_t d = tmem_base + uint32_t(Cfg::kSCol + q * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint32_t off = uint32_t(((kk / 4) * kHalfBytes + (kk % 4) * 32) / 16);
            umma_ss<1>(d, a + off, bb + off, idesc_qk, kk != 0);
          }
          umma_commit(&s_full[q]);
        }
      };
      auto issue_pv = [&](int q, int stage, int step, bool first_tile) {
        const uint64_t bb = dV + uint64_t(stage * kTile16);
        const uint32_t d = tmem_base + uint32_t(Cfg::kOCol + q * 128);
        const uint32_t a = tmem_base + uint32_t(Cfg::kSCol + q * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (((kk >> 1) & 1) != step) continue;  // step 0: keys 0..31 and 64..95 (kk 0,1,4,5); step 1: the rest
            umma_ts(d, a + uint32_t(kk * 8), bb + uint64_t(kk * 128), idesc_pv, !(first_tile && kk == 0));
          }
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      if (leader) umma_commit(&k_empty[0]);
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % kVS, ksn = (j + 1) % kKS;
        const bool more = j + 1 < n_kv;
        mbar_wait(&v_full[vs], (j / kVS) & 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          mbar_wait(&p_full[2 * q], j & 1);
          tc_fence_after();
          issue_pv(q, vs, 0, j == 0);
          mbar_wait(&p_full[2 * q + 1], j & 1);
          tc_fence_after();
          issue_pv(q, vs, 1, false);
          if (leader) {
            umma_commit(&pv_done[q]);
            if (q == 1) umma_commit(&v_empty[vs]);
          }
          if (more) {
            if (q == 0) {
              mbar_wait(&k_full[ksn], ((j + 1) / kKS) & 1);
              tc_fence_after();
            }
            issue_qk(q, ksn);
            if (q == 1 && leader) umma_commit(&k_empty[ksn]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    setmaxnreg_inc<Cfg8::kRegsLarge>();
    {
      // ===================== softmax: a pair of threads per query row, both tiles in turn =====================
      const int half = (warp - 4) >> 2;  // which 64 score columns / which half of the accumulator columns
      const int quad = warp & 3;         // TMEM lane quadrant = rows quad * 32 .. + 31 of a tile
      const int row_in_tile = quad * 32 + lane;
      const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
      const float c = p.scale_log2;
      const f32x2 c2 = pack2(c, c);
      float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
      for (int j = 0; j < n_kv; ++j) {
        const int valid = p.N - j * 128 - half * 64;  // columns of my half that are real keys (>= 64 except on the last tile)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
          const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128 + half * (kHeadDim / 2));
          float* xmine = xch + (j & 1) * 512 + (q * 2 + half) * 128 + row_in_tile;
          float* xother = xch + (j & 1) * 512 + (q * 2 + (half ^ 1)) * 128 + row_in_tile;
          mbar_wait(&s_full[q], j & 1);
          tc_fence_after();
          uint32_t s0[32], s1[32];
          tmem_ld32(t_s + half * 64, s0);
          tmem_ld32(t_s + half * 64 + 32, s1);
          tmem_ld_wait();
          if (valid < 64) {  // ragged last tile: keys past N score -inf -> probability 0
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i >= valid) s0[i] = 0xff800000u;
              if (32 + i >= valid) s1[i] = 0xff800000u;
            }
          }
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
            mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
          }
          *xmine = fmaxf(mx0, mx1);
          pair8_bar_sync(quad);  // also orders: both halves have read their scores before either writes P over them
          const float mx = fmaxf(fmaxf(mx0, mx1), *xother);
          const bool need = (mx - m[q]) * c > kAttnRescaleThreshold;  // true on the first tile (m = -inf); identical in both threads of a row
          const float m_new = need ? mx : m[q];
          const float alpha = need ? ex2((m[q] - m_new) * c) : 1.0f;
          if (j > 0 && __any_sync(0xffffffffu, need)) {
            // O_q holds PV(0..j-1), retired (s_full(j) flipped behind it).  Each half rescales its dh / 2 accumulator columns; PV(j)
            // of EITHER half adds into all columns, so neither half may hand its P over before both are done
#pragma unroll 1
            for (int cch = 0; cch < kHeadDim / 64; ++cch) {
              uint32_t v[32];
              tmem_ld32(t_o + cch * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(t_o + cch * 32, v);
            }
            tmem_st_wait();
            tc_fence_before();
            pair8_bar_sync(quad);  // the partner warp took the same branch: same rows, same maxima
          }
          const float mc = m_new * c;
          const f32x2 nmc2 = pack2(-mc, -mc);
          f32x2 sum2 = pack2(0.f, 0.f);
          uint32_t pk[16];
          attn_exp_quarter<kEmu>(s0, c2, nmc2, sum2, pk);
          tmem_st16(t_s + half * 32, pk);  // 32 probabilities as 16 bf16 pairs over the S columns
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * q]);
          attn_exp_quarter<kEmu>(s1, c2, nmc2, sum2, pk);
          tmem_st16(t_s + half * 32 + 16, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * q + 1]);
          float sum0, sum1;
          unpack2(sum2, sum0, sum1);
          l[q] = l[q] * alpha + (sum0 + sum1);
          m[q] = m_new;
        }
      }
      // ---- finalize: the two partial sums of a row meet, O / l -> bf16, each thread stores its dh / 2 columns token-major
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128 + half * (kHeadDim / 2));
        float* xmine = xch + (n_kv & 1) * 512 + (q * 2 + half) * 128 + row_in_tile;
        float* xother = xch + (n_kv & 1) * 512 + (q * 2 + (half ^ 1)) * 128 + row_in_tile;
        *xmine = l[q];
        pair8_bar_sync(quad);
        const float lsum = l[q] + *xother;
        mbar_wait(&pv_done[q], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / lsum;
        const int pos = q0 + q * 128 + row_in_tile;
        const bool row_ok = pos < p.N;
        const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
        __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim + half * (kHeadDim / 2);
#pragma unroll 1
        for (int cch = 0; cch < kHeadDim / 64; ++cch) {
          uint32_t v[32];
          tmem_ld32(t_o + cch * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float xo[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) xo[i] = __uint_as_float(v[i]) * inv_l;
            store_row_chunk_bf16x32(dst + cch * 32, xo);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
