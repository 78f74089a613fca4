// Joint [text;image] flash attention, schedule 6: schedule 3 (one CTA per (head, 256 query rows), two 128-row query tiles ping-pong,
// S / P / O in TMEM, split-P hand-over) with the softmax of each query tile spread over TWO warpgroups -- 20 warps per CTA, four
// softmax warps per scheduler instead of two.
//
// Why (profiles/r2n_sdpa_ncu.md, r1j_attn_trace.md): per KV tile the exponentials of one query tile are 776 XU cycles per scheduler,
// but one warp per scheduler only keeps the XU ~54 % busy (dependent MUFU -> FADD2 / F2FP chains, in-order issue), so they take
// ~1430 cycles and the S-ready -> P-ready chain bounds the kernel at ~3100 cycles per KV iteration for 2048 cycles of MMAs.  The
// library kernel runs the same instruction stream on 16 warps and reaches ~2470.  Here a row is handled by a PAIR of threads:
// warpgroup A of a tile takes score columns 0..63 (-> P half 0, accumulator columns 0..dh/2 in the rescale and the epilogue),
// warpgroup B columns 64..127.  The pair exchanges its half-row maxima through shared memory behind a 64-thread named barrier (one per
// (tile, lane quadrant)), keeps one shared reference maximum, and each thread sums its own half of the row; the two partial sums
// meet once, at the end.  P half h is handed over by warpgroup h alone (the split-P barriers of schedule 3 map one to one).
#pragma once
#include <cuda.h>

#include "attention3.cuh"

namespace tfx {

template <int kHeadDim>
struct Attn6Cfg {
  static constexpr int kThreads = 640;  // wg0: TMA, MMA, TMEM alloc, spare; wg1/wg2: softmax halves of q-tile 0; wg3/wg4: of q-tile 1
  static constexpr int kXchBytes = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] fp32: half-row maxima (and, at the end, sums)
  static constexpr int kSmemBytes = Attn3Cfg<kHeadDim>::kSmemBytes - Attn3Cfg<kHeadDim>::kXchBytes + kXchBytes;
  static constexpr int kRegsSmall = 64, kRegsLarge = 104;  // 128 * 64 + 512 * 104 = 640 * 96, the launch-time allocation
};

// 64-thread named barrier of the two warps that share the rows of (query tile q, lane quadrant): ids 1..8
__device__ __forceinline__ void pair_bar_sync(int q, int quad) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q * 4 + quad) : "memory"); }

template <int kHeadDim, int kEmu>
__global__ void __launch_bounds__(Attn6Cfg<kHeadDim>::kThreads, 1)
attention6_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn3Cfg<kHeadDim>;
  using Cfg6 = Attn6Cfg<kHeadDim>;
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
      mbar_init(&p_full[2 * i], 4);      // the four warps of the tile's warpgroup A
      mbar_init(&p_full[2 * i + 1], 4);  // ... of warpgroup B
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
    setmaxnreg_dec<Cfg6::kRegsSmall>();
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
      auto issue_pv = [&](int q, int stage, int kk0, int kk1, bool first_tile) {
        const uint64_t bb = dV + uint64_t(stage * kTile16);
        const uint32_t d = tmem_base + uint32_t(Cfg::kOCol + q * 128);
        const uint32_t a = tmem_base + uint32_t(Cfg::kSCol + q * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < kk0 || kk >= kk1) continue;
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
          issue_pv(q, vs, 0, 4, j == 0);
          mbar_wait(&p_full[2 * q + 1], j & 1);
          tc_fence_after();
          issue_pv(q, vs, 4, 8, false);
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
    setmaxnreg_inc<Cfg6::kRegsLarge>();
    {
      // ===================== softmax: a pair of threads per query row (warpgroup A: columns 0..63, B: 64..127) =====================
      const int sw = warp - 4;
      const int q = sw >> 3;           // query tile
      const int half = (sw >> 2) & 1;  // which 64 score columns / which half of the accumulator columns
      const int quad = warp & 3;       // TMEM lane quadrant = rows quad * 32 .. + 31 of the tile
      const int row_in_tile = quad * 32 + lane;
      const int pos = q0 + q * 128 + row_in_tile;
      const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
      const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
      const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128 + half * (kHeadDim / 2));
      float* xmine = xch + (q * 2 + half) * 128 + row_in_tile;         // + parity * 512
      float* xother = xch + (q * 2 + (half ^ 1)) * 128 + row_in_tile;
      const float c = p.scale_log2;
      const f32x2 c2 = pack2(c, c);
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        const int valid = p.N - j * 128 - half * 64;  // columns of my half that are real keys (>= 64 except on the last tile)
        mbar_wait(&s_full[q], j & 1);
        tc_fence_after();
        uint32_t sr[2][32];
        tmem_ld32(t_s + half * 64, sr[0]);
        tmem_ld32(t_s + half * 64 + 32, sr[1]);
        tmem_ld_wait();
        if (valid < 64) {  // ragged last tile: keys past N score -inf -> probability 0
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cc * 32 + i >= valid) sr[cc][i] = 0xff800000u;
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
        }
        // the row maximum of the tile = max of the two halves, exchanged through shared memory (slot by tile parity: the partner
        // cannot overwrite a slot before this thread has passed the next barrier)
        xmine[(j & 1) * 512] = fmaxf(mx0, mx1);
        pair_bar_sync(q, quad);  // also orders: both halves have read their scores before either writes P over them
        const float mx = fmaxf(fmaxf(mx0, mx1), xother[(j & 1) * 512]);
        const bool need = (mx - m) * c > kAttnRescaleThreshold;  // true on the first tile (m = -inf); identical in both threads of a row
        const float m_new = need ? mx : m;
        const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // O_q holds PV(0..j-1), retired (s_full(j) flipped behind it).  Each half rescales its dh / 2 accumulator columns; PV(j) of
          // EITHER half adds into all columns, so neither half may hand its P over before both are done
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
          pair_bar_sync(q, quad);  // the partner warp took the same branch: same rows, same maxima
        }
        const float mc = m_new * c;
        const f32x2 nmc2 = pack2(-mc, -mc);
        f32x2 sum2 = pack2(0.f, 0.f);
        uint32_t pk[32];
        attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
        tmem_st32(t_s + half * 32, pk);  // my 64 probabilities as 32 bf16 pairs: P half `half` over the S columns
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * q + half]);
        float sum0, sum1;
        unpack2(sum2, sum0, sum1);
        l = l * alpha + (sum0 + sum1);
        m = m_new;
      }
      // ---- finalize: the two partial sums of a row meet, O / l -> bf16, each thread stores its dh / 2 columns token-major
      xmine[(n_kv & 1) * 512] = l;
      pair_bar_sync(q, quad);
      l += xother[(n_kv & 1) * 512];
      mbar_wait(&pv_done[q], (n_kv - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l;
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

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
