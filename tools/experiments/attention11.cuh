// Joint [text;image] flash attention, schedule 11 (experiment): schedule 9 (one 128-row query tile per CTA, three score buffers, K and V
// on separate producer warps) with SIXTEEN softmax warps: two tile groups (even / odd KV tiles, as schedule 9) x two column halves (a
// pair of threads per row, as schedule 6).  Four softmax warps per scheduler hide each other's MUFU / FFMA2 / TMEM latencies; with the
// score buffers decoupling a tile's softmax from the tensor pipe, that latency hiding is all that is missing (profiles/r2ae_attn_sbuf3.md:
// schedule 9 is bound by softmax throughput, two warps per scheduler run at ~68 % of their XU / issue floor).  dh = 128 only.
#pragma once
#include <cuda.h>

#include "../../textflux_b200/csrc/attention3.cuh"

namespace tfx {

template <int kHeadDim>
struct Attn11Cfg {
  static constexpr int kTileBytes = 128 * kHeadDim * 2;
  static constexpr int kKStages = 3;  // = score buffers: K(j + 3) replaces K(j) once QK(j) has run
  static constexpr int kVStages = 2;
  static constexpr int kSBufs = 3;
  static constexpr int kThreads = 640;  // wg0: K TMA, MMA, TMEM alloc, V TMA; wg1..4: (tile group, column half) = (0,0) (0,1) (1,0) (1,1)
  static constexpr int kXchBytes = 14 * 128 * 4;  // half-row maxima [2][2][2], references [2], partial sums [2][2] per row
  static constexpr int kSmemBytes = (1 + kKStages + kVStages) * kTileBytes + 1024 + 256 + kXchBytes;
  static constexpr int kOCol = 384;  // S_b at b * 128 (P_b aliases its first 64 columns), O behind them
  static constexpr int kRegsSmall = 64, kRegsLarge = 104;  // 128 * 64 + 512 * 104 = 640 * 96, the launch-time allocation
};

#ifndef TFX_ATTN8
// 32 score columns -> 16 packed bf16 pairs (as attention8.cuh)
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
#endif

template <int kHeadDim, int kEmu, bool kSplitIssue = false, bool kTrace = false>
__global__ void __launch_bounds__(Attn11Cfg<kHeadDim>::kThreads, 1)
attention11_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn11Cfg<kHeadDim>;
  constexpr int kHalves = kHeadDim / 64;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages, kSB = Cfg::kSBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [kHalves][128][64]
  uint8_t* sK = sQ + Cfg::kTileBytes;              // [kKS][kHalves][128][64]
  uint8_t* sV = sK + kKS * Cfg::kTileBytes;        // [kVS][kHalves][128 kv][64 dh]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVS * Cfg::kTileBytes);
  uint64_t* q_full = bars;               // [1]
  uint64_t* k_full = q_full + 1;         // [kKS]
  uint64_t* k_empty = k_full + kKS;      // [kKS]
  uint64_t* v_full = k_empty + kKS;      // [kVS]
  uint64_t* v_empty = v_full + kVS;      // [kVS]
  uint64_t* s_full = v_empty + kVS;      // [kSB]
  uint64_t* p_full = s_full + kSB;       // [kSB][2 halves]
  uint64_t* pv_done = p_full + 2 * kSB;  // [kSB]: PV(j) retired, by score buffer (S(j) ready implies PV(j - 3) retired: no phase can be skipped unseen)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + kSB);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 wg][128] reference maxima, [2 wg][128] sums

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + head;
  const int n_kv = (p.N + 127) / 128;
  pdl_launch_dependents();
  // wait accounting (kTrace builds, one CTA in the middle of the grid): cycles each role spends inside its barrier waits
  const bool tracing = kTrace && p.trace != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
  long long acc[4] = {0, 0, 0, 0};
#define TFX_WAIT(slot, stmt)                     \
  do {                                           \
    if (kTrace && tracing) {                     \
      const long long t0_ = clock64();           \
      stmt;                                      \
      acc[slot] += clock64() - t0_;              \
    } else {                                     \
      stmt;                                      \
    }                                            \
  } while (0)

  if (warp == 3 && lane == 0) prefetch_tensormap(&tmV);
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kVS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < kSB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 4);      // column half 0: the four warps of that half
      mbar_init(&p_full[2 * i + 1], 4);
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
    setmaxnreg_dec<Cfg::kRegsSmall>();
    if (warp == 0) {
      // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
      const bool leader = elect_one();
      if (leader) {
        mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
        for (int h = 0; h < kHalves; ++h) tma_load_3d(&tmQ, q_full, sQ + h * kHalfBytes, h * 64, q0, bh, kEvictFirst);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % kKS;
        TFX_WAIT(0, mbar_wait(&k_empty[ks], ((j / kKS) & 1) ^ 1));
        if (leader) {
          mbar_arrive_expect_tx(&k_full[ks], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmK, &k_full[ks], sK + ks * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        __syncwarp();
      }
      if (tracing && lane == 0) p.trace[10] = acc[0];
    } else if (warp == 3) {
      // ===================== V producer: its own warp, so that a V slot waiting for PV(j - 2) never holds back K(j + 1) =====================
      const bool leader = elect_one();
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % kVS;
        TFX_WAIT(0, mbar_wait(&v_empty[vs], ((j / kVS) & 1) ^ 1));
        if (leader) {
          mbar_arrive_expect_tx(&v_full[vs], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmV, &v_full[vs], sV + vs * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        __syncwarp();
      }
      if (tracing && lane == 0) p.trace[11] = acc[0];
    } else if (warp == 1) {
      // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (dh contiguous)
      const bool leader = elect_one();
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
      const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSW128);
      const uint64_t dV = make_smem_desc(smem_u32(sV), kHalfBytes, 1024, kLayoutSW128);
      constexpr uint32_t kTile16 = Cfg::kTileBytes / 16;
      auto issue_qk = [&](int buf, int stage) {
        const uint64_t bb = dK + uint64_t(stage * kTile16);
        const uint32_t d = tmem_base + uint32_t(buf * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint32_t off = uint32_t(((kk / 4) * kHalfBytes + (kk % 4) * 32) / 16);
            umma_ss<1>(d, dQ + off, bb + off, idesc_qk, kk != 0);
          }
          umma_commit(&s_full[buf]);
          umma_commit(&k_empty[stage]);
        }
      };
      auto issue_pv = [&](int buf, int stage, int kk0, int kk1, bool first_tile) {
        const uint64_t bb = dV + uint64_t(stage * kTile16);
        const uint32_t d = tmem_base + uint32_t(Cfg::kOCol);
        const uint32_t a = tmem_base + uint32_t(buf * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < kk0 || kk >= kk1) continue;
            umma_ts(d, a + uint32_t(kk * 8), bb + uint64_t(kk * 128), idesc_pv, !(first_tile && kk == 0));
          }
        }
      };
      const long long t_role = clock64();
      if (kSplitIssue) {
        // PV issuer only: the QKs come from warp 2 (below), ordered against the PVs through pv_done / s_full instead of program order
        for (int j = 0; j < n_kv; ++j) {
          const int buf = j % kSB, vs = j % kVS;
          const uint32_t ph = (j / kSB) & 1;
          TFX_WAIT(0, mbar_wait(&v_full[vs], (j / kVS) & 1));
          TFX_WAIT(1, mbar_wait(&p_full[2 * buf], ph));
          tc_fence_after();
          issue_pv(buf, vs, 0, 4, j == 0);
          TFX_WAIT(2, mbar_wait(&p_full[2 * buf + 1], ph));
          tc_fence_after();
          issue_pv(buf, vs, 4, 8, false);
          if (leader) {
            umma_commit(&pv_done[buf]);
            umma_commit(&v_empty[vs]);
          }
          __syncwarp();
        }
        if (tracing && lane == 0) { p.trace[1] = acc[0]; p.trace[2] = acc[1]; p.trace[3] = acc[2]; p.trace[0] = clock64() - t_role; }
      } else {
      mbar_wait(q_full, 0);
      for (int j = 0; j < kSB && j < n_kv; ++j) {  // K stage = score buffer = j % 3
        mbar_wait(&k_full[j], 0);
        tc_fence_after();
        issue_qk(j, j);
        __syncwarp();
      }
      for (int j = 0; j < n_kv; ++j) {
        const int buf = j % kSB, vs = j % kVS;
        const uint32_t ph = (j / kSB) & 1;
        mbar_wait(&v_full[vs], (j / kVS) & 1);
        mbar_wait(&p_full[2 * buf], ph);
        tc_fence_after();
        issue_pv(buf, vs, 0, 4, j == 0);
        mbar_wait(&p_full[2 * buf + 1], ph);
        tc_fence_after();
        issue_pv(buf, vs, 4, 8, false);
        if (leader) {
          umma_commit(&pv_done[buf]);
          umma_commit(&v_empty[vs]);
        }
        if (j + kSB < n_kv) {  // the buffer PV(j) has just read takes the scores of tile j + 3
          mbar_wait(&k_full[buf], ((j + kSB) / kKS) & 1);
          tc_fence_after();
          issue_qk(buf, buf);
        }
        __syncwarp();
      }
      }
    } else if (warp == 2 && kSplitIssue) {
      // ===================== QK issuer (split-issue variant): QK(j) into buffer j % 3 once PV(j - 3) has retired =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      const bool leader = elect_one();
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
      const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSW128);
      constexpr uint32_t kTile16 = Cfg::kTileBytes / 16;
      mbar_wait(q_full, 0);
      const long long t_qk = clock64();
      for (int j = 0; j < n_kv; ++j) {
        const int buf = j % kSB;
        if (j >= kSB) TFX_WAIT(0, mbar_wait(&pv_done[buf], ((j - kSB) / kSB) & 1));
        TFX_WAIT(1, mbar_wait(&k_full[buf], (j / kKS) & 1));
        tc_fence_after();
        const uint64_t bb = dK + uint64_t(buf * kTile16);
        const uint32_t d = tmem_base + uint32_t(buf * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint32_t off = uint32_t(((kk / 4) * kHalfBytes + (kk % 4) * 32) / 16);
            umma_ss<1>(d, dQ + off, bb + off, idesc_qk, kk != 0);
          }
          umma_commit(&s_full[buf]);
          umma_commit(&k_empty[buf]);
        }
        __syncwarp();
      }
      if (tracing && lane == 0) { p.trace[5] = acc[0]; p.trace[6] = acc[1]; p.trace[12] = clock64() - t_qk; p.trace[13] = n_kv; }
    }
  } else {
    setmaxnreg_inc<Cfg::kRegsLarge>();
    // ===================== softmax: 16 warps.  Tile group g (even / odd KV tiles) x column half h (score columns h * 64 ..) ==========
    // A row of a tile is handled by a PAIR of threads (same lane quadrant, same scheduler): half-row maxima meet through shared memory.
    const int sw = warp - 4;
    const int g = sw >> 3;           // tile group: KV tiles j = g, g + 2, ...
    const int half = (sw >> 2) & 1;  // which 64 score columns, which P half, which accumulator columns in a rescale
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol);
    float* xh = xch + row_in_tile;                 // [2 parity][2 g][2 half][128]: half-row maxima of the tile in flight
    float* xm = xch + 8 * 128 + row_in_tile;       // [2 g][128]: reference after the group's latest tile
    float* xl = xch + 10 * 128 + row_in_tile;      // [2 g][2 half][128]: partial sums (final exchange)
    const int bar_pair = 1 + g * 4 + quad;         // 96 threads: the pair (sync) + the other group's publishing warp (arrive)
    const int bar_first = 9 + quad;                // 64 threads: tile 0 has no publisher
    const int bar_pub = 1 + (g ^ 1) * 4 + quad;
    const float c = p.scale_log2;
    const f32x2 c2 = pack2(c, c);
    float m = -INFINITY, l = 0.f;
    for (int j = g; j < n_kv; j += 2) {
      const int buf = j % kSB;
      const uint32_t t_s = t_lane + uint32_t(buf * 128);
      const int valid = p.N - j * 128 - half * 64;
      const int par = (j >> 1) & 1;
      mbar_wait(&s_full[buf], (j / kSB) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(t_s + half * 64, s0);
      tmem_ld32(t_s + half * 64 + 32, s1);
      tmem_ld_wait();
      if (valid < 64) {
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
      xh[((par * 2 + g) * 2 + half) * 128] = fmaxf(mx0, mx1);
      // pair barrier: half maxima exchanged, both halves hold their scores in registers (P may now overwrite them), and -- from tile 1
      // on -- the other group's reference after tile j - 1 has been published
      if (j == 0) asm volatile("bar.sync %0, 64;" ::"r"(bar_first) : "memory");
      else asm volatile("bar.sync %0, 96;" ::"r"(bar_pair) : "memory");
      const float mx = fmaxf(fmaxf(mx0, mx1), xh[((par * 2 + g) * 2 + (half ^ 1)) * 128]);
      if (j > 0) {
        const float mp = xm[(g ^ 1) * 128];
        if (mp != m) {
          l *= ex2((m - mp) * c);
          m = mp;
        }
      }
      const bool need = (mx - m) * c > kAttnRescaleThreshold;  // identical in both threads of a row
      const float m_new = need ? mx : m;
      const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        mbar_wait(&pv_done[(j - 1) % kSB], ((j - 1) / kSB) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cch = 0; cch < kHeadDim / 64; ++cch) {  // my half of the accumulator columns
          uint32_t v[32];
          tmem_ld32(t_o + half * (kHeadDim / 2) + cch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(t_o + half * (kHeadDim / 2) + cch * 32, v);
        }
        tmem_st_wait();
        tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(bar_first) : "memory");  // both halves of O rescaled before either hands P over / publishes
      }
      l *= alpha;
      m = m_new;
      if (half == 0 && j + 1 < n_kv) {
        xm[g * 128] = m;
        __threadfence_block();
        asm volatile("bar.arrive %0, 96;" ::"r"(bar_pub) : "memory");
      }
      const float mc = m * c;
      const f32x2 nmc2 = pack2(-mc, -mc);
      f32x2 sum2 = pack2(0.f, 0.f);
      {
        uint32_t pk[16];
        attn_exp_quarter<kEmu>(s0, c2, nmc2, sum2, pk);
        tmem_st16(t_s + half * 32, pk);
        attn_exp_quarter<kEmu>(s1, c2, nmc2, sum2, pk);
        tmem_st16(t_s + half * 32 + 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * buf + half]);
      float sum0, sum1;
      unpack2(sum2, sum0, sum1);
      l += sum0 + sum1;
    }
    // ---- finalize: every partial sum against the final reference (the group of the last tile holds it); each warpgroup stores 32 columns
    const int last = n_kv - 1;
    const int g_last = last & 1;
    if (half == 0) xm[g * 128] = m;
    xl[(g * 2 + half) * 128] = l;
    if (g == g_last) {
      mbar_wait(&pv_done[last % kSB], (last / kSB) & 1);
      tc_fence_after();
      tc_fence_before();
    }
    __threadfence_block();
    asm volatile("bar.sync 13, 512;" ::: "memory");
    tc_fence_after();
    const float m_fin = xm[g_last * 128];
    float lsum = 0.f;
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      const float f = ex2((xm[gg * 128] - m_fin) * c);
      lsum += (xl[(gg * 2) * 128] + xl[(gg * 2 + 1) * 128]) * f;
    }
    const float inv_l = 1.0f / lsum;
    const int pos = q0 + row_in_tile;
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    const int cch = g * 2 + half;  // 4 warpgroups x 32 columns = dh 128
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim + cch * 32;
    {
      uint32_t v[32];
      tmem_ld32(t_o + cch * 32, v);
      tmem_ld_wait();
      if (row_ok) {
        float xo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i] = __uint_as_float(v[i]) * inv_l;
        store_row_chunk_bf16x32(dst, xo);
      }
    }
  }

#undef TFX_WAIT
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
