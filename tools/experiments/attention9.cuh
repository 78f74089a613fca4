// Joint [text;image] flash attention, schedule 9 (experiment): ONE 128-row query tile per CTA, THREE score buffers, the two softmax
// warpgroups alternate over the KV tiles of that one query tile.
//
// Why (profiles/r2q_attn_rowsplit.md, r2s_attn_dbuf.md, r2ad_attn_timemux.md): schedules 3 / 5 are bound by the chain
// S ready -> softmax -> P -> PV(j) -> QK(j+1) -> S ready of a query tile, because P aliases the only score buffer the tile has
// (2 x S + 2 x O fill the 512 TMEM columns), and two tiles in flight cover only ~70 % of the tensor pipe.  With one query tile per
// CTA the same 512 columns hold O + three score buffers: QK(j+1), QK(j+2) are queued while the softmax of tile j runs, a softmax
// warpgroup takes every second KV tile (so a tile's ~1500-cycle softmax latency is hidden behind the other warpgroup's tile), and the
// tensor pipe only ever waits when the softmax THROUGHPUT (two warps per scheduler) falls behind.
// What it costs: K / V tiles are streamed per 128 query rows instead of per 256 (twice the L2 -> shared-memory traffic per FLOP), and
// the two warpgroups share one running reference maximum per row (a 64-thread producer / consumer named barrier per KV tile and lane
// quadrant; the partial row sums meet once at the end).
#pragma once
#include <cuda.h>

#include "../../textflux_b200/csrc/attention3.cuh"

namespace tfx {

template <int kHeadDim>
struct Attn9Cfg {
  static constexpr int kTileBytes = 128 * kHeadDim * 2;
  static constexpr int kKStages = 3;  // = score buffers: K(j + 3) replaces K(j) once QK(j) has run
  static constexpr int kVStages = kHeadDim == 128 ? 3 : 2;  // at dh = 128 this fills the 227 KB only without the 1 KB alignment slack (checked at run time)
  static constexpr int kSBufs = 3;
  static constexpr int kThreads = 384;  // wg0: TMA, MMA, TMEM alloc, spare; wg1: even KV tiles; wg2: odd KV tiles
  static constexpr int kXchBytes = 4 * 128 * 4;  // [2] reference maxima + [2] partial sums per row
  static constexpr int kSmemBytes = (1 + kKStages + kVStages) * kTileBytes + (kHeadDim == 128 ? 0 : 1024) + 256 + kXchBytes;
  static constexpr int kOCol = 384;  // S_b at b * 128 (P_b aliases its first 64 columns), O behind them
  static constexpr int kRegsSmall = Attn3Cfg<kHeadDim>::kRegsSmall, kRegsLarge = Attn3Cfg<kHeadDim>::kRegsLarge;
};

// producer / consumer named barriers between the two warps that own the same 32 rows (ids 1..8: 1 + quad * 2 + publishing warpgroup),
// and a plain one for the final exchange (ids 9..12)
__device__ __forceinline__ void pair9_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair9_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int kHeadDim, int kEmu, bool kSplitIssue = false, bool kTrace = false>
__global__ void __launch_bounds__(Attn9Cfg<kHeadDim>::kThreads, 1)
attention9_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn9Cfg<kHeadDim>;
  constexpr int kHalves = kHeadDim / 64;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages, kSB = Cfg::kSBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (kHeadDim == 128 && smem != smem_raw) __trap();  // no slack at dh = 128: the dynamic shared memory window must start 1 KB-aligned
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
      mbar_init(&p_full[2 * i], 4);
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
    // ===================== softmax: warpgroup w takes KV tiles j = w, w + 2, ...; one thread per query row =====================
    const int w = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol);
    float* xm_mine = xch + w * 128 + row_in_tile;
    float* xm_other = xch + (w ^ 1) * 128 + row_in_tile;
    float* xl_mine = xch + 256 + w * 128 + row_in_tile;
    float* xl_other = xch + 256 + (w ^ 1) * 128 + row_in_tile;
    const int bar_pub = 1 + quad * 2 + w, bar_sub = 1 + quad * 2 + (w ^ 1);
    const float c = p.scale_log2;
    const f32x2 c2 = pack2(c, c);
    float m = -INFINITY, l = 0.f;  // m: the reference the row's probabilities (both warpgroups') and O are expressed against
    *xm_mine = -INFINITY;
    const long long t_sm = clock64();
    for (int j = w; j < n_kv; j += 2) {
      const int buf = j % kSB;
      const uint32_t t_s = t_lane + uint32_t(buf * 128);
      const int valid = p.N - j * 128;
      TFX_WAIT(0, mbar_wait(&s_full[buf], (j / kSB) & 1));
      tc_fence_after();
      uint32_t sr[4][32];
      attn_load_scores(t_s, valid, sr);
      const float mx = attn_row_max(sr);
      if (j > 0) {  // the reference after tile j - 1 (the other warpgroup's): it has already rescaled O if it moved it
        TFX_WAIT(1, pair9_sync(bar_sub));
        const float mp = *xm_other;
        if (mp != m) {
          l *= ex2((m - mp) * c);
          m = mp;
        }
      }
      const bool need = (mx - m) * c > kAttnRescaleThreshold;  // true on the first tile (m = -inf)
      const float m_new = need ? mx : m;
      const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // nothing past PV(j - 1) can be in flight (PV(j) needs this tile's P): wait for it, then O is ours
        mbar_wait(&pv_done[(j - 1) % kSB], ((j - 1) / kSB) & 1);
        tc_fence_after();
        attn_rescale_o<kHeadDim>(t_o, alpha);
        tc_fence_before();
      }
      l *= alpha;
      m = m_new;
      if (j + 1 < n_kv) {  // publish before the exponentials: the other warpgroup's tile j + 1 only waits for this
        *xm_mine = m;
        __threadfence_block();
        pair9_arrive(bar_pub);
      }
      const float mc = m * c;
      const f32x2 nmc2 = pack2(-mc, -mc);
      f32x2 sum2 = pack2(0.f, 0.f);
      uint32_t pk[32];
      attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
      tmem_st32(t_s, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * buf]);
      attn_exp_half<kEmu>(sr[2], sr[3], c2, nmc2, sum2, pk);
      tmem_st32(t_s + 32, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * buf + 1]);
      float sum0, sum1;
      unpack2(sum2, sum0, sum1);
      l += sum0 + sum1;
    }
    if (tracing && quad == 0 && lane == 0) { p.trace[7 + 7 * w] = acc[0]; p.trace[8 + 7 * w] = acc[1]; p.trace[9 + 7 * w] = clock64() - t_sm; }
    // ---- finalize: the warpgroup of the last tile holds the final reference; the partial sums meet; each warpgroup stores dh / 2 columns
    const int last = n_kv - 1;
    const bool mine_last = (last & 1) == w;
    *xm_mine = m;
    *xl_mine = l;
    if (mine_last) {
      mbar_wait(&pv_done[last % kSB], (last / kSB) & 1);
      tc_fence_after();
      tc_fence_before();
    }
    __threadfence_block();
    pair9_sync(9 + quad);
    tc_fence_after();
    const float m_o = *xm_other, l_o = *xl_other;
    const float m_fin = mine_last ? m : m_o;
    const float lsum = mine_last ? l + l_o * ex2((m_o - m_fin) * c) : l * ex2((m - m_fin) * c) + l_o;
    const float inv_l = 1.0f / lsum;
    const int pos = q0 + row_in_tile;
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim + w * (kHeadDim / 2);
#pragma unroll 1
    for (int cch = 0; cch < kHeadDim / 64; ++cch) {
      uint32_t v[32];
      tmem_ld32(t_o + w * (kHeadDim / 2) + cch * 32, v);
      tmem_ld_wait();
      if (row_ok) {
        float xo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i] = __uint_as_float(v[i]) * inv_l;
        store_row_chunk_bf16x32(dst + cch * 32, xo);
      }
    }
  }

#undef TFX_WAIT
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
