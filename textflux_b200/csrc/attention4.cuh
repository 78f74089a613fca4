// Joint [text;image] flash attention, schedule 4 ("stream"): schedule 3 (attention3.cuh: two 128-row query tiles per CTA
// ping-pong, S/P/O in TMEM, warp-uniform issuer, split P hand-over, setmaxnreg) plus a work decomposition that removes the
// wave-quantisation tail of the one-CTA-per-(head, query pair) grid.
//
// Work units are (batch, head, 256-query-row pair); each costs n_kv = ceil(N/128) KV iterations.  On 148 SMs with one
// CTA per SM the grid of schedule 3 runs ceil(units / 148) waves: 240 units at N = 2560 and 480 at N = 5120 both waste 19 %
// of the machine in a mostly empty last wave.  Here the first floor(units/148)*148 units run as before; the remaining R
// units are treated as ONE stream of R * n_kv iterations cut into equal contiguous shares, one per SM (the stream-K idea
// applied to the KV loop).  A share covers the tail of one unit and/or the head of the next, so it is launched as (up to)
// two CTAs, each a contiguous KV range [kv0, kv1) of one unit: the first segments of all shares start together, and the
// second segments are queued longest first (the host sorts them), so the SM whose first segment ends first picks up the
// longest second segment -- its own partner, since the two lengths of a share add up to the share.  A CTA that owns a whole unit stores the normalised output
// directly; a CTA that owns part of one stores its un-normalised (O, m, l) to a workspace, takes a ticket on the unit's
// counter, and the last arriver of the unit merges the parts in part order (deterministic) exactly as flash-decoding
// does:  M = max m_p,  O = sum_p 2^((m_p - M) c) O_p / sum_p 2^((m_p - M) c) l_p.
#pragma once
#include <cuda.h>

#include "attention3.cuh"

namespace tfx {

constexpr int kAttn4MaxParts = 8;
constexpr int kAttn4MaxShares = 160;  // >= SM count

struct Attn4Params {
  AttnParams a;
  int n_qpairs;       // ceil(N / 256)
  int n_units;        // B * H * n_qpairs
  int n_full;         // units [0, n_full) run whole, one CTA each
  int n_rem;          // R = n_units - n_full units are cut into shares
  int stream_ctas;    // number of shares (<= SM count); share c = stream iterations [c*share, min((c+1)*share, R*n_kv))
  int share;          // KV iterations per share (<= n_kv, so a share touches at most two units)
  int n_seg2;         // shares that cross a unit boundary; grid = n_full + stream_ctas + n_seg2
  uint16_t seg2_share[kAttn4MaxShares];  // those shares, longest second segment first
  float* ws_o;        // [stream_ctas][2 segments][2 q tiles][dh][128 rows] fp32, un-normalised O
  float* ws_ml;       // [stream_ctas][2][2][2 (m, l)][128]
  int* counters;      // [R] tickets, self-resetting
};

template <int kHeadDim>
struct Attn4Cfg : Attn3Cfg<kHeadDim> {
  static constexpr size_t kWsOFloats = 2ull * 2 * kHeadDim * 128;  // per share
  static constexpr size_t kWsMlFloats = 2ull * 2 * 2 * 128;
};

__device__ __forceinline__ float ldcg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int kHeadDim, int kEmu>
__global__ void __launch_bounds__(Attn3Cfg<kHeadDim>::kThreads, 1)
attention4_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ Attn4Params pp) {
  using Cfg = Attn3Cfg<kHeadDim>;
  const AttnParams& p = pp.a;
  constexpr int kHalves = Cfg::kHalves;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages;

  // ---- which piece of work is this CTA (block-uniform; decided before any barrier or TMEM allocation)
  const int n_kv_all = (p.N + 127) / 128;
  int unit, kv0 = 0, kv1 = n_kv_all, n_parts = 1, part = 0, rem_unit = -1, c_first = 0, which = 0;
  {
    const int idx = blockIdx.x;
    if (idx < pp.n_full) {
      unit = idx;
    } else {
      const int s = idx - pp.n_full;
      which = s >= pp.stream_ctas ? 1 : 0;
      const int c = which ? int(pp.seg2_share[s - pp.stream_ctas]) : s;
      const long long total = (long long)pp.n_rem * n_kv_all;
      const long long start = (long long)c * pp.share;
      const long long end = (start + pp.share < total) ? start + pp.share : total;
      if (start >= end) return;
      const int u0 = int(start / n_kv_all), u1 = int((end - 1) / n_kv_all);
      if (which == 0) {
        rem_unit = u0;
        kv0 = int(start - (long long)u0 * n_kv_all);
        const long long e = end - (long long)u0 * n_kv_all;
        kv1 = e < n_kv_all ? int(e) : n_kv_all;
      } else {
        if (u1 == u0) return;
        rem_unit = u1;
        kv0 = 0;
        kv1 = int(end - (long long)u1 * n_kv_all);
      }
      unit = pp.n_full + rem_unit;
      c_first = int(((long long)rem_unit * n_kv_all) / pp.share);
      const int c_last = int((((long long)rem_unit + 1) * n_kv_all - 1) / pp.share);
      n_parts = c_last - c_first + 1;
      part = c - c_first;
    }
  }
  const int n_it = kv1 - kv0;  // >= 1
  const int qp = unit % pp.n_qpairs;
  const int bh = unit / pp.n_qpairs;
  const int b = bh / p.H, head = bh - b * p.H;
  const int q0 = qp * 256;

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
  uint64_t* h0_done = pv_done + 2;       // [2]  kAttnNoMax: the PV MMAs over keys 0..63 of the tile have retired
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(h0_done + 2);
  int* last_flag = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
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
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&h0_done[i], 1);
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
        mbar_arrive_expect_tx(q_full, 2 * Cfg::kTileBytes);
        for (int q = 0; q < 2; ++q)
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmQ, q_full, sQ + q * Cfg::kTileBytes + h * kHalfBytes, h * 64, q0 + q * 128, bh, kEvictFirst);
      }
      for (int jj = 0; jj < n_it; ++jj) {
        const int j = kv0 + jj;
        const int ks = jj % kKS, vs = jj % kVS;
        mbar_wait(&k_empty[ks], ((jj / kKS) & 1) ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&k_full[ks], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmK, &k_full[ks], sK + ks * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        mbar_wait(&v_empty[vs], ((jj / kVS) & 1) ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&v_full[vs], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmV, &v_full[vs], sV + vs * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
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
      for (int jj = 0; jj < n_it; ++jj) {
        const int vs = jj % kVS, ksn = (jj + 1) % kKS;
        const bool more = jj + 1 < n_it;
        mbar_wait(&v_full[vs], (jj / kVS) & 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          mbar_wait(&p_full[2 * q], jj & 1);
          tc_fence_after();
          issue_pv(q, vs, 0, 4, jj == 0);
          if (kAttnNoMax && leader) umma_commit(&h0_done[q]);
          mbar_wait(&p_full[2 * q + 1], jj & 1);
          tc_fence_after();
          issue_pv(q, vs, 4, 8, false);
          if (leader) {
            umma_commit(&pv_done[q]);
            if (q == 1) umma_commit(&v_empty[vs]);
          }
          if (more) {
            if (q == 0) {
              mbar_wait(&k_full[ksn], ((jj + 1) / kKS) & 1);
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
    setmaxnreg_inc<Cfg::kRegsLarge>();
    // ===================== softmax warpgroups: one thread per query row =====================
    const int q = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int pos = q0 + q * 128 + row_in_tile;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128);
    const float c = p.scale_log2;
    float m = -INFINITY, l = 0.f, pend = -INFINITY;
    for (int jj = 0; jj < n_it; ++jj) {
      const int valid = p.N - (kv0 + jj) * 128;  // >= 128 on every tile but possibly the last of the sequence
      mbar_wait(&s_full[q], jj & 1);
      tc_fence_after();
      auto handover = [&](int half, const uint32_t (&pk)[32]) {
        tmem_st32(t_s + half * 32, pk);  // P (bf16 pairs) over the S columns already in registers
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * q + half]);
      };
      auto wait_h0 = [&]() {
        mbar_wait(&h0_done[q], jj & 1);
        tc_fence_after();
      };
      attn_softmax_tile<kHeadDim, kEmu, true>(t_s, t_o, c, valid, jj == 0, m, l, pend, handover, [](int) {}, wait_h0);
    }
    mbar_wait(&pv_done[q], (n_it - 1) & 1);
    tc_fence_after();
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim;
    if (n_parts == 1) {
      // ---- whole unit: O / l -> bf16, token-major store
      const float inv_l = 1.0f / l;
#pragma unroll 1
      for (int cch = 0; cch < kHeadDim / 32; ++cch) {
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
    } else {
      // ---- part of a unit: park (O, m, l), take a ticket; the last arriver merges all parts in part order
      const int my_c = c_first + part;
      const int my_which = which;
      float* wo = pp.ws_o + ((size_t)(my_c * 2 + my_which) * 2 + q) * (size_t)(kHeadDim * 128);
      float* wml = pp.ws_ml + ((size_t)(my_c * 2 + my_which) * 2 + q) * 256;
#pragma unroll 1
      for (int cch = 0; cch < kHeadDim / 32; ++cch) {
        uint32_t v[32];
        tmem_ld32(t_o + cch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) wo[(cch * 32 + i) * 128 + row_in_tile] = __uint_as_float(v[i]);  // [d][row]: coalesced per d
      }
      wml[row_in_tile] = m;
      wml[128 + row_in_tile] = l;
      __threadfence();
      softmax_bar_sync();
      if (threadIdx.x == 128) {
        const int old = atomicAdd(pp.counters + rem_unit, 1);
        const int last = (old == n_parts - 1) ? 1 : 0;
        if (last) pp.counters[rem_unit] = 0;  // every part has arrived: ready for the next launch
        *last_flag = last;
      }
      softmax_bar_sync();
      if (*last_flag) {
        __threadfence();
        float mp[kAttn4MaxParts], sc[kAttn4MaxParts];
        float M = -INFINITY;
#pragma unroll
        for (int pi = 0; pi < kAttn4MaxParts; ++pi) {
          if (pi < n_parts) {
            // part pi of this unit is the share c_first + pi; it is that share's first segment unless the share began in the
            // previous unit, i.e. unless it is the unit's first part and does not start exactly at the unit boundary
            const int cc = c_first + pi;
            const bool second = (pi == 0) && ((long long)cc * pp.share < (long long)rem_unit * n_kv_all);
            const float* ml = pp.ws_ml + ((size_t)(cc * 2 + (second ? 1 : 0)) * 2 + q) * 256;
            mp[pi] = ldcg_f32(ml + row_in_tile);
            sc[pi] = ldcg_f32(ml + 128 + row_in_tile);  // l_p for now
            M = fmaxf(M, mp[pi]);
          }
        }
        float L = 0.f;
#pragma unroll
        for (int pi = 0; pi < kAttn4MaxParts; ++pi) {
          if (pi < n_parts) {
            const float w = ex2((mp[pi] - M) * c);
            L = fmaf(sc[pi], w, L);
            sc[pi] = w;
          }
        }
        const float inv_l = 1.0f / L;
#pragma unroll 1
        for (int cch = 0; cch < kHeadDim / 32; ++cch) {
          float acc[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
          for (int pi = 0; pi < kAttn4MaxParts; ++pi) {
            if (pi < n_parts) {
              const int cc = c_first + pi;
              const bool second = (pi == 0) && ((long long)cc * pp.share < (long long)rem_unit * n_kv_all);
              const float* po = pp.ws_o + ((size_t)(cc * 2 + (second ? 1 : 0)) * 2 + q) * (size_t)(kHeadDim * 128);
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[i] = fmaf(ldcg_f32(po + (cch * 32 + i) * 128 + row_in_tile), sc[pi], acc[i]);
            }
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] *= inv_l;
            store_row_chunk_bf16x32(dst + cch * 32, acc);
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
