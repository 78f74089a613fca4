// Joint [text;image] flash attention, schedule 5 ("persistent stream"): the inner loop of schedule 3 (attention3.cuh) run by
// ONE persistent CTA per SM that walks a list of work items, so that what a launch of schedule 3 loses outside its KV
// loop is overlapped or removed.  tools/attn_timeline.py on schedule 3 (profiles/r2c_attn_timeline.md): per CTA 0.3 us set-up
// + 1.5 us Q/K load + 0.5 us first QK + 2.1 us output store + 3.5 us between a CTA's exit and its successor's start = 8 us
// against 37 / 73 / 118 us of KV loop at N = 2560 / 5120 / 8704, and SMs busy 77 / 80 / 90 % of the kernel's span because
// 240 / 480 / 816 units do not divide by 148.
//
// Work decomposition (the one of attention4.cuh, made persistent): units are (batch, head, 256-query-row pair) costing
// n_kv = ceil(N/128) KV iterations each.  CTA c runs whole units c, c + G, c + 2G, ... (G = grid size = SM count: the same
// unit-to-time order as a one-CTA-per-unit grid, so the CTAs running at any moment share a few heads' K/V in L2), then
// its equal share of the R = units mod G remaining units, treated as one stream of R * n_kv iterations: iterations
// [c * share, (c + 1) * share), i.e. the tail of one unit and/or the head of the next.  Whole units store their normalised
// output; parts of a unit park un-normalised (O, m, l) in a workspace and the last arriver of the unit merges them in
// part order (deterministic, no CTA ever waits for another).
//
// Pipeline across items: the producer loads the next item's Q as soon as the last QK of the current item has retired
// (q_empty), its K/V tiles ride the same rings; the issuer puts the next item's first QK right behind the current item's
// last PV, so the tensor pipe never drains; a softmax warpgroup ends an item by pulling O out of TMEM into registers,
// releasing O (o_free) and only then normalising / storing, while the other query tile keeps the tensor pipe busy.
#pragma once
#include <cuda.h>

#include "attention3.cuh"
#include "attention4.cuh"

namespace tfx {

struct Attn5Params {
  AttnParams a;
  int n_qpairs;     // ceil(N / 256)
  int n_units;      // B * H * n_qpairs
  int grid;         // G: persistent CTAs
  int n_waves;      // whole units per CTA: units [0, n_waves * G)
  int n_rem;        // R = n_units - n_waves * G
  int share;        // KV iterations of the remainder stream per CTA (<= n_kv); 0 when R = 0
  int stream_ctas;  // CTAs with a non-empty share
  float* ws_o;      // [G][2 segments][2 q tiles][dh][128 rows] fp32, un-normalised O
  float* ws_ml;     // [G][2][2][2 (m, l)][128]
  int* counters;    // [R] tickets, self-resetting
};

struct Attn5Item {
  int unit, kv0, kv1;
  int n_parts, part, c_first, rem_unit, which;
};

// item k of CTA c; returns false past the end of the CTA's list (block-uniform)
__device__ __forceinline__ bool attn5_item(const Attn5Params& pp, int n_kv_all, int c, int k, Attn5Item& it) {
  it.kv0 = 0; it.kv1 = n_kv_all; it.n_parts = 1; it.part = 0; it.c_first = 0; it.rem_unit = -1; it.which = 0;
  if (k < pp.n_waves) {
    it.unit = k * pp.grid + c;
    return true;
  }
  const int which = k - pp.n_waves;
  if (which > 1 || c >= pp.stream_ctas) return false;
  const long long total = (long long)pp.n_rem * n_kv_all;
  const long long start = (long long)c * pp.share;
  const long long end = (start + pp.share < total) ? start + pp.share : total;
  const int u0 = int(start / n_kv_all), u1 = int((end - 1) / n_kv_all);
  if (which == 0) {
    it.rem_unit = u0;
    it.kv0 = int(start - (long long)u0 * n_kv_all);
    const long long e = end - (long long)u0 * n_kv_all;
    it.kv1 = e < n_kv_all ? int(e) : n_kv_all;
  } else {
    if (u1 == u0) return false;
    it.rem_unit = u1;
    it.kv0 = 0;
    it.kv1 = int(end - (long long)u1 * n_kv_all);
  }
  it.which = which;
  it.unit = pp.n_waves * pp.grid + it.rem_unit;
  it.c_first = int(((long long)it.rem_unit * n_kv_all) / pp.share);
  const int c_last = int((((long long)it.rem_unit + 1) * n_kv_all - 1) / pp.share);
  it.n_parts = c_last - it.c_first + 1;
  it.part = c - it.c_first;
  return true;
}

// number of KV iterations of item k of CTA c (0 past the end of the list): all the MMA issuer needs to know
__device__ __forceinline__ int attn5_len(const Attn5Params& pp, int n_kv_all, int c, int k) {
  Attn5Item it;
  return attn5_item(pp, n_kv_all, c, k, it) ? it.kv1 - it.kv0 : 0;
}

// kTrace builds (tools/attn_timeline5.py): cta_trace [CTA][item (8)][8] globaltimer ns -- 0 first scores of the
// item seen (q0 warpgroup), 1 its last P handed over, 2 its last PV retired, 3 O pulled out of TMEM, 4 item's output done,
// 5 issuer: first PV of the item issued, 6 producer: Q of the item requested, 7 iterations
template <int kHeadDim, int kEmu, bool kTrace = false>
__global__ void __launch_bounds__(Attn3Cfg<kHeadDim>::kThreads, 1)
attention5_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ Attn5Params pp) {
  using Cfg = Attn3Cfg<kHeadDim>;
  const AttnParams& p = pp.a;
  constexpr int kHalves = Cfg::kHalves;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages;
  const int n_kv_all = (p.N + 127) / 128;
  const int cta = blockIdx.x;
  long long* ctr = (kTrace && p.cta_trace) ? p.cta_trace + (long long)cta * 64 : nullptr;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2][kHalves][128][64]
  uint8_t* sK = sQ + 2 * Cfg::kTileBytes;          // [kKS][kHalves][128][64]
  uint8_t* sV = sK + kKS * Cfg::kTileBytes;        // [kVS][kHalves][128 kv][64 dh]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVS * Cfg::kTileBytes);
  uint64_t* q_full = bars;               // [1]  once per item
  uint64_t* q_empty = q_full + 1;        // [1]  once per item: every QK of the item has retired
  uint64_t* k_full = q_empty + 1;        // [kKS]
  uint64_t* k_empty = k_full + kKS;      // [kKS]
  uint64_t* v_full = k_empty + kKS;      // [kVS]
  uint64_t* v_empty = v_full + kVS;      // [kVS]
  uint64_t* s_full = v_empty + kVS;      // [2]  once per KV iteration
  uint64_t* p_full = s_full + 2;         // [2 q][2 halves]
  uint64_t* pv_done = p_full + 4;        // [2]  once per KV iteration
  uint64_t* o_free = pv_done + 2;        // [2]  once per item: O_q has been pulled into registers
  uint64_t* h0_done = o_free + 2;        // [2]  kAttnNoMax: the PV MMAs over keys 0..63 of the tile have retired
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
    mbar_init(q_empty, 1);
    for (int i = 0; i < kKS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kVS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_free[i], 4);
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
      int kt = 0;  // K/V tiles loaded so far: ring slot kt % stages, phase (kt / stages) & 1
      Attn5Item it;
      for (int k = 0; attn5_item(pp, n_kv_all, cta, k, it); ++k) {
        const int bh = it.unit / pp.n_qpairs;
        const int q0 = (it.unit - bh * pp.n_qpairs) * 256;
        if (k > 0) mbar_wait(q_empty, (k - 1) & 1);  // the previous item's QK MMAs no longer read sQ
        if (kTrace && ctr && leader && k < 8) ctr[k * 8 + 6] = (long long)globaltimer_ns();
        if (leader) {
          mbar_arrive_expect_tx(q_full, 2 * Cfg::kTileBytes);
          for (int q = 0; q < 2; ++q)
            for (int h = 0; h < kHalves; ++h)
              tma_load_3d(&tmQ, q_full, sQ + q * Cfg::kTileBytes + h * kHalfBytes, h * 64, q0 + q * 128, bh, kEvictFirst);
        }
        for (int j = it.kv0; j < it.kv1; ++j, ++kt) {
          const int ks = kt % kKS, vs = kt % kVS;
          mbar_wait(&k_empty[ks], ((kt / kKS) & 1) ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&k_full[ks], Cfg::kTileBytes);
            for (int h = 0; h < kHalves; ++h)
              tma_load_3d(&tmK, &k_full[ks], sK + ks * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
          }
          mbar_wait(&v_empty[vs], ((kt / kVS) & 1) ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&v_full[vs], Cfg::kTileBytes);
            for (int h = 0; h < kHalves; ++h)
              tma_load_3d(&tmV, &v_full[vs], sV + vs * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
          }
          __syncwarp();
        }
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
      int g = 0;  // KV iterations issued so far (all items): s_full / p_full / pv_done phase = g & 1, ring slot g % stages
      int n_it = attn5_len(pp, n_kv_all, cta, 0);
      if (n_it > 0) {
        // first QK of the first item
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
        if (leader) {
          umma_commit(&k_empty[0]);
          if (n_it == 1) umma_commit(q_empty);
        }
        __syncwarp();
      }
      for (int k = 0; n_it > 0; ++k) {
        const int next_len = attn5_len(pp, n_kv_all, cta, k + 1);  // 0: this is the CTA's last item
        for (int jj = 0; jj < n_it; ++jj, ++g) {
          const int vs = g % kVS, ksn = (g + 1) % kKS;
          const bool last = jj + 1 == n_it;
          const bool more = !last || next_len > 0;                            // a QK follows: next tile of this item, or tile 0 of the next
          const bool next_is_last_qk = last ? next_len == 1 : jj + 2 == n_it;  // ... and it is the last QK of the item it belongs to
          mbar_wait(&v_full[vs], (g / kVS) & 1);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (jj == 0 && k > 0) mbar_wait(&o_free[q], (k - 1) & 1);  // O_q of the previous item is in its owner's registers
            mbar_wait(&p_full[2 * q], g & 1);
            tc_fence_after();
            if (kTrace && ctr && leader && jj == 0 && q == 0 && k < 8) ctr[k * 8 + 5] = (long long)globaltimer_ns();
            issue_pv(q, vs, 0, 4, jj == 0);
            if (kAttnNoMax && leader) umma_commit(&h0_done[q]);
            mbar_wait(&p_full[2 * q + 1], g & 1);
            tc_fence_after();
            issue_pv(q, vs, 4, 8, false);
            if (leader) {
              umma_commit(&pv_done[q]);
              if (q == 1) umma_commit(&v_empty[vs]);
            }
            if (more) {
              if (q == 0) {
                if (last) mbar_wait(q_full, (k + 1) & 1);  // next item's Q
                mbar_wait(&k_full[ksn], ((g + 1) / kKS) & 1);
                tc_fence_after();
              }
              issue_qk(q, ksn);
              if (q == 1 && leader) {
                umma_commit(&k_empty[ksn]);
                if (next_is_last_qk) umma_commit(q_empty);
              }
            }
            __syncwarp();
          }
        }
        n_it = next_len;
      }
    }
  } else {
    setmaxnreg_inc<Cfg::kRegsLarge>();
    // ===================== softmax warpgroups: one thread per query row =====================
    const int q = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128);
    const float c = p.scale_log2;
    int g = 0;
    for (int k = 0;; ++k) {
      int n_it, kv_first;
      {
        Attn5Item it0;
        if (!attn5_item(pp, n_kv_all, cta, k, it0)) break;
        n_it = it0.kv1 - it0.kv0;
        kv_first = it0.kv0;
      }
      float m = -INFINITY, l = 0.f, pend = -INFINITY;
      for (int jj = 0; jj < n_it; ++jj, ++g) {
        const int valid = p.N - (kv_first + jj) * 128;  // >= 128 on every tile but possibly the last of the sequence
        mbar_wait(&s_full[q], g & 1);
        tc_fence_after();
        if (kTrace && ctr && jj == 0 && warp == 4 && lane == 0 && k < 8) { ctr[k * 8 + 0] = (long long)globaltimer_ns(); ctr[k * 8 + 7] = n_it; }
        auto handover = [&](int half, const uint32_t (&pk)[32]) {
          tmem_st32(t_s + half * 32, pk);  // P (bf16 pairs) over the S columns already in registers
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * q + half]);
        };
        auto wait_h0 = [&]() {
          mbar_wait(&h0_done[q], g & 1);
          tc_fence_after();
        };
        attn_softmax_tile<kHeadDim, kEmu, true>(t_s, t_o, c, valid, jj == 0, m, l, pend, handover, [](int) {}, wait_h0);
      }
      // ---- end of item: the last PV has to retire, then O_q leaves TMEM so that the next item's first PV may overwrite it
      if (kTrace && ctr && warp == 4 && lane == 0 && k < 8) ctr[k * 8 + 1] = (long long)globaltimer_ns();
      mbar_wait(&pv_done[q], (g - 1) & 1);
      tc_fence_after();
      if (kTrace && ctr && warp == 4 && lane == 0 && k < 8) ctr[k * 8 + 2] = (long long)globaltimer_ns();
      uint32_t o[kHeadDim / 32][32];
#pragma unroll
      for (int cch = 0; cch < kHeadDim / 32; ++cch) tmem_ld32(t_o + cch * 32, o[cch]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[q]);
      if (kTrace && ctr && warp == 4 && lane == 0 && k < 8) ctr[k * 8 + 3] = (long long)globaltimer_ns();

      Attn5Item it;  // recomputed here rather than kept live across the KV loop
      attn5_item(pp, n_kv_all, cta, k, it);
      const int bh = it.unit / pp.n_qpairs;
      const int b = bh / p.H, head = bh - b * p.H;
      const int pos = (it.unit - bh * pp.n_qpairs) * 256 + q * 128 + row_in_tile;
      const bool row_ok = pos < p.N;
      const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
      __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim;
      if (it.n_parts == 1) {
        // ---- whole unit: O / l -> bf16, token-major store
        const float inv_l = 1.0f / l;
        if (row_ok) {
#pragma unroll
          for (int cch = 0; cch < kHeadDim / 32; ++cch) {
            float xo[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) xo[i] = __uint_as_float(o[cch][i]) * inv_l;
            store_row_chunk_bf16x32(dst + cch * 32, xo);
          }
        }
      } else {
        // ---- part of a unit: park (O, m, l), take a ticket; the last arriver merges all parts in part order
        float* wo = pp.ws_o + ((size_t)(cta * 2 + it.which) * 2 + q) * (size_t)(kHeadDim * 128);
        float* wml = pp.ws_ml + ((size_t)(cta * 2 + it.which) * 2 + q) * 256;
#pragma unroll
        for (int cch = 0; cch < kHeadDim / 32; ++cch)
#pragma unroll
          for (int i = 0; i < 32; ++i) wo[(cch * 32 + i) * 128 + row_in_tile] = __uint_as_float(o[cch][i]);  // [d][row]: coalesced per d
        wml[row_in_tile] = m;
        wml[128 + row_in_tile] = l;
        __threadfence();
        softmax_bar_sync();
        if (threadIdx.x == 128) {
          const int old = atomicAdd(pp.counters + it.rem_unit, 1);
          const int last = (old == it.n_parts - 1) ? 1 : 0;
          if (last) pp.counters[it.rem_unit] = 0;  // every part has arrived: ready for the next launch
          *last_flag = last;
        }
        softmax_bar_sync();
        if (*last_flag) {
          __threadfence();
          float mp[kAttn4MaxParts], sc[kAttn4MaxParts];
          float M = -INFINITY;
#pragma unroll
          for (int pi = 0; pi < kAttn4MaxParts; ++pi) {
            if (pi < it.n_parts) {
              // part pi of this unit is share c_first + pi; it is that share's first segment unless the share began in the
              // previous unit, i.e. unless it is the unit's first part and does not start exactly at the unit boundary
              const int cc = it.c_first + pi;
              const bool second = (pi == 0) && ((long long)cc * pp.share < (long long)it.rem_unit * n_kv_all);
              const float* ml = pp.ws_ml + ((size_t)(cc * 2 + (second ? 1 : 0)) * 2 + q) * 256;
              mp[pi] = ldcg_f32(ml + row_in_tile);
              sc[pi] = ldcg_f32(ml + 128 + row_in_tile);  // l_p for now
              M = fmaxf(M, mp[pi]);
            }
          }
          float L = 0.f;
#pragma unroll
          for (int pi = 0; pi < kAttn4MaxParts; ++pi) {
            if (pi < it.n_parts) {
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
              if (pi < it.n_parts) {
                const int cc = it.c_first + pi;
                const bool second = (pi == 0) && ((long long)cc * pp.share < (long long)it.rem_unit * n_kv_all);
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
      if (kTrace && ctr && warp == 4 && lane == 0 && k < 8) ctr[k * 8 + 4] = (long long)globaltimer_ns();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
