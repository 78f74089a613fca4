// Joint [text;image] flash attention, CTA-pair schedule (sm_100a, cta_group::2), head_dim 128.
//
// The ping-pong schedules (attention.cuh, attention3.cuh) keep two query tiles per CTA in TMEM, so each tile has ONE
// score buffer and its chain  softmax(j) -> PV(j), QK(j+1) -> softmax(j+1)  is serial: the clock traces
// (profiles/r1j_attn_trace.md) put the tensor pipe at ~67 % busy with nothing left to shorten inside the chain.
// Here a CTA owns ONE 128-row query tile, which leaves room for THREE score buffers next to the O accumulator
// (3 x 128 + 128 = 512 TMEM columns): QK runs three KV tiles ahead of PV and the softmax never waits for the tensor pipe.
// Two CTAs of a cluster (adjacent query tiles of the same head) issue every MMA as one cta_group::2 instruction with
// M = 256, so each CTA stages only HALF of every K tile (64 keys) and half of every V tile (64 of the 128 dh columns):
// L2->smem traffic and B-operand smem reads per SM stay where they were with two query tiles per CTA.
//
// Measured (profiles/r1k_attention_pair.md): correct, but 15-25 % SLOWER than schedule 3: with both warpgroups always
// busy the two softmax warps of an SMSP contend for the MUFU / FMA / ALU pipes, and one 128 x 128 tile costs ~1450-2000
// SMSP cycles of softmax against 1024 cycles of MMA -- the kernel is bound by the softmax instruction stream, not by the
// serial chain this schedule removes.  Kept as attn_variant 7 (not the default).
//
// Per CTA: warp 0 TMA producer, warp 1 MMA issuer (pair leader only), warp 2 TMEM allocation; warpgroup 1 (warps 4-7)
// does the softmax of the even KV tiles, warpgroup 2 (warps 8-11) of the odd ones, one thread per query row.  The
// running (lazily rescaled) row maximum is handed from tile to tile through shared memory; each warpgroup keeps its own
// partial row sum relative to the reference it used last, and the two are merged at the end.
//   S(j)    = Q K_j^T     UMMA SS, M 256 x N 128 (each CTA supplies 64 key rows)   -> TMEM buffer j % 3 of both CTAs
//   P(j)    = 2^(S c - m c)  bf16, written over the first 64 columns of the same buffer
//   O      += P(j) V_j    UMMA TS, A = each CTA's own P, B = V (MN-major, each CTA supplies 64 dh columns)
#pragma once
#include <cuda.h>

#include "attention.cuh"
#include "attention3.cuh"
#include "ptx.cuh"

namespace tfx {

struct AttnPairCfg {
  static constexpr int kHeadDim = 128;
  static constexpr int kQBytes = 128 * kHeadDim * 2;      // this CTA's query tile: [2 dh halves][128 rows][64]
  static constexpr int kKBytes = 64 * kHeadDim * 2;       // half a K tile: [2 dh halves][64 keys][64]
  static constexpr int kVBytes = 128 * 64 * 2;            // half a V tile: [128 keys][64 dh columns]
  static constexpr int kKStages = 4;
  static constexpr int kVStages = 4;
  static constexpr int kThreads = 384;
  static constexpr int kXchBytes = 128 * 4 + 2 * 128 * 8;  // running max [128] + final (m, l) exchange [2][128]
  static constexpr int kSmemBytes = kQBytes + kKStages * kKBytes + kVStages * kVBytes + 1024 + 256 + kXchBytes;
  static constexpr int kOCol = 384;                        // S buffer b at b * 128, O at 384
  static constexpr int kRegsSmall = 72, kRegsLarge = 216;
};

template <int kEmu>
__global__ void __launch_bounds__(AttnPairCfg::kThreads, 1)
attention_pair_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = AttnPairCfg;
  constexpr int kHeadDim = Cfg::kHeadDim;
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;
  uint8_t* sV = sK + kKS * Cfg::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVS * Cfg::kVBytes);
  uint64_t* q_full = bars;               // [1]   leader: both CTAs' Q landed
  uint64_t* k_full = q_full + 1;         // [kKS] leader: both halves of a K tile landed
  uint64_t* k_empty = k_full + kKS;      // [kKS] both CTAs (multicast commit): stage may be refilled
  uint64_t* v_full = k_empty + kKS;      // [kVS] leader
  uint64_t* v_empty = v_full + kVS;      // [kVS] both CTAs
  uint64_t* s_full = v_empty + kVS;      // [3]   both CTAs: score buffer written
  uint64_t* p_full = s_full + 3;         // [3]   leader: P of both CTAs in place (8 warp arrivals)
  uint64_t* pv_done = p_full + 3;        // [2]   both CTAs: PV(j) retired, barrier j & 1 (a parity wait is only unambiguous
                                         //       within one phase: s_full(j) implies PV(j-3) retired, not PV(j-2))
  uint64_t* o_final = pv_done + 2;       // [1]   both CTAs: the last PV retired
  uint64_t* m_ready = o_final + 1;       // [2]   local: running max of tile j published (parity of j picks the barrier)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(m_ready + 2);
  float* m_sh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [128]
  float2* fin_sh = reinterpret_cast<float2*>(m_sh + 128);                          // [2 warpgroups][128] (m_ref, l)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int q0 = blockIdx.x * 128;
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
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); }
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_init(o_final, 1);
    mbar_init(&m_ready[0], 4);
    mbar_init(&m_ready[1], 4);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_base_ptr, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_wait();

  if (warp < 4) {
    setmaxnreg_dec<Cfg::kRegsSmall>();
    if (warp == 0) {
      // ===================== TMA producer (every CTA loads its halves; tx bytes complete on the leader) =====================
      const bool issuer = elect_one();
      if (issuer) {
        if (is_leader) mbar_arrive_expect_tx(q_full, 2 * Cfg::kQBytes);
        for (int h = 0; h < 2; ++h)
          tma_load_3d_2sm(&tmQ, q_full, sQ + h * (128 * 128), h * 64, q0, bh, kEvictFirst);
      }
      auto load_k = [&](int t) {
        const int ks = t % kKS;
        mbar_wait_cluster(&k_empty[ks], ((t / kKS) & 1) ^ 1);
        if (issuer) {
          if (is_leader) mbar_arrive_expect_tx(&k_full[ks], 2 * Cfg::kKBytes);
          for (int h = 0; h < 2; ++h)
            tma_load_3d_2sm(&tmK, &k_full[ks], sK + ks * Cfg::kKBytes + h * (64 * 128), h * 64, t * 128 + int(rank) * 64, bh, kEvictLast);
        }
      };
      auto load_v = [&](int t) {
        const int vs = t % kVS;
        mbar_wait_cluster(&v_empty[vs], ((t / kVS) & 1) ^ 1);
        if (issuer) {
          if (is_leader) mbar_arrive_expect_tx(&v_full[vs], 2 * Cfg::kVBytes);
          tma_load_3d_2sm(&tmV, &v_full[vs], sV + vs * Cfg::kVBytes, int(rank) * 64, t * 128, bh, kEvictLast);
        }
      };
      // same order as the issuer consumes: K runs three tiles ahead of V
      for (int t = 0; t < 3 && t < n_kv; ++t) load_k(t);
      for (int t = 0; t < n_kv; ++t) {
        load_v(t);
        if (t + 3 < n_kv) load_k(t + 3);
        __syncwarp();
      }
    } else if (warp == 1 && is_leader) {
      // ===================== MMA issuer (pair leader): warp-uniform control flow, one elected lane issues ==============
      constexpr uint32_t idesc_qk = make_idesc_bf16(256, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(256, kHeadDim, 0, 1);  // B = V is MN-major (dh contiguous)
      const bool issuer = elect_one();
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
      const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSW128);
      const uint64_t dV = make_smem_desc(smem_u32(sV), 64 * 128, 1024, kLayoutSW128);
      auto issue_qk = [&](int t) {
        const int ks = t % kKS, buf = t % 3;
        mbar_wait_cluster(&k_full[ks], (t / kKS) & 1);
        tc_fence_after();
        if (issuer) {
          const uint64_t bb = dK + uint64_t(ks * (Cfg::kKBytes / 16));
          const uint32_t d = tmem_base + uint32_t(buf * 128);
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint32_t qoff = uint32_t(((kk / 4) * (128 * 128) + (kk % 4) * 32) / 16);
            const uint32_t koff = uint32_t(((kk / 4) * (64 * 128) + (kk % 4) * 32) / 16);
            umma_ss<2>(d, dQ + qoff, bb + koff, idesc_qk, kk != 0);
          }
          umma_commit_2sm(&s_full[buf], 0b11);
          umma_commit_2sm(&k_empty[ks], 0b11);
        }
        __syncwarp();
      };
      mbar_wait_cluster(q_full, 0);
      for (int t = 0; t < 3 && t < n_kv; ++t) issue_qk(t);
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % kVS, buf = j % 3;
        mbar_wait_cluster(&v_full[vs], (j / kVS) & 1);
        mbar_wait_cluster(&p_full[buf], (j / 3) & 1);
        tc_fence_after();
        if (issuer) {
          const uint64_t bb = dV + uint64_t(vs * (Cfg::kVBytes / 16));
          const uint32_t d = tmem_base + uint32_t(Cfg::kOCol);
          const uint32_t a = tmem_base + uint32_t(buf * 128);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // 16 keys (= 2 KiB of this CTA's V half) per step
            umma_ts_2sm(d, a + uint32_t(kk * 8), bb + uint64_t(kk * 128), idesc_pv, (j | kk) != 0);
          umma_commit_2sm(&pv_done[j & 1], 0b11);
          umma_commit_2sm(&v_empty[vs], 0b11);
          if (j + 1 == n_kv) umma_commit_2sm(o_final, 0b11);
        }
        __syncwarp();
        if (j + 3 < n_kv) issue_qk(j + 3);  // overwrites buffer j % 3 = P(j): ordered behind PV(j) by the in-order tensor pipe
      }
    }
  } else {
    // ===================== softmax: warpgroup w takes KV tiles j = w, w + 2, ...; one thread per query row ==============
    setmaxnreg_inc<Cfg::kRegsLarge>();
    const int w = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol);
    const float c = p.scale_log2;
    const float kRescaleThreshold = 8.0f;  // log2 units: keep a stale row max until it is off by more than 2^8
    float m_ref = -INFINITY, l = 0.f;      // reference of my last tile, my partial row sum relative to it
    for (int j = w; j < n_kv; j += 2) {
      const int buf = j % 3;
      const uint32_t t_s = t_lane + uint32_t(buf * 128);
      const int valid = p.N - j * 128;  // >= 128 on every tile but possibly the last
      mbar_wait_cluster(&s_full[buf], (j / 3) & 1);
      tc_fence_after();
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
      // reference maximum after tile j-1, published by the other warpgroup
      float m_prev = -INFINITY;
      if (j > 0) {
        mbar_wait(&m_ready[(j - 1) & 1], ((j - 1) >> 1) & 1);
        m_prev = m_sh[row_in_tile];
      }
      // lazy rescaling: the reference only moves when the true max ran away from it
      const bool need = (mx - m_prev) * c > kRescaleThreshold;  // true on the first tile (m_prev = -inf)
      const float m_new = need ? mx : m_prev;
      if (j + 1 < n_kv) {
        m_sh[row_in_tile] = m_new;
        __syncwarp();
        if (lane == 0) mbar_arrive(&m_ready[j & 1]);
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // O holds PV(0..j-1) relative to m_prev; PV(j) cannot start before this warpgroup hands P(j) over
        const float alpha = need ? ex2((m_prev - m_new) * c) : 1.0f;
        mbar_wait_cluster(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
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
      if (m_new != m_ref) l *= ex2((m_ref - m_new) * c);  // bring my partial sum to the current reference (0 * 0 at start)
      m_ref = m_new;
      const float mc = m_new * c;
      const f32x2 c2 = pack2(c, c), nmc2 = pack2(-mc, -mc);
      f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int cch = half * 2 + cc;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const f32x2 x2 = fma2(pack2(__uint_as_float(sr[cch][i]), __uint_as_float(sr[cch][i + 1])), c2, nmc2);
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
        tmem_st32(t_s + half * 32, pk);  // P (bf16 pairs) over the S columns already in registers
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&p_full[buf], 0);
      float sum0, sum1;
      unpack2(sum2, sum0, sum1);
      l += sum0 + sum1;
    }
    // ---- merge the two partial row sums, then O / l -> bf16, token-major store (each warpgroup stores 64 dh columns)
    fin_sh[w * 128 + row_in_tile] = make_float2(m_ref, l);
    softmax_bar_sync();
    const float2 other = fin_sh[(w ^ 1) * 128 + row_in_tile];
    const float m_fin = fmaxf(m_ref, other.x);
    const float l_tot = l * ex2((m_ref - m_fin) * c) + other.y * ex2((other.x - m_fin) * c);
    const float inv_l = 1.0f / l_tot;
    mbar_wait_cluster(o_final, 0);
    tc_fence_after();
    const int pos = q0 + row_in_tile;
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim + w * 64;
#pragma unroll 1
    for (int cch = 0; cch < 2; ++cch) {
      uint32_t v[32];
      tmem_ld32(t_o + w * 64 + cch * 32, v);
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

  // ---- teardown: nobody leaves while the peer may still multicast into this CTA's barriers / read its smem
  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace tfx
