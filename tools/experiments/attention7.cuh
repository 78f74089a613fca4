// Joint [text;image] flash attention, schedule 7: schedule 3 (one CTA per (head, 256 query rows), two 128-row query tiles, S / P / O in
// TMEM, 12 warps) walked in 64-key steps with the score buffer of each query tile DOUBLE-BUFFERED.
//
// Why (profiles/r1j_attn_trace.md, r2n_sdpa_ncu.md): with one 128-column score buffer per query tile, P aliases S, so per tile the work
// is a serial chain  S ready -> softmax -> P -> PV(j) -> QK(j+1) -> S ready ...: ~1850 cycles of softmax plus ~1200 of MMAs per KV
// tile, and with two tiles in flight the tensor pipe is busy 2048 of ~3100 cycles.  TMEM has no room for a second 128-column buffer
// (2 x S + 2 x O = 512 columns) -- but it has room for TWO 64-column buffers per tile.  Here QK(s + 2) of a tile is issued right behind
// PV(s) into the buffer PV(s) has just read, while the tile's softmax is already working on step s + 1 in the other buffer: the softmax
// never waits for a QK that has not been issued yet, and the MMA issuer never waits for a softmax that could not have started.
//   per 64-key step and tile:  QK = 8 (dh / 16) MMAs of 128 x 64 x 16 (32 cycles each), PV = 4 MMAs of 128 x dh x 16 (64 cycles each)
//   TMEM: tile q: S/P buffers at q * 128 + {0, 64}, O at 256 + q * 128.   K / V travel as 128-key TMA tiles as before.
// The lazily rescaled maximum, the exponentials and the hand-over are attention_common.cuh's, on 64 columns per step.
#pragma once
#include <cuda.h>

#include "attention3.cuh"

namespace tfx {

template <int kHeadDim, int kEmu, bool kTrace = false>
__global__ void __launch_bounds__(Attn3Cfg<kHeadDim>::kThreads, 1)
attention7_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn3Cfg<kHeadDim>;
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
  uint64_t* s_full = v_empty + kVS;      // [2 q][2 buffers]
  uint64_t* p_full = s_full + 4;         // [2 q][2 buffers]
  uint64_t* pv_done = p_full + 4;        // [2 q]: one completion per step
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + head;
  const int n_kv = (p.N + 127) / 128;    // 128-key K / V tiles
  const int n_steps = (p.N + 63) / 64;   // 64-key steps (the second half of a ragged last tile may be empty: skipped)
  // trace builds: clock64 stamps of CTA (0,0,0) per (step, tile): 0 s_full seen, 1 scores in registers, 2 maximum known, 3 P handed
  // over, 5 issuer saw P, 6 PV issued, 7 next QK issued (tools/attn_trace7.py)
  const bool tracing = kTrace && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kVS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
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
      // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (dh contiguous)
      const bool leader = elect_one();
      const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
      const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSW128);
      const uint64_t dV = make_smem_desc(smem_u32(sV), kHalfBytes, 1024, kLayoutSW128);
      constexpr uint32_t kTile16 = Cfg::kTileBytes / 16;
      // S_q[buffer sub] = Q_q K(stage)[keys sub * 64 .. + 63]^T
      auto issue_qk = [&](int q, int stage, int sub) {
        const uint64_t a = dQ + uint64_t(q * kTile16), bb = dK + uint64_t(stage * kTile16 + sub * (64 * 128 / 16));
        const uint32_t d = tmem_base + uint32_t(Cfg::kSCol + q * 128 + sub * 64);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint32_t off = uint32_t(((kk / 4) * kHalfBytes + (kk % 4) * 32) / 16);
            umma_ss<1>(d, a + off, bb + off, idesc_qk, kk != 0);
          }
          umma_commit(&s_full[2 * q + sub]);
        }
      };
      // O_q += P_q[buffer sub] V(stage)[keys sub * 64 .. + 63]
      auto issue_pv = [&](int q, int stage, int sub, bool first) {
        const uint64_t bb = dV + uint64_t(stage * kTile16 + sub * 4 * 128);
        const uint32_t d = tmem_base + uint32_t(Cfg::kOCol + q * 128);
        const uint32_t a = tmem_base + uint32_t(Cfg::kSCol + q * 128 + sub * 64);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // 16 kv rows (= 2 KiB of V, 8 packed columns of P) per MMA
            umma_ts(d, a + uint32_t(kk * 8), bb + uint64_t(kk * 128), idesc_pv, !(first && kk == 0));
          umma_commit(&pv_done[q]);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      // steps 0 and 1 (both halves of K tile 0) fill the two buffers of either tile
      issue_qk(0, 0, 0);
      issue_qk(1, 0, 0);
      if (n_steps > 1) {
        issue_qk(0, 0, 1);
        issue_qk(1, 0, 1);
      }
      if (leader) umma_commit(&k_empty[0]);
      __syncwarp();
      for (int s = 0; s < n_steps; ++s) {
        const int j = s >> 1, sub = s & 1, vs = j % kVS;
        const uint32_t par = (s >> 1) & 1;  // a buffer's barriers complete once per two steps
        if (sub == 0) mbar_wait(&v_full[vs], (j / kVS) & 1);
        const int s2 = s + 2, j2 = s2 >> 1, ks2 = j2 % kKS;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          mbar_wait(&p_full[2 * q + sub], par);
          tc_fence_after();
          if (tracing && leader) p.trace[(s * 2 + q) * kAttnTraceSlots + 5] = clock64();
          issue_pv(q, vs, sub, s == 0);
          if (tracing && leader) p.trace[(s * 2 + q) * kAttnTraceSlots + 6] = clock64();
          if (q == 1 && leader && (sub == 1 || s == n_steps - 1)) umma_commit(&v_empty[vs]);
          if (s2 < n_steps) {
            if (q == 0 && sub == 0) {  // first use of K tile j2 (its second half, step s2 + 1, comes one iteration later)
              mbar_wait(&k_full[ks2], (j2 / kKS) & 1);
              tc_fence_after();
            }
            issue_qk(q, ks2, sub);  // into the buffer PV(q, s) has just read
            if (q == 1 && leader && (sub == 1 || s2 == n_steps - 1)) umma_commit(&k_empty[ks2]);
          }
          if (tracing && leader) p.trace[(s * 2 + q) * kAttnTraceSlots + 7] = clock64();
          __syncwarp();
        }
      }
    }
  } else {
    setmaxnreg_inc<Cfg::kRegsLarge>();
    {
      // ===================== softmax warpgroups: one thread per query row, 64 keys per step =====================
      const int q = (warp - 4) >> 2;
      const int quad = warp & 3;
      const int row_in_tile = quad * 32 + lane;
      const int pos = q0 + q * 128 + row_in_tile;
      const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
      const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
      const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128);
      const float c = p.scale_log2;
      const f32x2 c2 = pack2(c, c);
      float m = -INFINITY, l = 0.f;
      const bool tr = tracing && quad == 0 && lane == 0;
      for (int s = 0; s < n_steps; ++s) {
        const int sub = s & 1;
        const int valid = p.N - s * 64;  // >= 64 on every step but possibly the last
        mbar_wait(&s_full[2 * q + sub], (s >> 1) & 1);
        tc_fence_after();
        if (tr) p.trace[(s * 2 + q) * kAttnTraceSlots + 0] = clock64();
        uint32_t sr[2][32];
        tmem_ld32(t_s + sub * 64, sr[0]);
        tmem_ld32(t_s + sub * 64 + 32, sr[1]);
        tmem_ld_wait();
        if (tr) p.trace[(s * 2 + q) * kAttnTraceSlots + 1] = clock64();
        if (valid < 64) {  // ragged last step: keys past N score -inf -> probability 0
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cc * 32 + i >= valid) sr[cc][i] = 0xff800000u;
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[0][i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[1][i]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[1][i + 1]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const bool need = (mx - m) * c > kAttnRescaleThreshold;  // true on the first step (m = -inf)
        const float m_new = need ? mx : m;
        const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
        if (s > 0 && __any_sync(0xffffffffu, need)) {
          // O_q holds PV(0..s-1); unlike schedule 3, s_full(s) only implies PV(s-2): wait for PV(s-1) itself
          mbar_wait(&pv_done[q], (s - 1) & 1);
          tc_fence_after();
          attn_rescale_o<kHeadDim>(t_o, alpha);
        }
        if (tr) p.trace[(s * 2 + q) * kAttnTraceSlots + 2] = clock64();
        const float mc = m_new * c;
        const f32x2 nmc2 = pack2(-mc, -mc);
        f32x2 sum2 = pack2(0.f, 0.f);
        uint32_t pk[32];
        attn_exp_half<kEmu>(sr[0], sr[1], c2, nmc2, sum2, pk);
        tmem_st32(t_s + sub * 64, pk);  // P (bf16 pairs) over the first 32 columns of the buffer just read
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * q + sub]);
        if (tr) p.trace[(s * 2 + q) * kAttnTraceSlots + 3] = clock64();
        float sum0, sum1;
        unpack2(sum2, sum0, sum1);
        l = l * alpha + (sum0 + sum1);
        m = m_new;
      }
      // ---- finalize: O / l -> bf16, token-major store.  pv_done completes once per step and the last s_full seen only implies
      // PV(n_steps - 3): walk the two phases that may still be open, in order (a parity wait must not lag two phases behind)
      if (n_steps >= 2) mbar_wait(&pv_done[q], (n_steps - 2) & 1);
      mbar_wait(&pv_done[q], (n_steps - 1) & 1);
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
