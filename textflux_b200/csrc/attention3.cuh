// Joint [text;image] flash attention, schedule 3: a ping-pong schedule (two 128-row query tiles per CTA, S/P/O in TMEM)
// built around what the clock traces of its predecessor showed (profiles/r1j_attn_trace.md):
//   * the issuing thread was the slow link: `warp == 1 && lane == 0` is a divergent branch, so every tcgen05.mma was
//     wrapped in a value-broadcast loop (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~15 dependent SASS instructions, about as
//     long as the 64 cycles the MMA itself runs).  Here the whole warp walks the loop in a warp-uniform context, the
//     operand descriptors are precomputed 64-bit bases bumped by constants, and one elected lane issues.
//   * K and V rings are separate (K(j+2) may land as soon as QK(., j) retired, one iteration earlier than with a joint ring).
//   * P is handed to the tensor pipe in two halves: PV over keys 0..63 starts while the exponentials of keys 64..127
//     are still being evaluated.
//   * 12 warps with setmaxnreg: the producer/issuer warpgroup drops to 72 registers, the two softmax warpgroups take
//     216 each, which holds the whole score row + packed P without spills.
//   * no wait on pv_done inside the loop: s_full(j) already implies PV(j-1) retired (commit tracks all earlier MMAs).
// Arithmetic: fp32 scores, lazily rescaled running max (threshold 2^8), a fraction of the
// exponentials on the FMA pipe (packed Cody-Waite + cubic), P rounded to bf16 for the PV MMA, fp32 O.
#pragma once
#include <cuda.h>

#include "attention_common.cuh"
#include "ptx.cuh"

namespace tfx {

template <int kHeadDim>
struct Attn3Cfg {
  static constexpr int kHalves = kHeadDim / 64;
  static constexpr int kTileBytes = 128 * kHeadDim * 2;  // one 128-row Q/K/V tile
  static constexpr int kKStages = 2;
  static constexpr int kVStages = 2;
  static constexpr int kThreads = 384;                   // wg0: TMA, MMA, TMEM alloc, spare; wg1/wg2: softmax of q-tile 0/1
  static constexpr int kXchBytes = 2 * 2 * 128 * 4;      // scratch behind the barriers (attention4: merge flag)
  static constexpr int kSmemBytes = (2 + kKStages + kVStages) * kTileBytes + 1024 + 256 + kXchBytes;
  static constexpr int kSCol = 0;    // S_q at q*128 (P_q aliases its first 64 columns)
  static constexpr int kOCol = 256;  // O_q at 256 + q*128
  static constexpr int kRegsSmall = 72, kRegsLarge = 216;  // 80 / 216 would use the register file exactly and never gets its setmaxnreg.inc granted (measured: hangs)
};

template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// Trace slots (kTrace builds only; CTA (0,0,0)): per KV tile j and query tile q, 8 clock64 stamps:
//   0 s_full seen by the softmax warp   1 score row in registers   2 row max known   3 first half of P handed over
//   4 second half handed over          5 issuer saw half 0         6 PV issued       7 next QK issued
constexpr int kAttnTraceSlots = 8;

// named barrier of the 256 softmax threads (warps 4..11)
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int kHeadDim, int kEmu, bool kSplitP, bool kTrace>
__global__ void __launch_bounds__(Attn3Cfg<kHeadDim>::kThreads, 1)
attention3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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
  uint64_t* s_full = v_empty + kVS;      // [2]
  uint64_t* p_full = s_full + 2;         // [2 q][2 halves]
  uint64_t* pv_done = p_full + 4;        // [2]
  uint64_t* h0_done = pv_done + 2;       // [2]  kAttnNoMax: the PV MMAs over keys 0..63 of the tile have retired
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(h0_done + 2);
  constexpr bool kTwoBars = kSplitP;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.H + head;
  const int n_kv = (p.N + 127) / 128;
  const bool tracing = kTrace && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  long long* ctr = nullptr;
  if (kTrace && p.cta_trace != nullptr) {
    ctr = p.cta_trace + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
    if (threadIdx.x == 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      ctr[0] = (long long)globaltimer_ns();
      ctr[7] = smid;
    }
  }
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
  if (kTrace && ctr && threadIdx.x == 0) ctr[1] = (long long)globaltimer_ns();

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
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kHeadDim, 0, 1);  // B = V is MN-major (dh contiguous)
      const bool leader = elect_one();
      // descriptor bases; the start-address field (bits 0..13, 16-byte units) is bumped by compile-time constants
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
            // 16 kv rows (= 2 KiB) per step; dh halves are kHalfBytes apart (leading-dimension byte offset)
            umma_ts(d, a + uint32_t(kk * 8), bb + uint64_t(kk * 128), idesc_pv, !(first_tile && kk == 0));
          }
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (kTrace && ctr && leader) ctr[2] = (long long)globaltimer_ns();
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
          if (tracing && leader) p.trace[(j * 2 + q) * kAttnTraceSlots + 5] = clock64();
          if (kTwoBars) {
            issue_pv(q, vs, 0, 4, j == 0);
            if (kAttnNoMax && leader) umma_commit(&h0_done[q]);
            mbar_wait(&p_full[2 * q + 1], j & 1);
            tc_fence_after();
            issue_pv(q, vs, 4, 8, false);
          } else {
            issue_pv(q, vs, 0, 8, j == 0);
          }
          if (leader) {
            umma_commit(&pv_done[q]);
            if (q == 1) umma_commit(&v_empty[vs]);
          }
          if (tracing && leader) p.trace[(j * 2 + q) * kAttnTraceSlots + 6] = clock64();
          if (more) {
            if (q == 0) {
              mbar_wait(&k_full[ksn], ((j + 1) / kKS) & 1);
              tc_fence_after();
            }
            issue_qk(q, ksn);
            if (q == 1 && leader) umma_commit(&k_empty[ksn]);
          }
          if (tracing && leader) p.trace[(j * 2 + q) * kAttnTraceSlots + 7] = clock64();
          __syncwarp();
        }
      }
    }
  } else {
    setmaxnreg_inc<Cfg::kRegsLarge>();
    {
      // ===================== softmax warpgroups: one thread per query row =====================
      const int q = (warp - 4) >> 2;
      const int quad = warp & 3;
      const int row_in_tile = quad * 32 + lane;
      const int pos = q0 + q * 128 + row_in_tile;
      const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
      const uint32_t t_s = t_lane + uint32_t(Cfg::kSCol + q * 128);
      const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol + q * 128);
      const float c = p.scale_log2;
      const bool tr = tracing && quad == 0 && lane == 0;
      float m = -INFINITY, l = 0.f, pend = -INFINITY;
      for (int j = 0; j < n_kv; ++j) {
        const int valid = p.N - j * 128;  // >= 128 on every tile but possibly the last
        mbar_wait(&s_full[q], j & 1);
        tc_fence_after();
        if (tr) p.trace[(j * 2 + q) * kAttnTraceSlots + 0] = clock64();
        if (kTrace && ctr && j == 0 && warp == 4 && lane == 0) ctr[3] = (long long)globaltimer_ns();
        auto handover = [&](int half, const uint32_t (&pk)[32]) {
          tmem_st32(t_s + half * 32, pk);  // P (bf16 pairs) over the S columns already in registers
          if (kSplitP || half == 1) {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(kSplitP ? &p_full[2 * q + half] : &p_full[2 * q]);
          }
        };
        auto stamp = [&](int slot) {
          if (tr) p.trace[(j * 2 + q) * kAttnTraceSlots + slot] = clock64();
        };
        auto wait_h0 = [&]() {
          mbar_wait(&h0_done[q], j & 1);
          tc_fence_after();
        };
        attn_softmax_tile<kHeadDim, kEmu, kSplitP>(t_s, t_o, c, valid, j == 0, m, l, pend, handover, stamp, wait_h0);
      }
      // ---- finalize: O / l -> bf16, token-major store
      if (kTrace && ctr && warp == 4 && lane == 0) ctr[4] = (long long)globaltimer_ns();
      mbar_wait(&pv_done[q], (n_kv - 1) & 1);
      tc_fence_after();
      if (kTrace && ctr && warp == 4 && lane == 0) ctr[5] = (long long)globaltimer_ns();
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
      if (kTrace && ctr && warp == 4 && lane == 0) ctr[6] = (long long)globaltimer_ns();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
