// Joint [text;image] flash attention, schedule 10 (experiment): schedule 9 (one 128-row query tile per CTA, three score buffers in TMEM)
// with THREE softmax warpgroups -- 16 warps per CTA, warpgroup w owns score buffer w and takes KV tiles j = w mod 3.
//
// Why: ncu of schedule 9 (profiles/r2ae_attn_sbuf3.md) shows the same ~69 % tensor-pipe utilisation as schedules 3 / 5 although
// the S-ready -> P-ready -> PV -> QK chain is gone: what bounds all of them is the THROUGHPUT of two softmax warps per scheduler
// (one 128 x 128 tile per ~2900 cycles per warp: dependent MUFU / FFMA2 / F2FP chains of an in-order warp, XU 52 %, issue 45-49 %).
// A third warp per scheduler adds the latency hiding; the price is 144 registers per softmax thread, so the score row is read from
// TMEM twice (maximum, then exponentials 64 columns at a time) instead of sitting in registers whole.
#pragma once
#include <cuda.h>

#include "../../textflux_b200/csrc/attention3.cuh"

namespace tfx {

template <int kHeadDim>
struct Attn10Cfg {
  static constexpr int kTileBytes = 128 * kHeadDim * 2;
  static constexpr int kKStages = 3;  // = score buffers: K(j + 3) replaces K(j) once QK(j) has run
  static constexpr int kVStages = 2;
  static constexpr int kSBufs = 3;
  static constexpr int kWG = 3;          // softmax warpgroups = score buffers
  static constexpr int kThreads = 128 + 128 * kWG;  // wg0: K TMA, MMA, TMEM alloc, V TMA; wg 1 + w: KV tiles j = w mod 3
  static constexpr int kXchBytes = 2 * kWG * 128 * 4;  // [kWG] reference maxima + [kWG] partial sums per row
  static constexpr int kSmemBytes = (1 + kKStages + kVStages) * kTileBytes + 1024 + 256 + kXchBytes;
  static constexpr int kOCol = 384;  // S_b at b * 128 (P_b aliases its first 64 columns), O behind them
  static constexpr int kRegsSmall = 56, kRegsLarge = 144;  // 128 * 56 + 384 * 144 = 62464 of the 65536 registers (launch: 512 * 128)
};

// producer / consumer named barriers between warps that own the same 32 rows (ids 1..12: 1 + quad * 3 + publishing warpgroup), id 13:
// all softmax threads, for the final exchange
#ifndef TFX_ATTN9
__device__ __forceinline__ void pair9_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair9_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
#endif

template <int kHeadDim, int kEmu>
__global__ void __launch_bounds__(Attn10Cfg<kHeadDim>::kThreads, 1)
attention10_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  using Cfg = Attn10Cfg<kHeadDim>;
  constexpr int kHalves = kHeadDim / 64;
  constexpr int kHalfBytes = 128 * 128;  // 128 rows x 128 B
  constexpr int kKS = Cfg::kKStages, kVS = Cfg::kVStages, kSB = Cfg::kSBufs, kWG = Cfg::kWG;
  static_assert(kWG == kSB, "warpgroup w owns score buffer w");
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
        mbar_wait(&k_empty[ks], ((j / kKS) & 1) ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&k_full[ks], Cfg::kTileBytes);
          for (int h = 0; h < kHalves; ++h)
            tma_load_3d(&tmK, &k_full[ks], sK + ks * Cfg::kTileBytes + h * kHalfBytes, h * 64, j * 128, bh, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 3) {
      // ===================== V producer: its own warp, so that a V slot waiting for PV(j - 2) never holds back K(j + 1) =====================
      const bool leader = elect_one();
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % kVS;
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
  } else {
    setmaxnreg_inc<Cfg::kRegsLarge>();
    // ===================== softmax: warpgroup w takes KV tiles j = w, w + kWG, ... (score buffer w); one thread per query row ========
    // Two passes over the scores in TMEM (maximum, then exponentials 64 columns at a time): the row never sits in registers whole,
    // so three warpgroups fit the register file at 144 registers per thread.
    const int w = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    const uint32_t t_o = t_lane + uint32_t(Cfg::kOCol);
    const uint32_t t_s = t_lane + uint32_t(w * 128);
    const int w_prev = (w + kWG - 1) % kWG;
    float* xm = xch + row_in_tile;        // [kWG][128] reference maxima
    float* xl = xch + kWG * 128 + row_in_tile;  // [kWG][128] partial sums
    const int bar_pub = 1 + quad * kWG + w, bar_sub = 1 + quad * kWG + w_prev;
    const float c = p.scale_log2;
    const f32x2 c2 = pack2(c, c);
    float m = -INFINITY, l = 0.f;  // m: the reference the row's probabilities (every warpgroup's) and O are expressed against
    for (int j = w; j < n_kv; j += kWG) {
      const int valid = p.N - j * 128;
      mbar_wait(&s_full[w], (j / kSB) & 1);
      tc_fence_after();
      float mx;
      {
        uint32_t s0[32], s1[32];
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld32(t_s + hh * 64, s0);
          tmem_ld32(t_s + hh * 64 + 32, s1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (hh * 64 + i < valid) mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
            if (hh * 64 + 32 + i < valid) mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
          }
        }
        mx = fmaxf(mx0, mx1);
      }
      if (j > 0) {  // the reference after tile j - 1 (the previous warpgroup's): it has already rescaled O if it moved it
        pair9_sync(bar_sub);
        const float mp = xm[w_prev * 128];
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
      if (j + 1 < n_kv) {  // publish before the exponentials: the next warpgroup's tile j + 1 only waits for this
        xm[w * 128] = m;
        __threadfence_block();
        pair9_arrive(bar_pub);
      }
      const float mc = m * c;
      const f32x2 nmc2 = pack2(-mc, -mc);
      f32x2 sum2 = pack2(0.f, 0.f);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t s0[32], s1[32], pk[32];
        tmem_ld32(t_s + hh * 64, s0);
        tmem_ld32(t_s + hh * 64 + 32, s1);
        tmem_ld_wait();
        if (valid < 128) {  // ragged last tile: keys past N score -inf -> probability 0
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (hh * 64 + i >= valid) s0[i] = 0xff800000u;
            if (hh * 64 + 32 + i >= valid) s1[i] = 0xff800000u;
          }
        }
        attn_exp_half<kEmu>(s0, s1, c2, nmc2, sum2, pk);
        tmem_st32(t_s + hh * 32, pk);  // P half hh over score columns this thread has consumed (hh = 1: columns 32..63, read in pass hh = 0)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * w + hh]);
      }
      float sum0, sum1;
      unpack2(sum2, sum0, sum1);
      l += sum0 + sum1;
    }
    // ---- finalize: the warpgroup of the last tile holds the final reference; the partial sums meet; the columns are dealt out in 32s
    const int last = n_kv - 1;
    const int w_last = last % kWG;
    xm[w * 128] = m;
    xl[w * 128] = l;
    if (w == w_last) {
      mbar_wait(&pv_done[last % kSB], (last / kSB) & 1);
      tc_fence_after();
      tc_fence_before();
    }
    __threadfence_block();
    asm volatile("bar.sync 13, %0;" ::"n"(128 * kWG) : "memory");
    tc_fence_after();
    const float m_fin = xm[w_last * 128];
    float lsum = 0.f;
#pragma unroll
    for (int ww = 0; ww < kWG; ++ww) lsum += xl[ww * 128] * ex2((xm[ww * 128] - m_fin) * c);
    const float inv_l = 1.0f / lsum;
    const int pos = q0 + row_in_tile;
    const bool row_ok = pos < p.N;
    const long long row = (pos < p.T) ? (long long)b * p.T + pos : (long long)p.B * p.T + (long long)b * p.S + (pos - p.T);
    __nv_bfloat16* dst = p.out + row * p.ld_out + head * kHeadDim;
#pragma unroll 1
    for (int cch = w; cch < kHeadDim / 32; cch += kWG) {
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

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace tfx
