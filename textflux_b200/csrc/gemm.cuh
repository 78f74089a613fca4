// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   Y[m, n] = sum_k A[m, k] * W[n, k]  (+ bias[n])  -> epilogue
//
// A (activations, [M,K] row-major bf16) and W (nn.Linear weight, [N,K] row-major bf16) are both K-major, which is
// exactly the UMMA "TN" case: TMA drops 128B-swizzled [rows x 64] bf16 boxes into shared memory, one elected thread
// issues tcgen05.mma (kind::f16, fp32 accumulate) into TMEM, four epilogue warps read the accumulator back with
// tcgen05.ld and apply the fused epilogue.  Replaces every `nn.Linear` call site of the reference hot path
// (transformer_flux.py:694-696,894-895,920; attention_processor.py:237-260; attention.py:1218-1232)
// and, through the epilogues, K4-K8 and K10 (RMSNorm, RoPE, GELU, gate+residual, cat, Euler step).
//
// kCtaGroup = 1: one CTA owns a 128 x 256 tile.   kCtaGroup = 2: a CTA pair (cluster of 2) owns 256 x 256, the leader
// CTA issues cta_group::2 MMAs, each CTA loads its 128 rows of A and its 128 rows of W (halved smem traffic per SM).
//
// Up to two row groups per launch (text rows / image rows of the double-stream blocks: same N,K, different
// weights, bias, modulation vectors) so both streams share one persistent grid.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace tfx {

enum EpiMode : int {
  EPI_STORE = 0,     // out = bf16(acc + bias)
  EPI_MUL = 6,         // out = bf16(res * bf16(acc + bias)): the gated half of T5's DenseGatedActDense (hidden_gelu * hidden_linear)
  EPI_QUICK_GELU = 7,  // out = bf16(x * sigmoid(1.702 x)), x = bf16(acc + bias): CLIP's MLP activation
  EPI_STORE_F32 = 5,  // out (float*, ldo in floats) = acc, no bias: attention scores of the VAE mid block, kept in fp32 for the softmax
  EPI_GELU = 1,      // out = bf16(gelu_tanh(bf16(acc + bias)))
  EPI_GATE_RES = 2,  // out = bf16(res + bf16(gate * bf16(acc + bias)))
  EPI_QKV = 3,       // per-head RMSNorm + RoPE on q,k; scatter q,k,v head-major into the joint [text;image] buffers
  EPI_EULER = 4,     // v = bf16(acc + bias); latents' = bf16(latents + bf16(dt * v))   (scheduler step fused)
};

struct GemmGroup {
  int M;                // rows of this group
  int rows_per_sample;  // rows of one batch sample inside the group
  int pos_offset;       // joint-sequence position of a sample's first row (0 for text, T for image)
  const __nv_bfloat16* bias;  // [N]
  __nv_bfloat16* out;         // STORE/GELU/GATE_RES/EULER(noise_pred, may be null) destination (first row of group)
  long long ldo;
  const __nv_bfloat16* res;   // GATE_RES residual / EULER latents in
  long long ldr;
  const __nv_bfloat16* gate;  // GATE_RES: gate vector of sample b at gate + b * gate_stride
  long long gate_stride;
  const __nv_bfloat16* rms_q;  // QKV: RMSNorm weights [head_dim]
  const __nv_bfloat16* rms_k;
  __nv_bfloat16* out2;  // EULER: latents out
};

// Implicit-GEMM 3x3 convolution over an NHWC bf16 image (AutoencoderKL: models/resnet.py ResnetBlock2D.conv1/conv2,
// models/downsampling.py Downsample2D, models/upsampling.py Upsample2D.conv).  The A operand is not a matrix: an M tile is a 16 x 8
// patch of OUTPUT pixels, a k-block is (tap, 64 input channels), and its A tile is one TMA box of the image shifted by the tap --
// out-of-bounds pixels read as zero, which is the padding.  W is [Cout, 9 * Cin] with k = tap * Cin + c.
struct ConvGeom {
  int mode;        // 0 plain GEMM | 1: stride 1, pad 1, map [C, W, H, B] | 2: stride 2, pad (0,1,0,1), map [C, 2, W_in/2, 2, H_in/2] of ONE image
  int H, W;        // output height / width
  int tiles_x, tiles_y, n_patches;  // 16-wide x 8-high patches per image row / column; B * tiles_y * tiles_x
  int cin_blocks;  // Cin / 64
};
constexpr int kConvPatchW = 16, kConvPatchH = 8;

struct GemmParams {
  int N, K;
  int num_groups;
  GemmGroup g[2];
  int n_split;  // columns [0,n_split) run mode0, [n_split,N) run mode1 (n_split multiple of 256, or == N)
  int mode0, mode1;
  int col_offset1;  // mode1 output column = n - n_split + col_offset1
  // QKV scatter
  int D, head_dim, num_heads, n_joint;
  __nv_bfloat16 *q, *k, *v;  // [B, H, n_joint, head_dim]
  const float2* rope;        // [n_joint, head_dim/2] (cos, sin)
  float rms_eps;
  const float* dt_ptr;  // EULER: device scalar float(bf16(sigma_next - sigma))
  int debug_flags;      // bit 0: A loads with L2 evict_last, bit 1: B (weight) loads with L2 evict_first
  ConvGeom conv;        // conv.mode != 0: g[0].M = B * H * W output pixels, K = 9 * Cin, tmA0 is the image map (see ConvGeom)
  int m_band;           // tile order (b < 0: N bands of -b tiles, see gemm_tile_coords): 0 = M-fastest over all M tiles (a wave spans every M tile and a few N tiles: each weight tile
                        // is fetched once, A must stay in L2); b > 0 = bands of b M tiles, inside a band M-fastest over all N tiles
                        // (a wave spans b M tiles x all N tiles: for wide-K GEMMs whose A is larger than the L2)
  int k_snake;          // 1: the tiles of every second band (m_band != 0) walk their k-blocks back to front, so the W slices the previous
                        // band of tiles read last are the first ones the next band asks the L2 for (wide-K GEMMs whose operands outgrow the L2)
  int k_ext;            // 0 | 64: one extra k-block behind the K of A whose operands come from the extension descriptors
                        // (A side tmE0 / tmE1: [M, 64]; W side tmF0 / tmF1: [N, 64]).  Unfused LoRA rides here: with
                        // T = bf16(x A_lora^T) as the A extension and (alpha/r) B_lora as the W extension the accumulator holds
                        // x W^T + (alpha/r) T B^T before any epilogue runs -- PEFT's `base(x) + scaling * lora_B(lora_A(x))`
                        // (run_inference_lora.py:52-65 keeps the adapters unfused) without touching the base weight
};

constexpr int kGemmBlockN = 256;  // default tile width; 224 / 192 are instantiated to cut wave quantisation
constexpr int kGemmBlockK = 64;
constexpr int kGemmThreads = 384;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue (2 per TMEM lane quadrant)
constexpr int kGemmEpiWarps = 8;

template <int kCtaGroup, int kBN = kGemmBlockN>
struct GemmCfg {
  static_assert(kBN % 32 == 0 && kBN <= 256, "tile width must be a multiple of 32 (epilogue chunk) and fit one UMMA");
  static constexpr int kTileM = 128 * kCtaGroup;
  static constexpr int kBRows = kBN / kCtaGroup;  // rows of W each CTA loads
  static constexpr int kABytes = 128 * kGemmBlockK * 2;
  static constexpr int kBBytes = kBRows * kGemmBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kCtaGroup == 1) ? 4 : 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&x)[32]) { store_row_chunk_bf16x32(dst, x); }
// kReadOnly: data never written during the kernel (bias, gates, norm weights) -> ld.global.nc
template <bool kReadOnly = true>
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float (&x)[32]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u = kReadOnly ? __ldg(s4 + i) : s4[i];
    x[8 * i + 0] = bf16_lo(u.x); x[8 * i + 1] = bf16_hi(u.x);
    x[8 * i + 2] = bf16_lo(u.y); x[8 * i + 3] = bf16_hi(u.y);
    x[8 * i + 4] = bf16_lo(u.z); x[8 * i + 5] = bf16_hi(u.z);
    x[8 * i + 6] = bf16_lo(u.w); x[8 * i + 7] = bf16_hi(u.w);
  }
}

// acc chunk (32 fp32 columns of this thread's row) + bias -> bf16-rounded linear output, as nn.Linear returns it
__device__ __forceinline__ void linear_out(const uint32_t (&v)[32], const __nv_bfloat16* bias, int n, int N, float (&x)[32]) {
  float b[32];
  if (n + 32 <= N) {
    load_bf16x32(bias + n, b);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = (n + i < N) ? __bfloat162float(bias[n + i]) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = bf16_round(__uint_as_float(v[i]) + b[i]);
}

// One thread's share of a tile's epilogue: row m_local of group grp, the column chunks of half `half`.
// t_acc = TMEM address of the accumulator stage at this warp's lane quadrant.
template <int kBN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, int grp, int m_local, int n_tile0, uint32_t t_acc,
                                                   int half, int lane) {
  constexpr int kChunks = kBN / 32;
  const GemmGroup& G = p.g[grp];
  const bool row_ok = m_local < G.M;
  const int mode = (n_tile0 < p.n_split) ? p.mode0 : p.mode1;
  const int bidx = m_local / G.rows_per_sample;
  const int pos = G.pos_offset + (m_local - bidx * G.rows_per_sample);
  if (mode == EPI_QKV) {
    const int sect = n_tile0 / p.D;  // 0 q, 1 k, 2 v
    const int dh = p.head_dim;
    const int head0 = (n_tile0 - sect * p.D) / dh;
    const int heads_in_tile = kBN / dh;  // QKV launches use kBN = 256 (host-enforced): 2 or 4 heads per tile
    const int chunks_per_head = dh / 32;
    __nv_bfloat16* dst_base = (sect == 0) ? p.q : (sect == 1) ? p.k : p.v;
    const __nv_bfloat16* rmsw = (sect == 0) ? G.rms_q : G.rms_k;
    for (int hh = half * (heads_in_tile / 2); hh < (half + 1) * (heads_in_tile / 2); ++hh) {
      const int head = head0 + hh;
      if (n_tile0 + hh * dh >= p.N) break;
      __nv_bfloat16* dst = dst_base + ((long long)(bidx * p.num_heads + head) * p.n_joint + pos) * dh;
      float rstd = 0.f;
      if (sect < 2) {
        float ss = 0.f;
        for (int c = 0; c < chunks_per_head; ++c) {
          uint32_t v[32];
          float x[32];
          tmem_ld32(t_acc + uint32_t(hh * dh + c * 32), v);
          tmem_ld_wait();
          linear_out(v, G.bias, n_tile0 + hh * dh + c * 32, p.N, x);
#pragma unroll
          for (int i = 0; i < 32; ++i) ss = fmaf(x[i], x[i], ss);
        }
        rstd = rsqrtf(ss / float(dh) + p.rms_eps);
      }
      for (int c = 0; c < chunks_per_head; ++c) {
        uint32_t v[32];
        float x[32];
        tmem_ld32(t_acc + uint32_t(hh * dh + c * 32), v);
        tmem_ld_wait();
        linear_out(v, G.bias, n_tile0 + hh * dh + c * 32, p.N, x);
        if (sect < 2) {
          float w[32];
          load_bf16x32(rmsw + c * 32, w);
          // RMSNorm.forward: (x * rsqrt(var+eps)) -> bf16 -> * weight(bf16) -> bf16   (normalization.py:535-546)
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = bf16_round(bf16_round(x[i] * rstd) * w[i]);
          // apply_rotary_emb: fp32 x*cos + rot(x)*sin on interleaved pairs (embeddings.py:904-914)
          if (row_ok) {
            const float4* cs4 = reinterpret_cast<const float4*>(p.rope + (long long)pos * (dh / 2) + c * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 cs = __ldg(cs4 + i);  // (cos0, sin0, cos1, sin1)
              float a0 = x[4 * i + 0], a1 = x[4 * i + 1], b0 = x[4 * i + 2], b1 = x[4 * i + 3];
              x[4 * i + 0] = __fadd_rn(__fmul_rn(a0, cs.x), __fmul_rn(-a1, cs.y));
              x[4 * i + 1] = __fadd_rn(__fmul_rn(a1, cs.x), __fmul_rn(a0, cs.y));
              x[4 * i + 2] = __fadd_rn(__fmul_rn(b0, cs.z), __fmul_rn(-b1, cs.w));
              x[4 * i + 3] = __fadd_rn(__fmul_rn(b1, cs.z), __fmul_rn(b0, cs.w));
            }
          }
        }
        if (row_ok) store_bf16x32(dst + c * 32, x);
      }
    }
  } else {
    const int n_out0 = (n_tile0 < p.n_split) ? n_tile0 : (n_tile0 - p.n_split + p.col_offset1);
    for (int c = half * ((kChunks + 1) / 2); c < (half ? kChunks : (kChunks + 1) / 2); ++c) {
      const int n = n_tile0 + c * 32;
      if (n >= p.N) break;  // warp-uniform
      uint32_t v[32];
      float x[32];
      tmem_ld32(t_acc + uint32_t(c * 32), v);
      tmem_ld_wait();
      linear_out(v, G.bias, n, p.N, x);
      const int no = n_out0 + c * 32;
      if (mode == EPI_STORE_F32) {
        if (row_ok) {
          float* o = reinterpret_cast<float*>(G.out) + (long long)m_local * G.ldo + no;
          if (n + 32 <= p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(o)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else {
            for (int i = 0; i < 32 && n + i < p.N; ++i) o[i] = __uint_as_float(v[i]);
          }
        }
        continue;
      }
      if (mode == EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = gelu_tanh(x[i]);
      } else if (mode == EPI_QUICK_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = x[i] / (1.0f + __expf(-1.702f * x[i]));
      } else if (mode == EPI_MUL) {
        if (row_ok) {
          if (n + 32 <= p.N) {
            float r[32];
            load_bf16x32<false>(G.res + (long long)m_local * G.ldr + no, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] *= r[i];
          } else {
            for (int i = 0; i < 32 && n + i < p.N; ++i) x[i] *= __bfloat162float(G.res[(long long)m_local * G.ldr + no + i]);
          }
        }
      } else if (mode == EPI_GATE_RES) {
        if (row_ok) {
          float g[32], r[32];
          load_bf16x32(G.gate + (long long)bidx * G.gate_stride + n, g);
          load_bf16x32<false>(G.res + (long long)m_local * G.ldr + no, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = r[i] + bf16_round(g[i] * x[i]);
        }
      } else if (mode == EPI_EULER) {
        if (row_ok) {
          const float dt = __ldg(p.dt_ptr);
          if (n + 32 <= p.N) {
            float r[32], y[32];
            if (G.out) store_bf16x32(G.out + (long long)m_local * G.ldo + no, x);
            load_bf16x32<false>(G.res + (long long)m_local * G.ldr + no, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = r[i] + bf16_round(dt * x[i]);
            store_bf16x32(G.out2 + (long long)m_local * G.ldr + no, y);
          } else {
            for (int i = 0; i < 32 && n + i < p.N; ++i) {
              if (G.out) G.out[(long long)m_local * G.ldo + no + i] = __float2bfloat16_rn(x[i]);
              float r = __bfloat162float(G.res[(long long)m_local * G.ldr + no + i]);
              G.out2[(long long)m_local * G.ldr + no + i] = __float2bfloat16_rn(r + bf16_round(dt * x[i]));
            }
          }
        }
        continue;
      }
      if (row_ok) {
        if (n + 32 <= p.N) {
          store_bf16x32(G.out + (long long)m_local * G.ldo + no, x);
        } else {
          for (int i = 0; i < 32 && n + i < p.N; ++i) G.out[(long long)m_local * G.ldo + no + i] = __float2bfloat16_rn(x[i]);
        }
      }
    }
  }
}

// tile index -> (M tile, N tile) under GemmParams::m_band
__device__ __forceinline__ void gemm_tile_coords(int t, int MT, int NT, int band, int& mi, int& ni) {
  if (band < 0) {  // N bands of -band tiles: inside a band N-fastest over all M tiles (the band of W stays in L2, A streams once per band)
    const int bw = -band, per = MT * bw;
    const int nb = t / per, r = t - nb * per;
    const int cols = min(bw, NT - nb * bw);
    mi = r / cols;
    ni = nb * bw + (r - mi * cols);
    return;
  }
  if (band == 0 || band >= MT) { mi = t % MT; ni = t / MT; return; }
  const int per = band * NT;
  const int b = t / per, r = t - b * per;
  const int rows = min(band, MT - b * band);
  ni = r / rows;
  mi = b * band + (r - ni * rows);
}

template <int kCtaGroup, int kBN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmE0, const __grid_constant__ CUtensorMap tmE1,
                    const __grid_constant__ CUtensorMap tmF0, const __grid_constant__ CUtensorMap tmF1,
                    const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<kCtaGroup, kBN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0;
  const bool is_leader = cta_rank == 0;
  pdl_launch_dependents();  // let the next kernel's CTAs set up (barriers, TMEM) while this grid drains

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA0);
    prefetch_tensormap(&tmB0);
    if (p.num_groups > 1) {
      prefetch_tensormap(&tmA1);
      prefetch_tensormap(&tmB1);
    }
    if (p.k_ext) {
      prefetch_tensormap(&tmE0);
      prefetch_tensormap(&tmF0);
      if (p.num_groups > 1) {
        prefetch_tensormap(&tmE1);
        prefetch_tensormap(&tmF1);
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kGemmEpiWarps * kCtaGroup);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<kCtaGroup>(tmem_base_ptr, 512);
    tmem_relinquish<kCtaGroup>();
  }
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  // ---- tile schedule (identical in every role)
  const int mt0 = p.conv.mode ? (p.conv.n_patches + kCtaGroup - 1) / kCtaGroup : (p.g[0].M + Cfg::kTileM - 1) / Cfg::kTileM;
  const int mt1 = (p.num_groups > 1) ? (p.g[1].M + Cfg::kTileM - 1) / Cfg::kTileM : 0;
  const int MT = mt0 + mt1;
  const int NT = (p.N + kBN - 1) / kBN;
  const int num_tiles = MT * NT;
  const int KBm = (p.K + kGemmBlockK - 1) / kGemmBlockK;  // k-blocks whose A operand is the main descriptor (a ragged last one is zero-filled by TMA)
  const int KB = KBm + p.k_ext / kGemmBlockK;        // + the extension block (0 or 1)
  const int first_tile = blockIdx.x / kCtaGroup;
  const int tile_step = gridDim.x / kCtaGroup;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    // activations are re-read by every N tile, each weight tile only by the M tiles running at the same time
    const uint64_t hint_a = (p.debug_flags & 1) ? kEvictLast : kEvictNormal;
    const uint64_t hint_b = (p.debug_flags & 2) ? kEvictFirst : kEvictNormal;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      int mi, ni;
      gemm_tile_coords(t, MT, NT, p.m_band, mi, ni);
      // direction from the tile's band, not from which CTA computes it: a given output element sums in the same order whatever the grid
      const bool backwards = p.k_snake && (p.m_band > 0 ? ((mi / p.m_band) & 1) : p.m_band < 0 ? ((ni / -p.m_band) & 1) : 0);
      const int grp = (mi < mt0) ? 0 : 1;
      const int m0 = (grp ? mi - mt0 : mi) * Cfg::kTileM + int(cta_rank) * 128;
      const int n0 = ni * kBN + int(cta_rank) * Cfg::kBRows;
      const CUtensorMap* tAm = grp ? &tmA1 : &tmA0;
      const CUtensorMap* tAe = grp ? &tmE1 : &tmE0;
      const CUtensorMap* tBm = grp ? &tmB1 : &tmB0;
      const CUtensorMap* tBe = grp ? &tmF1 : &tmF0;
      int cimg = 0, cy0 = 0, cx0 = 0;  // convolution: this CTA's patch of output pixels
      if (p.conv.mode) {
        const int patch = mi * kCtaGroup + int(cta_rank);
        const int per = p.conv.tiles_x * p.conv.tiles_y;
        cimg = patch / per;
        const int r = patch - cimg * per;
        cy0 = (r / p.conv.tiles_x) * kConvPatchH;
        cx0 = (r % p.conv.tiles_x) * kConvPatchW;
        if (patch >= p.conv.n_patches) { cimg = 0; cy0 = p.conv.H + 64; }  // phantom half of the last pair: every pixel out of bounds
      }
      for (int step = 0; step < KB; ++step) {
        const int kb = backwards ? KB - 1 - step : step;  // the accumulation is order-free for the issuer: it only counts blocks
        if constexpr (kCtaGroup == 2) mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
        else mbar_wait(&empty_bar[stage], phase ^ 1);
        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * kCtaGroup);
        uint8_t* sa = smem_a + stage * Cfg::kABytes;
        uint8_t* sb = smem_b + stage * Cfg::kBBytes;
        const CUtensorMap* tA = kb < KBm ? tAm : tAe;
        const CUtensorMap* tB = kb < KBm ? tBm : tBe;
        const int k_col = (kb < KBm ? kb : kb - KBm) * kGemmBlockK;
        if (p.conv.mode) {
          const int tap = kb / p.conv.cin_blocks, cb = kb - tap * p.conv.cin_blocks;
          const int ky = tap / 3, kx = tap - 3 * ky;
          if (p.conv.mode == 1)
            tma_load_4d<kCtaGroup == 2>(tA, &full_bar[stage], sa, cb * kGemmBlockK, cx0 + kx - 1, cy0 + ky - 1, cimg, hint_a);
          else  // input pixel (2y + ky, 2x + kx) = (row parity ky & 1 of row pair y + (ky >> 1), same in x)
            tma_load_5d<kCtaGroup == 2>(tA, &full_bar[stage], sa, cb * kGemmBlockK, kx & 1, cx0 + (kx >> 1), ky & 1, cy0 + (ky >> 1), hint_a);
          if constexpr (kCtaGroup == 2) tma_load_2d_2sm(tB, &full_bar[stage], sb, k_col, n0, hint_b);
          else tma_load_2d(tB, &full_bar[stage], sb, k_col, n0, hint_b);
        } else if constexpr (kCtaGroup == 2) {
          tma_load_2d_2sm(tA, &full_bar[stage], sa, k_col, m0, hint_a);
          tma_load_2d_2sm(tB, &full_bar[stage], sb, k_col, n0, hint_b);
        } else {
          tma_load_2d(tA, &full_bar[stage], sa, k_col, m0, hint_a);
          tma_load_2d(tB, &full_bar[stage], sb, k_col, n0, hint_b);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp walks the loop in a warp-uniform context and one elected lane issues: under `lane == 0` the
    // compiler wraps every tcgen05.mma in a value-broadcast loop (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~15 dependent
    // instructions per MMA); this way the descriptors live in uniform registers and the MMAs issue back to back.
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc_bf16(Cfg::kTileM, kBN, 0, 0);
      const bool issuer = elect_one();
      const uint64_t da0 = make_smem_desc(smem_u32(smem_a), 16, 1024, kLayoutSW128);
      const uint64_t db0 = make_smem_desc(smem_u32(smem_b), 16, 1024, kLayoutSW128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        if constexpr (kCtaGroup == 2) mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
        else mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * 256);
        for (int kb = 0; kb < KB; ++kb) {
          if constexpr (kCtaGroup == 2) mbar_wait_cluster(&full_bar[stage], phase);
          else mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = da0 + uint64_t(stage * (Cfg::kABytes / 16));
          const uint64_t db = db0 + uint64_t(stage * (Cfg::kBBytes / 16));
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kGemmBlockK / 16; ++k) {
              // +32 bytes (= 2 in 16-byte units) per UMMA_K step inside the 128B swizzle atom
              umma_ss<kCtaGroup>(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0);
            }
            if constexpr (kCtaGroup == 2) umma_commit_2sm(&empty_bar[stage], 0b11);
            else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (issuer) {
          if constexpr (kCtaGroup == 2) umma_commit_2sm(&tmem_full_bar[acc], 0b11);
          else umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;  // which half of the tile's 256 columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      int mi, ni;
      gemm_tile_coords(t, MT, NT, p.m_band, mi, ni);
      const int grp = (mi < mt0) ? 0 : 1;
      int m_local = (grp ? mi - mt0 : mi) * Cfg::kTileM + int(cta_rank) * 128 + quad * 32 + lane;
      if (p.conv.mode) {  // row r of the patch is output pixel (y0 + r / 16, x0 + r % 16) of image `img`
        const int patch = mi * kCtaGroup + int(cta_rank), r = quad * 32 + lane;
        const int per = p.conv.tiles_x * p.conv.tiles_y;
        const int img = patch / per, pr = patch - img * per;
        const int y = (pr / p.conv.tiles_x) * kConvPatchH + r / kConvPatchW, x = (pr % p.conv.tiles_x) * kConvPatchW + r % kConvPatchW;
        m_local = (patch < p.conv.n_patches && y < p.conv.H && x < p.conv.W) ? (img * p.conv.H + y) * p.conv.W + x : p.g[0].M;
      }
      const int n_tile0 = ni * kBN;

      if constexpr (kCtaGroup == 2) mbar_wait_cluster(&tmem_full_bar[acc], acc_phase);
      else mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * 256);

      gemm_epilogue_tile<kBN>(p, grp, m_local, n_tile0, t_acc, half, lane);
      // release this accumulator stage back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  // ---- teardown
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<kCtaGroup>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Multicast variant.  The plain 2-CTA kernel above is bound by L2->SM operand traffic (64 B/cycle/SM for 256x256x64
// tiles, ~6.2 KB/cycle chip-wide measured, profiles/r1a_ncu_summary.md).  Here kPN CTA pairs form one cluster of 2*kPN
// CTAs and compute kPN horizontally adjacent 256 x kBN tiles: they need the same A rows, so every CTA fetches only
// 1/kPN of its 128 A rows and TMA-multicasts it to the CTAs of the other pairs that hold the same rows.  Per-CTA
// traffic per k-block drops from 32 KB to 16/kPN + 16 KB.
//   * full[s]      in each pair leader, count 1 (its producer) + tx bytes of BOTH CTAs' stage: all loads use the
//                  cta_group::2 TMA form, which completes on the destination's pair leader (A quarters arrive from kPN
//                  senders).  (A first version relayed "peer stage full" with a remote mbarrier arrive per stage; the
//                  release.cluster arrive serialised the relay thread and halved throughput.)
//   * empty[s]     per CTA, count kPN: every pair leader's tcgen05.commit is multicast to all CTAs of the cluster,
//                  because a stage is overwritten by remote multicasts as well as by the CTA's own loads
template <int kBN, int kPN>
struct GemmMcCfg {
  static constexpr int kCluster = 2 * kPN;
  static constexpr int kARowsLoad = 128 / kPN;
  static constexpr int kABytes = 128 * kGemmBlockK * 2;
  static constexpr int kBRows = kBN / 2;
  static constexpr int kBBytes = kBRows * kGemmBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
};

template <int kBN, int kPN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_mc_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                       const __grid_constant__ GemmParams p) {
  using Cfg = GemmMcCfg<kBN, kPN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = int(rank >> 1);
  const int r = int(rank & 1);
  const bool is_leader = r == 0;
  const uint32_t leader_rank = rank & ~1u;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA0);
    prefetch_tensormap(&tmB0);
    if (p.num_groups > 1) {
      prefetch_tensormap(&tmA1);
      prefetch_tensormap(&tmB1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kPN);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kGemmEpiWarps * 2);
    }
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

  const int mt0 = (p.g[0].M + 255) / 256;
  const int mt1 = (p.num_groups > 1) ? (p.g[1].M + 255) / 256 : 0;
  const int MT = mt0 + mt1;
  const int NT = (p.N + kBN - 1) / kBN;
  const int NS = (NT + kPN - 1) / kPN;  // super columns of kPN tiles
  const int num_super = MT * NS;
  const int KB = p.K / kGemmBlockK;
  const int first = blockIdx.x / Cfg::kCluster;
  const int step = gridDim.x / Cfg::kCluster;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (every CTA) =====================
    uint16_t a_mask = 0;
#pragma unroll
    for (int q = 0; q < kPN; ++q) a_mask |= uint16_t(1u << (2 * q + r));
    int stage = 0;
    uint32_t phase = 0;
    for (int t = first; t < num_super; t += step) {
      const int mi = t % MT, ni = (t / MT) * kPN + pair;
      const int grp = (mi < mt0) ? 0 : 1;
      const int m0 = (grp ? mi - mt0 : mi) * 256 + r * 128 + pair * Cfg::kARowsLoad;
      const int n0 = ni * kBN + r * Cfg::kBRows;
      const CUtensorMap* tA = grp ? &tmA1 : &tmA0;
      const CUtensorMap* tB = grp ? &tmB1 : &tmB0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
        uint8_t* sa = smem_a + stage * Cfg::kABytes + pair * (Cfg::kARowsLoad * 128);
        uint8_t* sb = smem_b + stage * Cfg::kBBytes;
        if constexpr (kPN == 1) tma_load_2d_2sm(tA, &full_bar[stage], sa, kb * kGemmBlockK, m0, kEvictNormal);
        else tma_load_2d_mcast_2sm(tA, &full_bar[stage], sa, kb * kGemmBlockK, m0, a_mask, kEvictNormal);
        tma_load_2d_2sm(tB, &full_bar[stage], sb, kb * kGemmBlockK, n0, kEvictNormal);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer (pair leaders) =====================
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, kBN, 0, 0);
      constexpr uint16_t all_mask = uint16_t((1u << Cfg::kCluster) - 1);
      const uint16_t pair_mask = uint16_t(3u << (2 * pair));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = first; t < num_super; t += step) {
        mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * 256);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes), 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes), 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kGemmBlockK / 16; ++k)
            umma_ss<2>(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0);
          umma_commit_2sm(&empty_bar[stage], all_mask);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full_bar[acc], pair_mask);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = first; t < num_super; t += step) {
      const int mi = t % MT, ni = (t / MT) * kPN + pair;
      const int grp = (mi < mt0) ? 0 : 1;
      const int m_local = (grp ? mi - mt0 : mi) * 256 + r * 128 + quad * 32 + lane;
      mbar_wait_cluster(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * 256);
      if (ni < NT) gemm_epilogue_tile<kBN>(p, grp, m_local, ni * kBN, t_acc, half, lane);  // phantom tiles only keep the protocol going
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], leader_rank);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace tfx
