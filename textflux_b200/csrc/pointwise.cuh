// HBM/latency-bound kernels of the hot path: sinusoidal embeddings, small-batch GEMV (temb MLPs and the 344*D-row
// adaLN modulation matrix), LayerNorm+modulate, RoPE table, Euler update.  The embedding / GEMV / Euler kernels keep the
// reference's bf16 rounding points (each eager op of the reference returns bf16); LayerNorm+modulate stays in fp32.
#pragma once
#include "ptx.cuh"

namespace tfx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) applied to  bf16(bf16(t) * 1000)
// (transformer_flux.py:1088-1090 then embeddings.py:27-78,1330-1334).  out [B,256] bf16 = [cos | sin].
// in_is_f32: guidance arrives fp32, timestep arrives bf16.
// skip: optional device flag; the kernel does nothing when *skip >= 0 (modulation cache hit, see mod_cache_lookup_kernel)
__global__ void timestep_embed_kernel(const void* __restrict__ t_in, int in_is_f32, int B, __nv_bfloat16* __restrict__ out,
                                      const int* __restrict__ skip = nullptr) {
  const int b = blockIdx.x;
  const int k = threadIdx.x;  // 0..127
  if (b >= B || k >= 128) return;
  if (skip && *skip >= 0) return;
  float t = in_is_f32 ? reinterpret_cast<const float*>(t_in)[b] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(t_in)[b]);
  t = bf16_round(bf16_round(t) * 1000.0f);
  const float exponent = (-9.210340371976184f * float(k)) / 128.0f;  // -ln(10000) * k / half_dim, fp32 like torch
  const float arg = t * expf(exponent);
  out[b * 256 + k] = __float2bfloat16_rn(cosf(arg));
  out[b * 256 + 128 + k] = __float2bfloat16_rn(sinf(arg));
}

// ---------------------------------------------------------------------------------------------------------------
// Small-batch GEMV:  y[b, n] = post( sum_k pre(x[b, k]) * W[n, k] + bias[n] )  ;  optional  out = bf16(out + y)
// One warp per output row n, all (<= 8) batch rows at once; W streamed once with 16-byte loads.
// Used for TimestepEmbedding / PixArtAlphaTextProjection (embeddings.py:1009-1021,1927-1932) and for all 77 adaLN
// `linear(silu(temb))` calls at once (normalization.py:167,200,363) via the concatenated [344*D, D] matrix.
enum GemvFlags : int { GEMV_PRE_SILU = 1, GEMV_POST_SILU = 2, GEMV_ADD_TO_OUT = 4 };
constexpr int kGemvMaxB = 8;

__global__ void __launch_bounds__(256) gemv_kernel(const __nv_bfloat16* __restrict__ x, int B, int K,
                                                  const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ bias,
                                                  long long N, __nv_bfloat16* out, int flags,
                                                  const int* __restrict__ skip = nullptr) {
  extern __shared__ __nv_bfloat16 xs[];  // [B][K], pre-activation applied (bf16 like the reference's silu output)
  if (skip && *skip >= 0) return;  // modulation cache hit: the vectors this pass would produce are already stored
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    float v = __bfloat162float(x[i]);
    if (flags & GEMV_PRE_SILU) v = bf16_round(silu(v));
    xs[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long num_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int kvec = K >> 3;  // uint4 per row
  for (long long n = warp_global; n < N; n += num_warps) {
    const uint4* wrow = reinterpret_cast<const uint4*>(W + n * K);
    float acc[kGemvMaxB];
#pragma unroll
    for (int b = 0; b < kGemvMaxB; ++b) acc[b] = 0.f;
    for (int v0 = lane; v0 < kvec; v0 += 128) {
      uint4 w4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int vi = v0 + u * 32;
        w4[u] = (vi < kvec) ? __ldg(wrow + vi) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int vi = v0 + u * 32;
        if (vi >= kvec) continue;
        const uint32_t wu[4] = {w4[u].x, w4[u].y, w4[u].z, w4[u].w};
#pragma unroll
        for (int b = 0; b < kGemvMaxB; ++b) {
          if (b >= B) break;
          const uint4 xv = *reinterpret_cast<const uint4*>(xs + b * K + vi * 8);
          const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[b] = fmaf(bf16_lo(wu[e]), bf16_lo(xu[e]), acc[b]);
            acc[b] = fmaf(bf16_hi(wu[e]), bf16_hi(xu[e]), acc[b]);
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < kGemvMaxB; ++b) {
      if (b >= B) break;
      float s = warp_sum(acc[b]);
      if (lane == 0) {
        float y = bf16_round(s + __bfloat162float(bias[n]));
        if (flags & GEMV_POST_SILU) y = bf16_round(silu(y));
        if (flags & GEMV_ADD_TO_OUT) y = __bfloat162float(out[(long long)b * N + n]) + y;
        out[(long long)b * N + n] = __float2bfloat16_rn(y);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// y = LayerNorm(x, eps=1e-6, no affine) * (1 + scale) + shift      (normalization.py:169,201,365;
// transformer_flux.py:820-821,833-834).  One warp per row, the row lives in registers (kVec uint4 per lane).
// Rows [0, rows0) belong to samples of rows_per0 rows and read (shift,scale) at mod + b*mod_stride + {shift0,scale0};
// rows [rows0, rows) likewise with rows_per1 / {shift1,scale1}  (text stream / image stream).
struct LnModParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  int rows, D;
  int row_begin;  // first row to process (norm_out touches image rows only)
  int rows0, rows_per0, rows_per1;
  const __nv_bfloat16* mod;
  long long mod_stride;
  long long shift0, scale0, shift1, scale1;
  float eps;
};

// The row stays in registers as packed bf16 (kVec uint4 per lane = 48 registers at D = 3072) and is unpacked again in
// each of the three passes (mean, centred sum of squares, output): an fp32 copy of the row cost 96 more registers,
// 3 blocks per SM and a second, almost empty wave at 2560 rows (640 blocks on 444 slots); this form keeps 5 blocks
// per SM resident so the whole grid is one wave.
// bf16 -> fp32 unpack the compiler may not common up across the three passes (it would keep the fp32 row live)
__device__ __forceinline__ float lo_nocse(uint32_t u) { float f; asm volatile("shl.b32 %0, %1, 16;" : "=f"(f) : "r"(u)); return f; }
__device__ __forceinline__ float hi_nocse(uint32_t u) { float f; asm volatile("and.b32 %0, %1, 0xffff0000;" : "=f"(f) : "r"(u)); return f; }

__device__ __forceinline__ uint4 ldg_nc_v4_volatile(const uint4* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

template <int kVec>
__global__ void __launch_bounds__(128, (kVec <= 12) ? 5 : 2) ln_modulate_kernel(const LnModParams p) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = p.row_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  int b;
  long long shift_off, scale_off;
  if (row < p.rows0) {
    b = row / p.rows_per0; shift_off = p.shift0; scale_off = p.scale0;
  } else {
    b = (row - p.rows0) / p.rows_per1; shift_off = p.shift1; scale_off = p.scale1;
  }
  const uint4* sh = reinterpret_cast<const uint4*>(p.mod + (long long)b * p.mod_stride + shift_off);
  const uint4* sc = reinterpret_cast<const uint4*>(p.mod + (long long)b * p.mod_stride + scale_off);
  const uint4* xr = reinterpret_cast<const uint4*>(p.x + (long long)row * p.D);
  pdl_wait();  // x (and, on the unscheduled path, mod) come from the previous kernels
  uint4 u[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) u[i] = xr[lane + i * 32];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) sum += lo_nocse(w[e]) + hi_nocse(w[e]);
  }
  const float mean = warp_sum(sum) / float(p.D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d0 = lo_nocse(w[e]) - mean, d1 = hi_nocse(w[e]) - mean;
      sq = fmaf(d0, d0, sq);
      sq = fmaf(d1, d1, sq);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / float(p.D) + p.eps);
  uint4* yr = reinterpret_cast<uint4*>(p.y + (long long)row * p.D);
  // fp32 throughout, one rounding at the store (the eager reference rounds to bf16 after each of its four ops; this
  // is never further from the fp32 result than the reference is -- tests/test_gpu_ops.py::test_ln_modulate)
  // (shift, scale) are fetched two chunks ahead through volatile asm so that the compiler cannot hoist all 2*kVec loads
  // above the reductions (that costs 96 registers and spills)
  uint4 s4[2], c4[2];
  s4[0] = ldg_nc_v4_volatile(sh + lane);
  c4[0] = ldg_nc_v4_volatile(sc + lane);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    if (i + 1 < kVec) {
      s4[(i + 1) & 1] = ldg_nc_v4_volatile(sh + lane + (i + 1) * 32);
      c4[(i + 1) & 1] = ldg_nc_v4_volatile(sc + lane + (i + 1) * 32);
    }
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
    const uint32_t su[4] = {s4[i & 1].x, s4[i & 1].y, s4[i & 1].z, s4[i & 1].w};
    const uint32_t cu[4] = {c4[i & 1].x, c4[i & 1].y, c4[i & 1].z, c4[i & 1].w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float y0 = fmaf((lo_nocse(w[e]) - mean) * rstd, 1.0f + bf16_lo(cu[e]), bf16_lo(su[e]));
      const float y1 = fmaf((hi_nocse(w[e]) - mean) * rstd, 1.0f + bf16_hi(cu[e]), bf16_hi(su[e]));
      o[e] = pack_bf16(y0, y1);
    }
    yr[lane + i * 32] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// FluxPosEmbed (embeddings.py:952-973): float64 frequencies and angles, cast to fp32.  Output [n_joint, dh/2] of
// (cos, sin) - the reference's repeat_interleave(2) duplicates are not stored.  ids arrive bf16 [*,3].
struct RopeParams {
  const __nv_bfloat16* txt_ids;
  const __nv_bfloat16* img_ids;
  int T, S;
  int axes[3];
  int half_dim;  // dh/2
  float2* out;
};
__global__ void rope_table_kernel(const RopeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = (p.T + p.S) * p.half_dim;
  if (idx >= total) return;
  const int tok = idx / p.half_dim;
  int j = idx - tok * p.half_dim;
  int axis = 0;
  while (axis < 2 && j >= p.axes[axis] / 2) { j -= p.axes[axis] / 2; ++axis; }
  const __nv_bfloat16* ids = (tok < p.T) ? p.txt_ids + tok * 3 : p.img_ids + (tok - p.T) * 3;
  const double pos = double(__bfloat162float(ids[axis]));
  const double dim = double(p.axes[axis]);
  const double freq = 1.0 / pow(10000.0, double(2 * j) / dim);
  const double ang = pos * freq;
  p.out[idx] = make_float2(float(cos(ang)), float(sin(ang)));
}

// ---------------------------------------------------------------------------------------------------------------
// FlowMatchEulerDiscreteScheduler.step (scheduling_flow_match_euler_discrete.py:322-330):
//   prev = bf16( float(sample) + float( bf16( float(bf16(dt)) * float(v) ) ) )
__global__ void euler_step_kernel(const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ x,
                                  __nv_bfloat16* __restrict__ out, long long n, float dt_bf16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dv = bf16_round(dt_bf16 * __bfloat162float(v[i]));
  out[i] = __float2bfloat16_rn(__bfloat162float(x[i]) + dv);
}

// StochasticRFOvershotDiscreteScheduler.step, attn_map = None (scheduling_stochastic_rf_discrete_overshot.py:339-361):
//   x_o  = fp32(x) + bf16( bf16(t_o - t) * (-v) )          (0-dim fp32 scalar times bf16 tensor is a bf16 product)
//   prev = bf16( x_o * a + eps * b )                        (two fp32 products, one fp32 sum, no contraction)
//   x1   = fp32(x) - bf16( bf16(sigma) * v )
__global__ void overshoot_step_kernel(const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ x,
                                      const float* __restrict__ eps, __nv_bfloat16* __restrict__ prev, float* __restrict__ x1,
                                      long long n, float coef_bf16, float a, float b, float sigma_bf16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float vf = __bfloat162float(v[i]);
  const float xf = __bfloat162float(x[i]);
  const float xo = __fadd_rn(xf, bf16_round(__fmul_rn(coef_bf16, -vf)));
  prev[i] = __float2bfloat16_rn(__fadd_rn(__fmul_rn(xo, a), __fmul_rn(eps[i], b)));
  if (x1) x1[i] = __fsub_rn(xf, bf16_round(__fmul_rn(sigma_bf16, vf)));
}

__global__ void set_float_kernel(float* dst, float value) { *dst = value; }

// ---------------------------------------------------------------------------------------------------------------
// Modulation cache of the drop-in path.  temb and all adaLN vectors depend only on (timestep, guidance, pooled)
// (embeddings.py:1327-1339, normalization.py:167,200,363), and the unmodified pipeline hands the same triples over for
// every image it samples with the same schedule and prompt template (pipeline_flux_fill.py:2082-2094).  The engine
// therefore keeps the last `slots` results on the device, keyed on the exact input bits, and a captured step stays one
// static graph: lookup -> (modulation kernels that return at once on a hit) -> commit.
//   state[0] = slot that matches this step's inputs, or -1      state[1] = slot a miss is stored into
//   state[2] = number of valid slots                             state[3] = hits so far (statistics)
//   state[4] = round-robin replacement cursor
struct ModCacheParams {
  const __nv_bfloat16* t;       // [B] bf16
  const float* g;               // [B] fp32 or null
  const __nv_bfloat16* pooled;  // [B, P]
  int B, P, slots;
  uint16_t* key_t;              // [slots, B]
  uint32_t* key_g;              // [slots, B]
  uint16_t* key_p;              // [slots, B*P]
  int* state;
  __nv_bfloat16* mod;           // [B * mod_rows] the step's modulation vectors
  __nv_bfloat16* table;         // [slots, B * mod_rows]
  long long mod_elems;          // B * mod_rows
};

__global__ void __launch_bounds__(256) mod_cache_lookup_kernel(const ModCacheParams p) {
  __shared__ int s_hit;
  const uint16_t* t = reinterpret_cast<const uint16_t*>(p.t);
  const uint32_t* g = reinterpret_cast<const uint32_t*>(p.g);
  const uint16_t* pl = reinterpret_cast<const uint16_t*>(p.pooled);
  const int valid = p.state[2];
  if (threadIdx.x == 0) s_hit = -1;
  __syncthreads();
  for (int s = 0; s < valid; ++s) {
    int ok = 1;
    for (int i = threadIdx.x; i < p.B; i += blockDim.x)
      ok &= (p.key_t[s * p.B + i] == t[i]) && (!g || p.key_g[s * p.B + i] == g[i]);
    for (int i = threadIdx.x; ok && i < p.B * p.P; i += blockDim.x) ok &= p.key_p[(long long)s * p.B * p.P + i] == pl[i];
    if (__syncthreads_and(ok)) {
      if (threadIdx.x == 0) s_hit = s;
      break;  // block-uniform
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int hit = s_hit;
    p.state[0] = hit;
    if (hit >= 0) {
      p.state[3] += 1;
    } else {
      const int ins = p.state[4] % p.slots;  // round-robin replacement
      p.state[1] = ins;
      p.state[4] = ins + 1;
      if (valid < p.slots) p.state[2] = valid + 1;
    }
  }
}

// hit: mod <- table[hit]; miss: table[insert] <- mod and the slot's keys <- this step's inputs
__global__ void __launch_bounds__(256) mod_cache_commit_kernel(const ModCacheParams p) {
  const int hit = p.state[0];
  const long long n4 = p.mod_elems / 8;  // uint4 = 8 bf16; mod_rows is a multiple of 256
  uint4* mod4 = reinterpret_cast<uint4*>(p.mod);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (hit >= 0) {
    const uint4* src = reinterpret_cast<const uint4*>(p.table + (long long)hit * p.mod_elems);
    for (long long i = tid; i < n4; i += nth) mod4[i] = src[i];
    return;
  }
  const int ins = p.state[1];
  uint4* dst = reinterpret_cast<uint4*>(p.table + (long long)ins * p.mod_elems);
  for (long long i = tid; i < n4; i += nth) dst[i] = mod4[i];
  if (blockIdx.x == 0) {
    const uint16_t* t = reinterpret_cast<const uint16_t*>(p.t);
    const uint32_t* g = reinterpret_cast<const uint32_t*>(p.g);
    const uint16_t* pl = reinterpret_cast<const uint16_t*>(p.pooled);
    for (int i = threadIdx.x; i < p.B; i += blockDim.x) {
      p.key_t[ins * p.B + i] = t[i];
      p.key_g[ins * p.B + i] = g ? g[i] : 0u;
    }
    for (int i = threadIdx.x; i < p.B * p.P; i += blockDim.x) p.key_p[(long long)ins * p.B * p.P + i] = pl[i];
  }
}

}  // namespace tfx
