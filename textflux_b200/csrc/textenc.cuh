// Kernels of the prompt encoders (pipeline_flux_fill.py:1411-1503 calls transformers' T5EncoderModel and
// CLIPTextModel): embedding gather, T5LayerNorm (RMS) / LayerNorm over rows, and attention for short sequences (T <= 512, head_dim
// 64) with T5's relative-position bias or CLIP's causal mask.  Every Linear of both models is the tcgen05 GEMM of gemm.cuh.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <mma.h>

#include "ptx.cuh"

namespace tfx {

// out[r, :] = table[ids[r], :] (+ pos[r % T, :]): nn.Embedding lookups of T5Stack.embed_tokens / CLIPTextEmbeddings
__global__ void embed_rows_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ pos,
                                  __nv_bfloat16* __restrict__ out, int rows, int T, int D, int vocab) {
  const int vecs = D / 8;
  const long long total = (long long)rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = int(i / vecs), v = int(i % vecs);
    int id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    uint4 u = __ldg(reinterpret_cast<const uint4*>(table + (long long)id * D) + v);
    if (pos) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(pos + (long long)(r % T) * D) + v);
      u.x = pack_bf16(bf16_lo(u.x) + bf16_lo(q.x), bf16_hi(u.x) + bf16_hi(q.x));
      u.y = pack_bf16(bf16_lo(u.y) + bf16_lo(q.y), bf16_hi(u.y) + bf16_hi(q.y));
      u.z = pack_bf16(bf16_lo(u.z) + bf16_lo(q.z), bf16_hi(u.z) + bf16_hi(q.z));
      u.w = pack_bf16(bf16_lo(u.w) + bf16_lo(q.w), bf16_hi(u.w) + bf16_hi(q.w));
    }
    reinterpret_cast<uint4*>(out)[i] = u;
  }
}

// One warp per row of D elements (D % 256 == 0).
//   kRms:  T5LayerNorm (modeling_t5.py T5LayerNorm.forward): y = w * bf16(x * rsqrt(mean(x^2) + eps)), statistics in fp32
//   !kRms: nn.LayerNorm(D, eps) with affine: y = bf16((x - mean) * rstd * w + b), fp32 throughout
template <bool kRms>
__global__ void __launch_bounds__(256) norm_rows_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                        const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int rows, int D,
                                                        float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + (long long)row * D);
  uint4* dst = reinterpret_cast<uint4*>(y + (long long)row * D);
  const int nv = D / 256;  // 16-byte vectors per lane
  float s = 0.f, q = 0.f;
  for (int i = 0; i < nv; ++i) {
    const uint4 u = src[i * 32 + lane];
    const float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j]; q = fmaf(v[j], v[j], q); }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); q += __shfl_xor_sync(0xffffffffu, q, d); }
  const float mean = kRms ? 0.f : s / D;
  const float var = kRms ? q / D : fmaxf(q / D - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int i = 0; i < nv; ++i) {
    const uint4 u = src[i * 32 + lane];
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(w) + i * 32 + lane);
    float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
    const float gw[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y), bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
    if constexpr (kRms) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gw[j] * bf16_round(v[j] * rstd);
    } else {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(b) + i * 32 + lane);
      const float hb[8] = {bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y), bf16_lo(h.z), bf16_hi(h.z), bf16_lo(h.w), bf16_hi(h.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf((v[j] - mean) * rstd, gw[j], hb[j]);
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    dst[i * 32 + lane] = o;
  }
}

// ---- attention over a short sequence, head_dim 64 -----------------------------------------------------------------------------
// One CTA per (32 query rows, head, sample), 4 warps.  q, k, v are column slices of one token-major matrix (the fused QKV
// projection): row (b * T + t), stride ld.  Scores of the 32 rows against all Tp (= T rounded up to 16) keys go to shared memory in
// fp32 (bf16 HMMA through nvcuda::wmma, operands straight from global / L2), bias or mask is applied, softmax in fp32, P in bf16,
// then O = P V.  ~4 GFLOP per T5-XXL layer against ~200 GFLOP of projections, so the legacy tensor path is enough here.
//   kCausal = false (T5Attention.forward): scores + relative_attention_bias[bucket(j - i), h], no scaling, keys j >= T masked
//   kCausal = true  (CLIPAttention / eager_attention_forward): scores * scale, keys j > i masked
struct SmallAttnParams {
  const __nv_bfloat16 *q, *k, *v;  // first element of column block 0 of each (head h adds h * 64 columns)
  long long ld;                    // row stride of q / k / v
  __nv_bfloat16* out;              // [B * T, ld_out], head h at columns h * 64
  long long ld_out;
  int B, H, T, Tp;
  float scale;
  const __nv_bfloat16* rel_table;  // T5: [num_buckets, H]
  const int32_t* rel_lut;          // T5: bucket of relative position d = j - i at index d + T - 1, [2T - 1]
};

constexpr int kSmallAttnRows = 32;
// row stride of the fp32 score tile: at least 72 floats, because the tile is reused as the [32][72] fp32 staging area of O
__host__ __device__ inline int small_attn_lds(int Tp) { return Tp + 8 > 72 ? Tp + 8 : 72; }
inline size_t small_attn_smem(int Tp) { return (size_t)kSmallAttnRows * small_attn_lds(Tp) * 4 + (size_t)kSmallAttnRows * (Tp + 16) * 2 + (size_t)2 * Tp * 4; }

template <bool kCausal>
__global__ void __launch_bounds__(128) small_attention_kernel(SmallAttnParams p) {
  using namespace nvcuda;
  extern __shared__ __align__(32) uint8_t sm_raw[];
  const int Tp = p.Tp, lds = small_attn_lds(Tp), ldp = Tp + 16;
  float* S = reinterpret_cast<float*>(sm_raw);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(sm_raw + (size_t)kSmallAttnRows * lds * 4);
  float* relv = reinterpret_cast<float*>(sm_raw + (size_t)kSmallAttnRows * lds * 4 + (size_t)kSmallAttnRows * ldp * 2);
  const int q0 = blockIdx.x * kSmallAttnRows, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)b * p.T;
  const __nv_bfloat16* Q = p.q + (row0 + q0) * p.ld + h * 64;
  const __nv_bfloat16* K = p.k + row0 * p.ld + h * 64;
  const __nv_bfloat16* V = p.v + row0 * p.ld + h * 64;
  if constexpr (!kCausal) {
    for (int i = threadIdx.x; i < 2 * p.T - 1; i += 128) relv[i] = __bfloat162float(p.rel_table[(long long)p.rel_lut[i] * p.H + h]);
  }
  // ---- S = Q K^T: warp w owns key blocks w, w + 4, ...
  {
    wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> a[2][4];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) wmma::load_matrix_sync(a[rb][kk], Q + (long long)rb * 16 * p.ld + kk * 16, (unsigned)p.ld);
    for (int cb = warp; cb < Tp / 16; cb += 4) {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2];
      wmma::fill_fragment(acc[0], 0.f);
      wmma::fill_fragment(acc[1], 0.f);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> bk;
        wmma::load_matrix_sync(bk, K + (long long)cb * 16 * p.ld + kk * 16, (unsigned)p.ld);
        wmma::mma_sync(acc[0], a[0][kk], bk, acc[0]);
        wmma::mma_sync(acc[1], a[1][kk], bk, acc[1]);
      }
      wmma::store_matrix_sync(S + cb * 16, acc[0], lds, wmma::mem_row_major);
      wmma::store_matrix_sync(S + 16 * lds + cb * 16, acc[1], lds, wmma::mem_row_major);
    }
  }
  __syncthreads();
  // ---- bias / mask, softmax (fp32), P in bf16: warp w owns rows 8w .. 8w + 7
  for (int r = warp * 8; r < warp * 8 + 8; ++r) {
    const int i = q0 + r;
    float* s = S + r * lds;
    float m = -INFINITY;
    for (int j = lane; j < Tp; j += 32) {
      float x = s[j];
      bool ok = j < p.T;
      if constexpr (kCausal) { x *= p.scale; ok = ok && j <= i; }
      else if (i < p.T && ok) x += relv[j - i + p.T - 1];
      x = ok ? x : -INFINITY;
      s[j] = x;
      m = fmaxf(m, x);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (m == -INFINITY) m = 0.f;  // a padded query row past T under the causal mask still sees key 0; this guards T = 0 rows only
    float l = 0.f;
    for (int j = lane; j < Tp; j += 32) {
      const float e = __expf(s[j] - m);
      s[j] = e;
      l += e;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d);
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    for (int j = lane; j < Tp; j += 32) P[r * ldp + j] = __float2bfloat16_rn(s[j] * inv);
  }
  __syncthreads();
  // ---- O = P V: warp w owns row block w / 2, column blocks 2 (w % 2), 2 (w % 2) + 1
  {
    const int rb = warp >> 1, cb0 = (warp & 1) * 2;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2];
    wmma::fill_fragment(acc[0], 0.f);
    wmma::fill_fragment(acc[1], 0.f);
    for (int kb = 0; kb < Tp / 16; ++kb) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> a;
      wmma::load_matrix_sync(a, P + rb * 16 * ldp + kb * 16, ldp);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> bv;
        wmma::load_matrix_sync(bv, V + (long long)kb * 16 * p.ld + (cb0 + c) * 16, (unsigned)p.ld);
        wmma::mma_sync(acc[c], a, bv, acc[c]);
      }
    }
    __syncthreads();  // everyone is done reading P / S before S is reused as the fp32 staging tile of O
    float* O = S;     // [32][72]
    wmma::store_matrix_sync(O + rb * 16 * 72 + cb0 * 16, acc[0], 72, wmma::mem_row_major);
    wmma::store_matrix_sync(O + rb * 16 * 72 + (cb0 + 1) * 16, acc[1], 72, wmma::mem_row_major);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kSmallAttnRows * 32; idx += 128) {
    const int r = idx / 32, c2 = idx % 32;
    if (q0 + r < p.T) {
      const float* o = S + r * 72 + c2 * 2;
      *reinterpret_cast<uint32_t*>(p.out + (row0 + q0 + r) * p.ld_out + h * 64 + c2 * 2) = pack_bf16(o[0], o[1]);
    }
  }
}

}  // namespace tfx
