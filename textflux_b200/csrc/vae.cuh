// Pointwise / reduction kernels of the AutoencoderKL path (reference: models/autoencoders/vae.py Encoder /
// Decoder, models/resnet.py ResnetBlock2D, models/upsampling.py, models/attention_processor.py AttnProcessor2_0 for the mid block).
// Activations live as NHWC bf16 ([B, H, W, C], C a multiple of 64) so a pixel row is the K-major row the tcgen05 GEMM wants: the 3x3
// convolutions are the implicit-GEMM mode of gemm.cuh, 1x1 convolutions and the attention projections are plain GEMMs on [B*H*W, C].
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace tfx {

// ---- layout: NCHW (fp32 | bf16) -> NHWC bf16 with the channels zero-padded to Cp (a multiple of 8) -----------------------------
template <typename Tin>
__global__ void nchw_to_nhwc_kernel(const Tin* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int C, long long HW, int Cp) {
  const int vecs = Cp / 8;
  const long long total = (long long)B * HW * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % vecs);
    const long long pix = (i / vecs) % HW;
    const long long b = i / (vecs * HW);
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      x[j] = 0.f;
      if (c < C) {
        if constexpr (sizeof(Tin) == 4) x[j] = bf16_round(float(src[(b * C + c) * HW + pix]));
        else x[j] = __bfloat162float(src[(b * C + c) * HW + pix]);
      }
    }
    uint4 o;
    o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

// NHWC bf16 rows of stride ld -> NCHW bf16 [B, C, H, W] (C small: the 3 image channels or the 2 * latent_channels moments)
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, long long ld, __nv_bfloat16* __restrict__ dst, int B, int C,
                                    long long HW) {
  const long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % HW;
    const int c = int((i / HW) % C);
    const long long b = i / (HW * C);
    dst[i] = src[(b * HW + pix) * ld + c];
  }
}

// ---- nearest-neighbour 2x upsampling (Upsample2D: F.interpolate(scale_factor=2.0, mode="nearest")) on NHWC ----------------------
__global__ void upsample_nearest2x_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int H, int W, int C) {
  const int vecs = C / 8;
  const long long total = (long long)B * (2 * H) * (2 * W) * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % vecs);
    long long r = i / vecs;
    const int x = int(r % (2 * W)); r /= (2 * W);
    const int y = int(r % (2 * H));
    const long long b = r / (2 * H);
    reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + ((b * H + (y >> 1)) * W + (x >> 1)) * vecs + v);
  }
}

// ---- GroupNorm (+ SiLU) on NHWC: nn.GroupNorm(num_groups, C, eps=1e-6, affine) followed by nn.SiLU --------------------------------
// Three launches: per-(image, pixel chunk) partial sums (fixed summation order: no atomics), finalize (partials -> mean, rstd in
// double), apply.  All arithmetic in fp32 / double with ONE rounding to bf16 at the store (the reference rounds after the norm and
// again after the activation).
struct GnParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  int B, C, G;
  long long HW;
  const __nv_bfloat16* gamma;
  const __nv_bfloat16* beta;
  float eps;
  int silu;
  float* partial;  // [B, nchunk, G, 2]
  float* stats;    // [B, G, 2] (mean, rstd)
  int nchunk;
};

constexpr int kGnThreads = 256;

__global__ void __launch_bounds__(kGnThreads) gn_partial_kernel(GnParams p) {
  __shared__ float red[kGnThreads * 16];  // [pixel lane][channel][2]
  const int vecs = p.C / 8;               // 16-byte vectors per pixel; kGnThreads % vecs == 0 (host-checked)
  const int lanes = kGnThreads / vecs;
  const int v = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  const int b = blockIdx.y;
  const long long per = (p.HW + p.nchunk - 1) / p.nchunk;
  const long long p0 = blockIdx.x * per, p1 = min(p0 + per, p.HW);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  const uint4* src = reinterpret_cast<const uint4*>(p.x) + (long long)b * p.HW * vecs;
  auto acc = [&](const uint4& u) {
    const float x[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += x[j]; q[j] = fmaf(x[j], x[j], q[j]); }
  };
  long long pix = p0 + pl;
  for (; pix + 3LL * lanes < p1; pix += 4LL * lanes) {  // four loads in flight per thread
    const uint4 u0 = __ldg(src + pix * vecs + v), u1 = __ldg(src + (pix + lanes) * vecs + v);
    const uint4 u2 = __ldg(src + (pix + 2LL * lanes) * vecs + v), u3 = __ldg(src + (pix + 3LL * lanes) * vecs + v);
    acc(u0); acc(u1); acc(u2); acc(u3);
  }
  for (; pix < p1; pix += lanes) acc(__ldg(src + pix * vecs + v));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(pl * p.C + v * 8 + j) * 2 + 0] = s[j];
    red[(pl * p.C + v * 8 + j) * 2 + 1] = q[j];
  }
  __syncthreads();
  if (threadIdx.x < p.G) {
    const int g = threadIdx.x, cg = p.C / p.G;
    float ss = 0.f, qq = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int c = g * cg; c < (g + 1) * cg; ++c) { ss += red[(l * p.C + c) * 2]; qq += red[(l * p.C + c) * 2 + 1]; }
    float* o = p.partial + (((long long)b * p.nchunk + blockIdx.x) * p.G + g) * 2;
    o[0] = ss; o[1] = qq;
  }
}

// one block of 128 threads per (group, image): a handful of independent loads per thread, then a fixed-order tree (deterministic)
__global__ void __launch_bounds__(128) gn_finalize_kernel(GnParams p) {
  __shared__ double red[2][4];
  const int g = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0.0, q = 0.0;
  for (int c = threadIdx.x; c < p.nchunk; c += 128) {
    const float2 o = *reinterpret_cast<const float2*>(p.partial + (((long long)b * p.nchunk + c) * p.G + g) * 2);
    s += o.x; q += o.y;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); q += __shfl_xor_sync(0xffffffffu, q, d); }
  if (lane == 0) { red[0][warp] = s; red[1][warp] = q; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  s = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
  q = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  const double n = double(p.HW) * (p.C / p.G);
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  p.stats[(b * p.G + g) * 2 + 0] = float(mean);
  p.stats[(b * p.G + g) * 2 + 1] = float(1.0 / sqrt(var + double(p.eps)));
}

__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(GnParams p) {
  const int vecs = p.C / 8;
  const int b = blockIdx.y;
  const int cg = p.C / p.G;
  // this thread's 8 channels are the same for every vector it visits (stride is a multiple of vecs)
  const long long i0 = blockIdx.x * (long long)kGnThreads + threadIdx.x;
  const int v = int(i0 % vecs);
  float a[8], c[8];  // y = x * a + c
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = v * 8 + j, g = ch / cg;
    const float mean = p.stats[(b * p.G + g) * 2], rstd = p.stats[(b * p.G + g) * 2 + 1];
    const float ga = __bfloat162float(p.gamma[ch]), be = __bfloat162float(p.beta[ch]);
    a[j] = rstd * ga;
    c[j] = be - mean * rstd * ga;
  }
  const long long total = p.HW * vecs;
  const uint4* src = reinterpret_cast<const uint4*>(p.x) + (long long)b * total;
  uint4* dst = reinterpret_cast<uint4*>(p.y) + (long long)b * total;
  auto apply = [&](const uint4& u) {
    float x[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = fmaf(x[j], a[j], c[j]);
      if (p.silu) x[j] = __fdividef(x[j], 1.0f + __expf(-x[j]));  // approximate division: 2 ulp in fp32, invisible after the bf16 store
    }
    uint4 o;
    o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
    return o;
  };
  const long long stride = (long long)gridDim.x * kGnThreads;
  long long i = i0;
  for (; i + 3 * stride < total; i += 4 * stride) {  // four loads in flight per thread
    const uint4 u0 = __ldg(src + i), u1 = __ldg(src + i + stride), u2 = __ldg(src + i + 2 * stride), u3 = __ldg(src + i + 3 * stride);
    dst[i] = apply(u0); dst[i + stride] = apply(u1); dst[i + 2 * stride] = apply(u2); dst[i + 3 * stride] = apply(u3);
  }
  for (; i < total; i += stride) dst[i] = apply(__ldg(src + i));
}

// ---- row softmax of fp32 scores -> bf16 probabilities: P = softmax(scale * S) (F.scaled_dot_product_attention of the mid block) ---
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long lds, __nv_bfloat16* __restrict__ P,
                                                           long long ldp, int n, float scale_log2) {
  __shared__ float red[8];
  const float* s = S + blockIdx.x * lds;
  __nv_bfloat16* o = P + blockIdx.x * ldp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, s[i]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) l += exp2f((s[i] - m) * scale_log2);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) l += red[w];
  const float inv = 1.0f / l;
  for (int i = threadIdx.x; i < n; i += 256) o[i] = __float2bfloat16_rn(exp2f((s[i] - m) * scale_log2) * inv);
}

// ---- DiagonalGaussianDistribution(moments).sample() given the N(0,1) draw (models/autoencoders/vae.py:780-803), with the bf16
// rounding points of the eager ops: std = bf16(exp(bf16(0.5 * clamp(logvar, -30, 20)))); x = bf16(mean + bf16(std * noise)) ------------
__global__ void gaussian_sample_kernel(const __nv_bfloat16* __restrict__ moments, const __nv_bfloat16* __restrict__ noise,
                                       __nv_bfloat16* __restrict__ out, int B, int L, long long HW) {
  const long long total = (long long)B * L * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / (L * HW), r = i - b * L * HW;
    const float mean = __bfloat162float(moments[b * 2 * L * HW + r]);
    float lv = __bfloat162float(moments[b * 2 * L * HW + L * HW + r]);
    lv = fminf(fmaxf(lv, -30.f), 20.f);
    const float sd = bf16_round(expf(bf16_round(0.5f * lv)));  // libdevice expf, as CUDA eager's exp kernel calls it
    out[i] = __float2bfloat16_rn(mean + bf16_round(sd * __bfloat162float(noise[i])));
  }
}

}  // namespace tfx
