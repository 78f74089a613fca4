// Conditioning glue on either side of the denoising loop: the 2x2 latent
// pack / unpack, the 8x8 mask pixel-unshuffle + pack, and the VAE latent (de)normalisation that brackets them.
// Pure index permutations of 2-byte elements -- HBM/latency bound, bit-exact by construction:
//   pack    (pipeline_flux_fill.py:1743-1748):  out[b, i*(w/2)+j, c*4 + di*2 + dj] = in[b, c, 2i+di, 2j+dj]
//   unpack  (:1752-1765):                        the inverse
//   mask    (:1563-1580):  out[b, i*(w/2)+j, (py*8+px)*4 + di*2 + dj] = mask[b, 0, (2i+di)*8+py, (2j+dj)*8+px]
//   affine  (:1536 / :2127):  y = (x - shift) * scale   before pack,   y = x / scale + shift   after unpack,
//           each torch op rounding to the tensor dtype (bf16 inputs: two roundings; fp32 inputs: one, at the store).
//           The python scalars stay fp32, as CUDA eager keeps them (opmath); CPU eager rounds the add/sub scalar to
//           bf16 first, so the reference itself differs by <= 1 bf16 ulp between devices
//           (tests/test_gpu_conditioning.py checks bit-exactness against CUDA eager and 1 ulp against the CPU golden).
// One thread per OUTPUT 2x2 patch channel group: writes are contiguous along the channel axis of a token (pack) or
// along w (unpack); reads of a warp touch at most 2 source rows.
#pragma once
#include "ptx.cuh"

namespace tfx {

template <typename T>
__device__ __forceinline__ float cond_load(const T* p);
template <>
__device__ __forceinline__ float cond_load<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float cond_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// (x - shift) * scale with the rounding points of the eager reference
template <typename T>
__device__ __forceinline__ float cond_affine_in(float x, float shift, float scale) {
  if (sizeof(T) == 2) return bf16_round(bf16_round(x - shift) * scale);
  return (x - shift) * scale;
}

// dst[b, s, dst_off + c*4 + di*2 + dj] (bf16, row stride dst_ld) = f(src[b, c, 2i+di, 2j+dj]);  src [B, C, h, w] contiguous
template <typename T>
__global__ void pack_latents_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long dst_ld,
                                    long long dst_off, int B, int C, int h, int w, int affine, float shift, float scale) {
  const int h2 = h >> 1, w2 = w >> 1;
  const long long total = (long long)B * h2 * w2 * C * 2;  // one thread per (b, token, c, di): two elements (dj = 0, 1)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int di = int(idx & 1);
    long long r = idx >> 1;
    const int c = int(r % C); r /= C;
    const int j = int(r % w2); r /= w2;
    const int i = int(r % h2);
    const int b = int(r / h2);
    const T* s = src + (((long long)b * C + c) * h + (2 * i + di)) * w + 2 * j;
    float x0 = cond_load<T>(s), x1 = cond_load<T>(s + 1);
    if (affine) { x0 = cond_affine_in<T>(x0, shift, scale); x1 = cond_affine_in<T>(x1, shift, scale); }
    __nv_bfloat16* d = dst + ((long long)b * h2 * w2 + (long long)i * w2 + j) * dst_ld + dst_off + c * 4 + di * 2;
    *reinterpret_cast<uint32_t*>(d) = pack_bf16(x0, x1);
  }
}

// dst[b, c, 2i+di, 2j+dj] = g(src[b, s, c*4 + di*2 + dj]);  src rows of stride src_ld, bf16 in and out
__global__ void unpack_latents_kernel(const __nv_bfloat16* __restrict__ src, long long src_ld, __nv_bfloat16* __restrict__ dst,
                                      int B, int C, int h, int w, int affine, float shift, float scale) {
  const int h2 = h >> 1, w2 = w >> 1;
  const long long total = (long long)B * C * h * w2;  // one thread per (b, c, y, j): two elements along w
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int j = int(r % w2); r /= w2;
    const int y = int(r % h); r /= h;
    const int c = int(r % C);
    const int b = int(r / C);
    const int i = y >> 1, di = y & 1;
    const uint32_t u = *reinterpret_cast<const uint32_t*>(src + ((long long)b * h2 * w2 + (long long)i * w2 + j) * src_ld + c * 4 + di * 2);
    float x0 = bf16_lo(u), x1 = bf16_hi(u);
    if (affine) {  // latents / scaling_factor + shift_factor, each op rounding to bf16 (pipeline_flux_fill.py:2127)
      x0 = bf16_round(bf16_round(__fdiv_rn(x0, scale)) + shift);
      x1 = bf16_round(bf16_round(__fdiv_rn(x1, scale)) + shift);
    }
    *reinterpret_cast<uint32_t*>(dst + (((long long)b * C + c) * h + y) * w + 2 * j) = pack_bf16(x0, x1);
  }
}

// mask [B, 1, H, W] (H = h*vs, W = w*vs, vs = vae scale factor 8) -> dst[b, s, dst_off + (py*vs+px)*4 + di*2 + dj]
template <typename T>
__global__ void pack_mask_kernel(const T* __restrict__ mask, __nv_bfloat16* __restrict__ dst, long long dst_ld, long long dst_off,
                                 int B, int h, int w, int vs) {
  const int h2 = h >> 1, w2 = w >> 1, W = w * vs, H = h * vs;
  const int ch = vs * vs;
  const long long total = (long long)B * h2 * w2 * ch * 2;  // (b, token, py*vs+px, di): two elements (dj = 0, 1)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int di = int(idx & 1);
    long long r = idx >> 1;
    const int pc = int(r % ch); r /= ch;
    const int j = int(r % w2); r /= w2;
    const int i = int(r % h2);
    const int b = int(r / h2);
    const int py = pc / vs, px = pc % vs;
    const T* m = mask + ((long long)b * H + (long long)(2 * i + di) * vs + py) * W + px;
    const float x0 = cond_load<T>(m + (long long)(2 * j) * vs), x1 = cond_load<T>(m + (long long)(2 * j + 1) * vs);
    __nv_bfloat16* d = dst + ((long long)b * h2 * w2 + (long long)i * w2 + j) * dst_ld + dst_off + pc * 4 + di * 2;
    *reinterpret_cast<uint32_t*>(d) = pack_bf16(x0, x1);
  }
}

}  // namespace tfx
