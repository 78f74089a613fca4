// What does the softmax of the attention kernels cost a scheduler, with nothing else on the SM?
// W warps per scheduler (W = 1, 2, 3 warpgroups of 128 threads, one thread per row of a 128 x 128 score tile) run the tile body of
// attention_common.cuh -- tcgen05.ld of 128 score columns, row maximum, lazy reference, exponentials (a share on the FMA pipe),
// row sum, bf16 pack, tcgen05.st of P, the fences -- back to back on their own TMEM columns, no MMAs, no barriers.
// Variants by elimination: full | no maximum | no exponentials (load, max, store only) | no pack / store | emu 0 / 2 / 4.
// Prints cycles per warp and tile and per scheduler and tile (= cycles / W).  The MMAs of a tile take 1024 cycles.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/experiments/_build/softmax_rate tools/experiments/softmax_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../textflux_b200/csrc/attention_common.cuh"

using namespace tfx;

template <int kEmu, int kVariant>  // variant 0 full, 1 no max, 2 no exponentials, 3 no pack / store
__global__ void __launch_bounds__(384, 1) softmax_rate_kernel(int tiles, long long* out, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc<1>(&tmem_ptr, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const int wg = warp >> 2, quad = warp & 3;
  const uint32_t t_s = tmem + (uint32_t(quad * 32) << 16) + uint32_t(wg * 128);
  // something finite in the score columns
  {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(float((lane * 7 + i * 3) % 13) - 6.0f);
    for (int c = 0; c < 4; ++c) tmem_st32(t_s + c * 32, v);
    tmem_st_wait();
  }
  __syncthreads();
  const float c = 0.1275f;
  const f32x2 c2 = pack2(c, c);
  float m = -INFINITY, l = 0.f;
  const long long t0 = clock64();
  for (int j = 0; j < tiles; ++j) {
    uint32_t sr[4][32];
    attn_load_scores(t_s, 128, sr);
    float mx = 0.f;
    if (kVariant != 1) mx = attn_row_max(sr);
    const bool need = (mx - m) * c > kAttnRescaleThreshold;
    const float m_new = need ? mx : m;
    const float alpha = need ? ex2((m - m_new) * c) : 1.0f;
    const float mc = m_new * c;
    const f32x2 nmc2 = pack2(-mc, -mc);
    f32x2 sum2 = pack2(0.f, 0.f);
    uint32_t pk[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (kVariant == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) pk[i] = sr[2 * h][i] ^ sr[2 * h + 1][i];
      } else {
        attn_exp_half<kEmu>(sr[2 * h], sr[2 * h + 1], c2, nmc2, sum2, pk);
      }
      if (kVariant == 3) {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) x ^= pk[i];
        if (x == 0x12345678u) sink[0] = 1.f;
      } else {
        tmem_st32(t_s + h * 32, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
      }
    }
    float s0, s1;
    unpack2(sum2, s0, s1);
    l = l * alpha + s0 + s1;
    m = m_new;
    // put scores back for the next round (cheap: 16 of the columns) so that the data stay finite
    if (kVariant != 3) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = sr[0][i];
      tmem_st16(t_s, v);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 12 + warp] = t1 - t0;
  if (l == 12345.f) sink[1] = l;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

template <int kEmu, int kVariant>
void run(const char* name, int sms, long long* out, float* sink) {
  const int tiles = 2000;
  for (int w = 1; w <= 3; ++w) {
    for (int rep = 0; rep < 2; ++rep) {
      softmax_rate_kernel<kEmu, kVariant><<<sms, 128 * w>>>(tiles, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    double avg = 0;
    for (int b = 0; b < sms; ++b)
      for (int i = 0; i < 4 * w; ++i) avg += double(out[b * 12 + i]) / (sms * 4 * w);
    printf("%-34s %d warp(s) / scheduler: %7.1f cycles per warp and tile, %7.1f per scheduler and tile\n", name, w, avg / tiles, avg / tiles / w);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* out;
  float* sink;
  cudaMallocManaged(&out, sms * 12 * sizeof(long long));
  cudaMallocManaged(&sink, 16);
  run<2, 0>("full, emu 2", sms, out, sink);
  run<0, 0>("full, emu 0 (all MUFU)", sms, out, sink);
  run<4, 0>("full, emu 4 (half on FMA pipe)", sms, out, sink);
  run<2, 1>("no maximum, emu 2", sms, out, sink);
  run<2, 2>("no exponentials (ld, max, st)", sms, out, sink);
  run<2, 3>("no pack / store, emu 2", sms, out, sink);
  return 0;
}
