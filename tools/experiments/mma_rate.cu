// How fast can one SM's tensor pipe run the MMA mix of the attention kernels when NOTHING else holds it back?
// One CTA per SM; one thread issues, per "KV tile", 8 SS MMAs 128 x 128 x 16 (Q K^T: A, B from shared memory) into a score buffer and
// 8 TS MMAs 128 x 128 x 16 (P V: A from TMEM, B MN-major from shared memory) into the accumulator -- the exact instructions, descriptors
// and TMEM layout of attention3 / attention9, operands left uninitialised (timing only), no TMA, no softmax, no barriers on the way.
// Modes: 0 QK only, 1 PV only, 2 QK then PV per tile (program order of the attention issuer), 3 QK / PV interleaved one by one,
//        +4: 8 other warps hammer TMEM with tcgen05.ld / st the way the softmax warps do (128 columns loaded, 64 stored per warp and tile).
// Prints cycles per tile and the share of the 16 x 64 = 1024 cycles the MMAs need on paper.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/experiments/_build/mma_rate tools/experiments/mma_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../textflux_b200/csrc/ptx.cuh"

using namespace tfx;

__global__ void __launch_bounds__(384, 1) mma_rate_kernel(int mode, int tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kTile = 128 * 128 * 2, kHalfBytes = 128 * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTile;          // 3 stages
  uint8_t* sV = sK + 3 * kTile;      // 2 stages
  uint64_t* bar = reinterpret_cast<uint64_t*>(sV + 2 * kTile);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(tmem_ptr + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    *stop = 0;
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const bool hammer = (mode & 4) != 0;
  mode &= 3;
  if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);
    const bool leader = elect_one();
    const uint64_t dQ = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
    const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 1024, kLayoutSW128);
    const uint64_t dV = make_smem_desc(smem_u32(sV), kHalfBytes, 1024, kLayoutSW128);
    constexpr uint32_t kTile16 = kTile / 16;
    const long long t0 = clock64();
    for (int j = 0; j < tiles; ++j) {
      const int buf = j % 3, vs = j % 2;
      const uint64_t bk = dK + uint64_t(buf * kTile16), bv = dV + uint64_t(vs * kTile16);
      const uint32_t d_s = tmem + uint32_t(buf * 128), d_o = tmem + 384u, a_p = tmem + uint32_t(((j + 1) % 3) * 128);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = uint32_t(((kk / 4) * kHalfBytes + (kk % 4) * 32) / 16);
          if (mode == 0 || mode == 2) umma_ss<1>(d_s, dQ + off, bk + off, idesc_qk, kk != 0);
          if (mode == 3) {
            umma_ss<1>(d_s, dQ + off, bk + off, idesc_qk, kk != 0);
            umma_ts(d_o, a_p + uint32_t(kk * 8), bv + uint64_t(kk * 128), idesc_pv, 1);
          }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (mode == 1 || mode == 2) umma_ts(d_o, a_p + uint32_t(kk * 8), bv + uint64_t(kk * 128), idesc_pv, 1);
      }
      __syncwarp();
    }
    if (leader) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    *stop = 1;
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp >= 4 && hammer) {
    // softmax-like TMEM traffic: per round 128 score columns in, 64 columns (packed P) out, on this warp's lane quadrant
    const uint32_t t_lane = tmem + (uint32_t((warp & 3) * 32) << 16);
    uint32_t v[32];
    long long rounds = 0;
    while (!*stop) {
      const uint32_t t_s = t_lane + uint32_t((rounds % 3) * 128);
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(t_s + c * 32, v);
        tmem_ld_wait();
      }
      for (int c = 0; c < 2; ++c) {
        tmem_st32(t_s + c * 32, v);
        tmem_st_wait();
      }
      ++rounds;
    }
    if (warp == 4 && lane == 0) out[gridDim.x + blockIdx.x] = rounds;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem, 512);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 6 * 32768 + 1024 + 256, tiles = 2000;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMallocManaged(&out, 2 * sms * sizeof(long long));
  const char* names[] = {"QK only (SS, 8 per tile)", "PV only (TS, 8 per tile)", "QK x 8 then PV x 8 per tile", "QK / PV alternating"};
  for (int grid : {1, sms})
    for (int mode = 0; mode < 8; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        for (int i = 0; i < 2 * sms; ++i) out[i] = 0;
        mma_rate_kernel<<<grid, 384, smem>>>(mode, tiles, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      double avg = 0;
      for (int i = 0; i < grid; ++i) avg += double(out[i]) / grid;
      const int per_tile = (mode & 3) >= 2 ? 16 : 8;
      printf("grid %3d  %-30s %s: %7.1f cycles / tile = %5.1f cycles / MMA (%4.1f %% of the 64-cycle floor)%s\n", grid, names[mode & 3],
             (mode & 4) ? "+ TMEM ld/st from 8 warps" : "                         ", avg / tiles, avg / tiles / per_tile,
             100.0 * 64 * per_tile / (avg / tiles), "");
      if (mode & 4) printf("          (hammer rounds per tile: %.2f)\n", double(out[grid]) / tiles);
    }
  return 0;
}
