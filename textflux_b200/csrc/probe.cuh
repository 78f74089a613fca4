// Single-CTA tcgen05 descriptor probe: D[128, n] = A[128, k] * B^T with every descriptor field supplied by the host.
// Used during bring-up and kept as a regression test of the UMMA shared-memory / instruction descriptor encodings
// (K-major and MN-major B, A from shared memory or from TMEM) that gemm.cuh and attention.cuh rely on.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace tfx {

struct ProbeParams {
  const __nv_bfloat16* A;  // [128, k] row-major (only read directly when a_from_tmem)
  float* D;                // [128, n] row-major
  int n, k;
  int b_mn_major, a_from_tmem;
  uint32_t b_lbo, b_sbo, b_kstep_bytes;
};

// smem: A halves [k/64][128 rows][128 B]; then B: K-major -> [k/64][n rows][128 B]; MN-major -> [n/64][k rows][128 B]
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_halves = p.k / 64;
  uint8_t* sA = smem;
  uint8_t* sB = sA + a_halves * 128 * 128;
  const int b_boxes = p.b_mn_major ? p.n / 64 : p.k / 64;
  const int b_box_rows = p.b_mn_major ? p.k : p.n;
  const int b_box_bytes = b_box_rows * 128;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sB + b_boxes * b_box_bytes);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<1>(tmem_ptr, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, a_halves * 128 * 128 + b_boxes * b_box_bytes);
    for (int h = 0; h < a_halves; ++h) tma_load_2d(&tmA, bar_load, sA + h * 128 * 128, h * 64, 0, kEvictNormal);
    for (int h = 0; h < b_boxes; ++h) tma_load_2d(&tmB, bar_load, sB + h * b_box_bytes, h * 64, 0, kEvictNormal);
  }
  // A -> TMEM (packed bf16 pairs, row r in lane r, 2 elements per 32-bit column) at columns [256, 256 + k/2)
  if (p.a_from_tmem) {
    const int row = warp * 32 + lane;
    const uint32_t t_a = tmem + (uint32_t(warp * 32) << 16) + 256u;
    for (int c = 0; c < p.k / 32; ++c) {
      uint32_t pk[16];
      const uint4* src = reinterpret_cast<const uint4*>(p.A + row * p.k + c * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u = src[i];
        pk[4 * i] = u.x; pk[4 * i + 1] = u.y; pk[4 * i + 2] = u.z; pk[4 * i + 3] = u.w;
      }
      tmem_st16(t_a + c * 16, pk);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, p.n, 0, p.b_mn_major);
    for (int kk = 0; kk < p.k / 16; ++kk) {
      const uint32_t a_off = uint32_t((kk / 4) * 128 * 128 + (kk % 4) * 32);
      uint64_t db;
      if (p.b_mn_major) db = make_smem_desc(smem_u32(sB) + kk * p.b_kstep_bytes, p.b_lbo, p.b_sbo, kLayoutSW128);
      else db = make_smem_desc(smem_u32(sB) + uint32_t((kk / 4) * b_box_bytes + (kk % 4) * 32), p.b_lbo, p.b_sbo, kLayoutSW128);
      if (p.a_from_tmem) umma_ts(tmem, tmem + 256u + kk * 8, db, idesc, kk != 0);
      else umma_ss<1>(tmem, make_smem_desc(smem_u32(sA) + a_off, 16, 1024, kLayoutSW128), db, idesc, kk != 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < p.n / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.D[row * p.n + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

}  // namespace tfx
