// Pipe-rate micro-benchmark for the softmax instruction mix (sm_100a): cycles per warp-instruction on one SMSP for
// MUFU.EX2, F2FP (bf16x2 pack), FFMA2, FADD2, FMNMX3, IMAD and for mixes of them, with 1 or 2 warps on the SMSP.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t ex2f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t packh(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

constexpr int R = 16;   // independent chains
constexpr int IT = 64;  // iterations

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float v[R];
  f32x2 w[R];
  uint32_t u[R];
#pragma unroll
  for (int i = 0; i < R; ++i) { v[i] = seed + i * 0.001f + threadIdx.x * 1e-4f; w[i] = ((f32x2)__float_as_uint(v[i]) << 32) | __float_as_uint(v[i]); u[i] = __float_as_uint(v[i]); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);
      if (MODE == 1) u[i] = pack(v[i], __uint_as_float(u[i]));
      if (MODE == 2) w[i] = fma2(w[i], w[i], w[i]);
      if (MODE == 3) w[i] = add2(w[i], w[i]);
      if (MODE == 4) v[i] = max3(v[i], v[(i + 1) % R], seed);
      if (MODE == 5) { v[i] = ex2(v[i]); u[i] = pack(v[i], __uint_as_float(u[i])); }              // 1 MUFU + 1 F2FP
      if (MODE == 6) { v[i] = ex2(v[i]); w[i] = fma2(w[i], w[i], w[i]); w[i] = add2(w[i], w[i]); }  // 1 MUFU + 2 packed
      if (MODE == 7) { v[i] = ex2(v[i]); u[i] = pack(v[i], __uint_as_float(u[i])); w[i] = fma2(w[i], w[i], w[i]); }
      if (MODE == 8) u[i] = ex2h2(u[i]);
      if (MODE == 9) u[i] = u[i] * 0x800000u + u[(i + 1) % R];
      if (MODE == 10) v[i] = fmaf(v[i], v[i], v[i]);
      if (MODE == 11) { u[i] = pack(v[i], __uint_as_float(u[i])); w[i] = fma2(w[i], w[i], w[i]); }
      if (MODE == 13) u[i] = ex2f16x2(u[i]);
      if (MODE == 14) { u[i] = packh(v[i], __uint_as_float(u[i])); u[i] = ex2f16x2(u[i]); }
      if (MODE == 15) u[i] = hadd2(u[i], u[(i + 1) % R]);
      if (MODE == 12) { v[i] = ex2(v[i]); v[i] = max3(v[i], v[(i + 1) % R], seed); }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < R; ++i) acc += v[i] + __uint_as_float(u[i]) + __uint_as_float((uint32_t)w[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {1, 2, 4, 8}) {  // warps per CTA; 4 SMSPs -> warps 1..4 land on distinct SMSPs, 8 = 2 per SMSP
    k<MODE><<<1, 32 * warps>>>(out, cyc, 0.5f);
    k<MODE><<<1, 32 * warps>>>(out, cyc, 0.5f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/CTA %d: %.2f cycles per warp-instr group (%d instr)\n", name, warps, double(c) / (IT * R), ops_per_iter);
  }
}

int main() {
  run<0>("MUFU.EX2", 1);
  run<1>("F2FP.BF16 pack", 1);
  run<2>("FFMA2", 1);
  run<3>("FADD2", 1);
  run<4>("FMNMX3", 1);
  run<10>("FFMA", 1);
  run<9>("IMAD shift-add", 1);
  run<8>("MUFU.EX2 bf16x2", 1);
  run<5>("MUFU + F2FP", 2);
  run<6>("MUFU + FFMA2 + FADD2", 3);
  run<7>("MUFU + F2FP + FFMA2", 3);
  run<11>("F2FP + FFMA2", 2);
  run<12>("MUFU + FMNMX3", 2);
  run<13>("MUFU.EX2 f16x2", 1);
  run<14>("F2FP.F16 + MUFU.EX2 f16x2", 2);
  run<15>("HADD2", 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
