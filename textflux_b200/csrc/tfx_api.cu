// C-ABI implementation (include/textflux_b200.h): model state, workspace, TMA descriptors, the per-step launch
// sequence of FluxTransformer2DModel.forward (transformer_flux.py:1028-1212) and CUDA-graph replay of it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/textflux_b200.h"
#include <algorithm>

#include "attention3.cuh"
#include "attention4.cuh"
#include "attention5.cuh"
#ifdef TFX_ATTN8
#include "../../tools/experiments/attention8.cuh"
#endif
#ifdef TFX_ATTN9
#include "../../tools/experiments/attention9.cuh"
#endif
#ifdef TFX_ATTN10
#include "../../tools/experiments/attention10.cuh"
#endif
#ifdef TFX_ATTN11
#include "../../tools/experiments/attention11.cuh"
#endif
#include "conditioning.cuh"
#include "gemm.cuh"
#include "pointwise.cuh"
#include "probe.cuh"
#include "textenc.cuh"
#include "vae.cuh"

using namespace tfx;
typedef __nv_bfloat16 bf16;

namespace {

thread_local std::string g_last_error;

// NVTX range over a host-side scope (header-only nvtx3: a no-op unless a profiler is attached).  Every C-ABI entry point of the
// hot path and the sections of the launch sequence carry one, so a timeline shows call -> section -> kernels.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Fail {
  int code;
};

int set_error(std::string* dst, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (dst) *dst = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                    \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      set_error(err_, TFX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw Fail{TFX_ERR_CUDA};                                                                           \
    }                                                                                                     \
  } while (0)

#define REQUIRE(cond, code, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      set_error(err_, code, __VA_ARGS__);        \
      throw Fail{code};                          \
    }                                            \
  } while (0)

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn(std::string* err_) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
  REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, TFX_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows x 64 cols], 128B swizzle.
CUtensorMap make_map_2d(std::string* err_, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  CUtensorMap m;
  memset(&m, 0, sizeof m);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, TFX_ERR_INVALID,
          "TMA operand must be 16-byte aligned (ptr %p, ld %lld)", ptr, ld);
  CUresult r = get_encode_fn(err_)(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, TFX_ERR_CUDA, "cuTensorMapEncodeTiled(2d %lldx%lld ld %lld box %d) failed: %d", rows, cols, ld,
          box_rows, (int)r);
  return m;
}

// 3-D bf16 [outer, rows, cols] contiguous; box = [1, box_rows rows, 64 cols]
CUtensorMap make_map_3d(std::string* err_, const void* ptr, long long outer, long long rows, long long cols, int box_rows = 128) {
  CUtensorMap m;
  memset(&m, 0, sizeof m);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn(err_)(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, TFX_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return m;
}

int num_sms(int device) {
  static std::map<int, int> cache;
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  cache[device] = n;
  return n;
}

// ------------------------------------------------------------------------------------------------ launchers
// Optional per-launch device timing (eager mode only): one event pair per launch, summed per kernel family.
enum KernelFamily : int { KF_GEMM = 0, KF_ATTN = 1, KF_LN = 2, KF_GEMV = 3, KF_MISC = 4, KF_COUNT = 5 };
struct Profiler {
  struct Rec { int fam; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  void begin(int fam, cudaStream_t s) {
    Rec r; r.fam = fam;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    recs.push_back(r);
  }
  void end(cudaStream_t s) { cudaEventRecord(recs.back().b, s); }
  // returns microseconds and launch counts per family; destroys the events
  void collect(double (&us)[KF_COUNT], long long (&n)[KF_COUNT]) {
    for (int i = 0; i < KF_COUNT; ++i) { us[i] = 0; n[i] = 0; }
    for (auto& r : recs) {
      cudaEventSynchronize(r.b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      us[r.fam] += ms * 1000.0; n[r.fam] += 1;
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    recs.clear();
  }
};

struct LaunchCtx {
  cudaStream_t stream;
  int device;
  long long* counter;
  std::string* err_;
  Profiler* prof = nullptr;
  bool pdl = false;  // programmatic dependent launch: the kernel may start while its predecessor drains
};

// <<<>>> with optional launch attributes (PDL, cluster)
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, const LaunchCtx& c, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (c.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c.stream;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
struct ProfScope {
  const LaunchCtx& c;
  ProfScope(const LaunchCtx& c_, int fam) : c(c_) { if (c.prof) c.prof->begin(fam, c.stream); }
  ~ProfScope() { if (c.prof) c.prof->end(c.stream); }
};

// cudaFuncSetAttribute is per device: run once for every device a handle (or an op-level call) touches
void configure_kernels(std::string* err_) {
  static std::map<int, bool> done_on;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  bool& done = done_on[dev];
  if (done) return;
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1, 256>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<2, 256>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 224>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1, 224>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 224>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<2, 224>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1, 192>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<2, 192>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1, 128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<2, 128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1, 64>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<2, 64>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_mc_tcgen05_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmMcCfg<256, 1>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_mc_tcgen05_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmMcCfg<256, 2>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_mc_tcgen05_kernel<224, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmMcCfg<224, 2>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_mc_tcgen05_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmMcCfg<256, 4>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(gemm_mc_tcgen05_kernel<224, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmMcCfg<224, 4>::kSmemBytes));
#define TFX_ATTN3_ATTR(DH, EMU, SPLIT, TRACE) \
  CUDA_TRY(cudaFuncSetAttribute(attention3_tcgen05_kernel<DH, EMU, SPLIT, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg<DH>::kSmemBytes))
  TFX_ATTN3_ATTR(128, 0, false, false); TFX_ATTN3_ATTR(128, 2, false, false);
  TFX_ATTN3_ATTR(128, 0, true, false); TFX_ATTN3_ATTR(128, 2, true, false); TFX_ATTN3_ATTR(128, 3, true, false); TFX_ATTN3_ATTR(128, 4, true, false);
  TFX_ATTN3_ATTR(128, 2, true, true);
  TFX_ATTN3_ATTR(64, 0, true, false); TFX_ATTN3_ATTR(64, 2, true, false);
#undef TFX_ATTN3_ATTR
#define TFX_ATTN4_ATTR(DH, EMU) \
  CUDA_TRY(cudaFuncSetAttribute(attention4_tcgen05_kernel<DH, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg<DH>::kSmemBytes))
  TFX_ATTN4_ATTR(128, 0); TFX_ATTN4_ATTR(128, 2); TFX_ATTN4_ATTR(128, 3); TFX_ATTN4_ATTR(128, 4); TFX_ATTN4_ATTR(64, 0); TFX_ATTN4_ATTR(64, 2);
#undef TFX_ATTN4_ATTR
#define TFX_ATTN5_ATTR(DH, EMU) \
  CUDA_TRY(cudaFuncSetAttribute(attention5_tcgen05_kernel<DH, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg<DH>::kSmemBytes))
  CUDA_TRY(cudaFuncSetAttribute(attention5_tcgen05_kernel<128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg<128>::kSmemBytes));
  TFX_ATTN5_ATTR(128, 0); TFX_ATTN5_ATTR(128, 2); TFX_ATTN5_ATTR(128, 3); TFX_ATTN5_ATTR(128, 4); TFX_ATTN5_ATTR(64, 0); TFX_ATTN5_ATTR(64, 2);
#undef TFX_ATTN5_ATTR
#ifdef TFX_ATTN8
  CUDA_TRY(cudaFuncSetAttribute(attention8_tcgen05_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn8Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention8_tcgen05_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn8Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention8_tcgen05_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn8Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention8_tcgen05_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn8Cfg<64>::kSmemBytes));
#endif
#ifdef TFX_ATTN9
  CUDA_TRY(cudaFuncSetAttribute(attention9_tcgen05_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn9Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention9_tcgen05_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn9Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention9_tcgen05_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn9Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention9_tcgen05_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn9Cfg<64>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention9_tcgen05_kernel<128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn9Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention9_tcgen05_kernel<128, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn9Cfg<128>::kSmemBytes));
#endif
#ifdef TFX_ATTN10
  CUDA_TRY(cudaFuncSetAttribute(attention10_tcgen05_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn10Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention10_tcgen05_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn10Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention10_tcgen05_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn10Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention10_tcgen05_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn10Cfg<64>::kSmemBytes));
#endif
#ifdef TFX_ATTN11
  CUDA_TRY(cudaFuncSetAttribute(attention11_tcgen05_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn11Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention11_tcgen05_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn11Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention11_tcgen05_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn11Cfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(attention11_tcgen05_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn11Cfg<128>::kSmemBytes));
#endif
  CUDA_TRY(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  done = true;
}

// Tile width for an [M, N] output on `units` concurrent tiles (SMs or SM pairs): minimise waves x width, where a narrower
// tile must win by more than its lower operand reuse costs.  Measured (profiles/r1h_kernels.json, r1i ncu): 224-wide
// tiles are ~5 % less efficient per column than 256-wide ones, 192-wide ones 15 % (never worth it on these shapes).
int pick_block_n(long long m_tiles, int N, int units, bool allow_narrow) {
  const long long t256 = m_tiles * ((N + 255) / 256), t224 = m_tiles * ((N + 223) / 224);
  const double c256 = double((t256 + units - 1) / units) * 256.0;
  const double c224 = double((t224 + units - 1) / units) * 224.0 * 1.05;
  return (allow_narrow && c224 < c256) ? 224 : 256;
}

// Operands of the extension k-block (GemmParams::k_ext): A side [M, 64] per group, W side [N, 64] per group.
struct GemmExt {
  const CUtensorMap *e0, *e1, *f0, *f1;
};

template <int kCG, int kBN>
void launch_gemm_inst(const LaunchCtx& c, long long tiles, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                      const CUtensorMap& b1, const GemmExt* x, const GemmParams& p) {
  std::string* err_ = c.err_;
  const int sms = num_sms(c.device);
  long long ctas = (kCG == 2) ? 2 * (tiles < sms / 2 ? tiles : sms / 2) : (tiles < sms ? tiles : sms);
  CUDA_TRY(launch_ex(gemm_tcgen05_kernel<kCG, kBN>, dim3((unsigned)ctas), dim3(kGemmThreads), GemmCfg<kCG, kBN>::kSmemBytes, c, kCG,
                     a0, a1, b0, b1, x ? *x->e0 : a0, x ? *x->e1 : a1, x ? *x->f0 : b0, x ? *x->f1 : b1, p));
}

// b0/b1 must be descriptors whose box holds block_n / cta_group rows.
void launch_gemm(const LaunchCtx& c, int cta_group, int block_n, const CUtensorMap& a0, const CUtensorMap& a1,
                 const CUtensorMap& b0, const CUtensorMap& b1, const GemmParams& p, const GemmExt* ext = nullptr) {
  std::string* err_ = c.err_;
  // K need not fill its last 64-wide k-block: TMA zero-fills both operands past K (the descriptors check the 16-byte row stride)
  REQUIRE(p.K > 0, TFX_ERR_INVALID, "GEMM K=%d must be positive", p.K);
  REQUIRE(p.k_ext == 0 || (p.k_ext == kGemmBlockK && ext), TFX_ERR_INVALID, "k_ext %d needs extension descriptors", p.k_ext);
  REQUIRE(p.n_split == p.N || p.n_split % block_n == 0, TFX_ERR_INVALID, "n_split %d not tile aligned", p.n_split);
  const bool qkv = p.mode0 == EPI_QKV || (p.n_split < p.N && p.mode1 == EPI_QKV);
  REQUIRE(!qkv || block_n == 256, TFX_ERR_INVALID, "QKV epilogue needs 256-wide tiles");
  ProfScope ps(c, KF_GEMM);
  const int tile_m = 128 * cta_group;
  long long tiles = 0;
  for (int g = 0; g < p.num_groups; ++g) tiles += (p.g[g].M + tile_m - 1) / tile_m;
  if (p.conv.mode) tiles = (p.conv.n_patches + cta_group - 1) / cta_group;  // convolution: an M tile is one 16 x 8 pixel patch per CTA
  tiles *= (p.N + block_n - 1) / block_n;
  if (tiles == 0) return;
  const int key = cta_group * 1000 + block_n;
  switch (key) {
    case 1256: launch_gemm_inst<1, 256>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 2256: launch_gemm_inst<2, 256>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 1224: launch_gemm_inst<1, 224>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 2224: launch_gemm_inst<2, 224>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 1192: launch_gemm_inst<1, 192>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 2192: launch_gemm_inst<2, 192>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 1128: launch_gemm_inst<1, 128>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 2128: launch_gemm_inst<2, 128>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 1064: launch_gemm_inst<1, 64>(c, tiles, a0, a1, b0, b1, ext, p); break;
    case 2064: launch_gemm_inst<2, 64>(c, tiles, a0, a1, b0, b1, ext, p); break;
    default: REQUIRE(false, TFX_ERR_INVALID, "no GEMM instance for cta_group %d block_n %d", cta_group, block_n);
  }
  ++*c.counter;
}

// ---- multicast GEMM: clusters of 2*kPN CTAs (kPN pairs share their A rows)
template <int kBN, int kPN>
int max_mc_clusters(int device) {
  static std::map<int, int> cache;
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kGemmThreads);
  cfg.gridDim = dim3(2 * kPN * 64);
  cfg.dynamicSmemBytes = GemmMcCfg<kBN, kPN>::kSmemBytes;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * kPN;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_mc_tcgen05_kernel<kBN, kPN>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = num_sms(device) / (2 * kPN) * 3 / 4;  // conservative fallback
  }
  if (const char* e = getenv("TFX_MC_CLUSTERS")) n = atoi(e);
  if (getenv("TFX_DEBUG")) fprintf(stderr, "tfx: gemm_mc<%d,%d> max active clusters %d\n", kBN, kPN, n);
  cache[device] = n;
  return n;
}

template <int kBN, int kPN>
void launch_gemm_mc_inst(const LaunchCtx& c, long long m_tiles, const CUtensorMap& a0, const CUtensorMap& a1,
                         const CUtensorMap& b0, const CUtensorMap& b1, const GemmParams& p) {
  std::string* err_ = c.err_;
  const long long nt = (p.N + kBN - 1) / kBN;
  const long long supers = m_tiles * ((nt + kPN - 1) / kPN);
  long long clusters = max_mc_clusters<kBN, kPN>(c.device);
  if (supers < clusters) clusters = supers;
  CUDA_TRY(launch_ex(gemm_mc_tcgen05_kernel<kBN, kPN>, dim3((unsigned)(clusters * 2 * kPN)), dim3(kGemmThreads),
                     GemmMcCfg<kBN, kPN>::kSmemBytes, c, 2 * kPN, a0, a1, b0, b1, p));
}

// a0/a1: descriptors with 128/pn-row boxes; b0/b1: block_n/2-row boxes.
void launch_gemm_mc(const LaunchCtx& c, int pn, int block_n, const CUtensorMap& a0, const CUtensorMap& a1,
                    const CUtensorMap& b0, const CUtensorMap& b1, const GemmParams& p) {
  std::string* err_ = c.err_;
  REQUIRE(p.K % kGemmBlockK == 0 && p.K > 0, TFX_ERR_INVALID, "GEMM K=%d must be a positive multiple of %d", p.K, kGemmBlockK);
  REQUIRE(p.n_split == p.N || p.n_split % block_n == 0, TFX_ERR_INVALID, "n_split %d not tile aligned", p.n_split);
  const bool qkv = p.mode0 == EPI_QKV || (p.n_split < p.N && p.mode1 == EPI_QKV);
  REQUIRE(!qkv || block_n == 256, TFX_ERR_INVALID, "QKV epilogue needs 256-wide tiles");
  ProfScope ps(c, KF_GEMM);
  long long m_tiles = 0;
  for (int g = 0; g < p.num_groups; ++g) m_tiles += (p.g[g].M + 255) / 256;
  if (m_tiles == 0) return;
  const int key = pn * 1000 + block_n;
  switch (key) {
    case 1256: launch_gemm_mc_inst<256, 1>(c, m_tiles, a0, a1, b0, b1, p); break;
    case 2256: launch_gemm_mc_inst<256, 2>(c, m_tiles, a0, a1, b0, b1, p); break;
    case 2224: launch_gemm_mc_inst<224, 2>(c, m_tiles, a0, a1, b0, b1, p); break;
    case 4256: launch_gemm_mc_inst<256, 4>(c, m_tiles, a0, a1, b0, b1, p); break;
    case 4224: launch_gemm_mc_inst<224, 4>(c, m_tiles, a0, a1, b0, b1, p); break;
    default: REQUIRE(false, TFX_ERR_INVALID, "no multicast GEMM instance for pn %d block_n %d", pn, block_n);
  }
  ++*c.counter;
}

// Schedule 3 (attention3.cuh): 2 query tiles per CTA, one CTA per (batch, head, 256 query rows).
// emu: exponentials per 8 on the FMA pipe (0, 2, 3, 4); split: hand P over in two halves; trace: clock stamps of CTA 0
void launch_attention3(const LaunchCtx& c, int head_dim, int emu, bool split, bool trace, const CUtensorMap& tq,
                       const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p) {
  std::string* err_ = c.err_;
  REQUIRE(head_dim == 64 || head_dim == 128, TFX_ERR_INVALID, "attention_head_dim %d unsupported (64 or 128)", head_dim);
  ProfScope ps(c, KF_ATTN);
  dim3 grid((p.N + 255) / 256, p.H, p.B);
#define TFX_ATTN3(DH, EMU, SPLIT, TRACE) \
  CUDA_TRY(launch_ex(attention3_tcgen05_kernel<DH, EMU, SPLIT, TRACE>, grid, dim3(Attn3Cfg<DH>::kThreads), Attn3Cfg<DH>::kSmemBytes, c, 1, tq, tk, tv, p))
  if (head_dim == 128 && trace) {
    TFX_ATTN3(128, 2, true, true);
  } else if (head_dim == 128 && split) {
    switch (emu) {
      case 1:
      case 2: TFX_ATTN3(128, 2, true, false); break;
      case 3: TFX_ATTN3(128, 3, true, false); break;
      case 4: TFX_ATTN3(128, 4, true, false); break;
      default: TFX_ATTN3(128, 0, true, false); break;
    }
  } else if (head_dim == 128) {
    if (emu) TFX_ATTN3(128, 2, false, false); else TFX_ATTN3(128, 0, false, false);
  } else {
    if (emu) TFX_ATTN3(64, 2, true, false); else TFX_ATTN3(64, 0, true, false);
  }
#undef TFX_ATTN3
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}

// Schedule 4 (attention4.cuh): schedule 3 with the units of the last, partial wave cut into equal KV shares.
struct Attn4Workspace {
  float* ws_o = nullptr;
  float* ws_ml = nullptr;
  int* counters = nullptr;
  static size_t o_bytes(int head_dim) { return (size_t)kAttn4MaxShares * 2 * 2 * head_dim * 128 * sizeof(float); }
  static size_t ml_bytes() { return (size_t)kAttn4MaxShares * 2 * 2 * 2 * 128 * sizeof(float); }
  static size_t counter_bytes() { return (size_t)kAttn4MaxShares * sizeof(int); }
};

// Fills the decomposition fields of `pp` for B*H*ceil(N/256) units on `sms` SMs (see attention4.cuh).
void plan_attention4(Attn4Params& pp, int sms) {
  const AttnParams& a = pp.a;
  const int n_kv = (a.N + 127) / 128;
  if (sms > kAttn4MaxShares) sms = kAttn4MaxShares;
  pp.n_qpairs = (a.N + 255) / 256;
  pp.n_units = a.B * a.H * pp.n_qpairs;
  pp.n_full = pp.n_units / sms * sms;
  pp.n_rem = pp.n_units - pp.n_full;
  pp.stream_ctas = 0; pp.share = n_kv; pp.n_seg2 = 0;
  if (pp.n_rem == 0) return;
  const long long total = (long long)pp.n_rem * n_kv;
  int share = (int)((total + sms - 1) / sms);
  // a partial wave that is nearly full gains nothing from being cut (the merges cost about two iterations)
  if (share * 20 > n_kv * 17) { pp.n_full = pp.n_units; pp.n_rem = 0; return; }
  share = std::max(share, std::min(4, n_kv));                               // do not cut below 4 iterations per CTA
  share = std::max(share, (n_kv + kAttn4MaxParts - 2) / (kAttn4MaxParts - 1));  // at most kAttn4MaxParts parts per unit
  share = std::min(share, n_kv);
  pp.share = share;
  pp.stream_ctas = (int)((total + share - 1) / share);
  // shares that cross a unit boundary get a second CTA; longest second segment first (see attention4.cuh)
  std::vector<std::pair<int, int>> seg2;  // (-length, share)
  for (int c = 0; c < pp.stream_ctas; ++c) {
    const long long start = (long long)c * share, end = std::min(start + share, total);
    const long long u0 = start / n_kv, u1 = (end - 1) / n_kv;
    if (u1 != u0) seg2.push_back({-(int)(end - u1 * n_kv), c});
  }
  std::sort(seg2.begin(), seg2.end());
  pp.n_seg2 = (int)seg2.size();
  for (int i = 0; i < pp.n_seg2; ++i) pp.seg2_share[i] = (uint16_t)seg2[i].second;
}

void launch_attention4(const LaunchCtx& c, int head_dim, int emu, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnParams& p, const Attn4Workspace& ws) {
  std::string* err_ = c.err_;
  REQUIRE(head_dim == 64 || head_dim == 128, TFX_ERR_INVALID, "attention_head_dim %d unsupported (64 or 128)", head_dim);
  REQUIRE(ws.ws_o && ws.ws_ml && ws.counters, TFX_ERR_STATE, "attention workspace missing");
  ProfScope ps(c, KF_ATTN);
  Attn4Params pp;
  memset(&pp, 0, sizeof pp);
  pp.a = p;
  pp.ws_o = ws.ws_o; pp.ws_ml = ws.ws_ml; pp.counters = ws.counters;
  plan_attention4(pp, num_sms(c.device));
  dim3 grid(pp.n_full + pp.stream_ctas + pp.n_seg2);
#define TFX_ATTN4(DH, EMU) \
  CUDA_TRY(launch_ex(attention4_tcgen05_kernel<DH, EMU>, grid, dim3(Attn3Cfg<DH>::kThreads), Attn3Cfg<DH>::kSmemBytes, c, 1, tq, tk, tv, pp))
  if (head_dim == 128) {
    switch (emu) {
      case 1:
      case 2: TFX_ATTN4(128, 2); break;
      case 3: TFX_ATTN4(128, 3); break;
      case 4: TFX_ATTN4(128, 4); break;
      default: TFX_ATTN4(128, 0); break;
    }
  } else {
    if (emu) TFX_ATTN4(64, 2); else TFX_ATTN4(64, 0);
  }
#undef TFX_ATTN4
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}

// Schedule 5 (attention5.cuh): one persistent CTA per SM walking whole units, then its share of the remainder stream.
void plan_attention5(Attn5Params& pp, int sms) {
  const AttnParams& a = pp.a;
  const int n_kv = (a.N + 127) / 128;
  const int G = std::min(sms, kAttn4MaxShares);
  pp.n_qpairs = (a.N + 255) / 256;
  pp.n_units = a.B * a.H * pp.n_qpairs;
  pp.n_waves = pp.n_units / G;
  pp.n_rem = pp.n_units - pp.n_waves * G;
  pp.share = 0; pp.stream_ctas = 0;
  if (pp.n_rem > 0) {
    const long long total = (long long)pp.n_rem * n_kv;
    int share = (int)((total + G - 1) / G);
    if (share * 20 > n_kv * 17) share = n_kv;                                   // nearly full partial wave: whole units, no merges
    share = std::max(share, std::min(4, n_kv));                                   // do not cut below 4 iterations
    share = std::max(share, (n_kv + kAttn4MaxParts - 2) / (kAttn4MaxParts - 1));  // at most kAttn4MaxParts parts per unit
    share = std::min(share, n_kv);
    pp.share = share;
    pp.stream_ctas = (int)((total + share - 1) / share);
  }
  pp.grid = pp.n_waves > 0 ? G : std::max(pp.stream_ctas, 1);
}

void launch_attention5(const LaunchCtx& c, int head_dim, int emu, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnParams& p, const Attn4Workspace& ws, bool trace = false) {
  std::string* err_ = c.err_;
  REQUIRE(head_dim == 64 || head_dim == 128, TFX_ERR_INVALID, "attention_head_dim %d unsupported (64 or 128)", head_dim);
  REQUIRE(ws.ws_o && ws.ws_ml && ws.counters, TFX_ERR_STATE, "attention workspace missing");
  ProfScope ps(c, KF_ATTN);
  Attn5Params pp;
  memset(&pp, 0, sizeof pp);
  pp.a = p;
  pp.ws_o = ws.ws_o; pp.ws_ml = ws.ws_ml; pp.counters = ws.counters;
  plan_attention5(pp, num_sms(c.device));
  dim3 grid(pp.grid);
#define TFX_ATTN5(DH, EMU) \
  CUDA_TRY(launch_ex(attention5_tcgen05_kernel<DH, EMU>, grid, dim3(Attn3Cfg<DH>::kThreads), Attn3Cfg<DH>::kSmemBytes, c, 1, tq, tk, tv, pp))
  if (head_dim == 128 && trace) {
    CUDA_TRY(launch_ex(attention5_tcgen05_kernel<128, 2, true>, grid, dim3(Attn3Cfg<128>::kThreads), Attn3Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, pp));
  } else if (head_dim == 128) {
    switch (emu) {
      case 1:
      case 2: TFX_ATTN5(128, 2); break;
      case 3: TFX_ATTN5(128, 3); break;
      case 4: TFX_ATTN5(128, 4); break;
      default: TFX_ATTN5(128, 0); break;
    }
  } else {
    if (emu) TFX_ATTN5(64, 2); else TFX_ATTN5(64, 0);
  }
#undef TFX_ATTN5
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}

#ifdef TFX_ATTN8
void launch_attention8(const LaunchCtx& c, int head_dim, int emu, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnParams& p) {
  std::string* err_ = c.err_;
  ProfScope ps(c, KF_ATTN);
  dim3 grid((p.N + 255) / 256, p.H, p.B);
  if (head_dim == 64) {
    CUDA_TRY(launch_ex(attention8_tcgen05_kernel<64, 2>, grid, dim3(Attn8Cfg<64>::kThreads), Attn8Cfg<64>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 0) {
    CUDA_TRY(launch_ex(attention8_tcgen05_kernel<128, 0>, grid, dim3(Attn8Cfg<128>::kThreads), Attn8Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 3) {
    CUDA_TRY(launch_ex(attention8_tcgen05_kernel<128, 3>, grid, dim3(Attn8Cfg<128>::kThreads), Attn8Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else {
    CUDA_TRY(launch_ex(attention8_tcgen05_kernel<128, 2>, grid, dim3(Attn8Cfg<128>::kThreads), Attn8Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  }
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}
#endif

#ifdef TFX_ATTN9
void launch_attention9(const LaunchCtx& c, int head_dim, int emu, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnParams& p) {
  std::string* err_ = c.err_;
  ProfScope ps(c, KF_ATTN);
  dim3 grid((p.N + 127) / 128, p.H, p.B);
  if (head_dim == 64) {
    CUDA_TRY(launch_ex(attention9_tcgen05_kernel<64, 2>, grid, dim3(Attn9Cfg<64>::kThreads), Attn9Cfg<64>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 0) {
    CUDA_TRY(launch_ex(attention9_tcgen05_kernel<128, 0>, grid, dim3(Attn9Cfg<128>::kThreads), Attn9Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 1 && p.trace) {
    CUDA_TRY(launch_ex(attention9_tcgen05_kernel<128, 2, true, true>, grid, dim3(Attn9Cfg<128>::kThreads), Attn9Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 1) {  // experiment: emu code 1 = split-issue variant at emu 2
    CUDA_TRY(launch_ex(attention9_tcgen05_kernel<128, 2, true>, grid, dim3(Attn9Cfg<128>::kThreads), Attn9Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 3) {
    CUDA_TRY(launch_ex(attention9_tcgen05_kernel<128, 3>, grid, dim3(Attn9Cfg<128>::kThreads), Attn9Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else {
    CUDA_TRY(launch_ex(attention9_tcgen05_kernel<128, 2>, grid, dim3(Attn9Cfg<128>::kThreads), Attn9Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  }
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}
#endif

#ifdef TFX_ATTN10
void launch_attention10(const LaunchCtx& c, int head_dim, int emu, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const AttnParams& p) {
  std::string* err_ = c.err_;
  ProfScope ps(c, KF_ATTN);
  dim3 grid((p.N + 127) / 128, p.H, p.B);
  if (head_dim == 64) {
    CUDA_TRY(launch_ex(attention10_tcgen05_kernel<64, 2>, grid, dim3(Attn10Cfg<64>::kThreads), Attn10Cfg<64>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 0) {
    CUDA_TRY(launch_ex(attention10_tcgen05_kernel<128, 0>, grid, dim3(Attn10Cfg<128>::kThreads), Attn10Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else if (emu == 3) {
    CUDA_TRY(launch_ex(attention10_tcgen05_kernel<128, 3>, grid, dim3(Attn10Cfg<128>::kThreads), Attn10Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  } else {
    CUDA_TRY(launch_ex(attention10_tcgen05_kernel<128, 2>, grid, dim3(Attn10Cfg<128>::kThreads), Attn10Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p));
  }
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}
#endif

#ifdef TFX_ATTN11
void launch_attention11(const LaunchCtx& c, int head_dim, int emu, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const AttnParams& p) {
  std::string* err_ = c.err_;
#ifdef TFX_ATTN9
  if (head_dim != 128) return launch_attention9(c, head_dim, emu, tq, tk, tv, p);
#endif
  REQUIRE(head_dim == 128, TFX_ERR_INVALID, "schedule 11 (experiment) is built for head_dim 128 only");
  ProfScope ps(c, KF_ATTN);
  dim3 grid((p.N + 127) / 128, p.H, p.B);
#define TFX_ATTN11_L(EMU) \
  CUDA_TRY(launch_ex(attention11_tcgen05_kernel<128, EMU>, grid, dim3(Attn11Cfg<128>::kThreads), Attn11Cfg<128>::kSmemBytes, c, 1, tq, tk, tv, p))
  if (emu == 0) TFX_ATTN11_L(0);
  else if (emu == 3) TFX_ATTN11_L(3);
  else if (emu == 4) TFX_ATTN11_L(4);
  else TFX_ATTN11_L(2);
#undef TFX_ATTN11_L
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}
#endif

// attn_variant 0 ("auto"): schedule 5 where its work decomposition pays, schedule 3 elsewhere.  Measured (tools/bench_kernels.py,
// profiles/r2e_kernels_attn.json, 24 heads, TFLOP/s, schedule 3 / schedule 5): N = 2560 936 / 865, 4608 1287 / 1249, 5120 1131 /
// 1229, 8704 1321 / 1319, 12800 1313 / 1274 -- schedule 5 wins when every SM runs at least two whole units before the shared
// remainder and the remainder is cut into few parts (each part costs a 5 us park, each cut unit a 6-11 us merge).
int pick_attention_variant(const AttnParams& a, int sms) {
  Attn5Params pp;
  memset(&pp, 0, sizeof pp);
  pp.a = a;
  plan_attention5(pp, sms);
  const int n_kv = (a.N + 127) / 128;
  if (pp.n_waves < 2 || pp.n_rem == 0 || pp.share >= n_kv) return 5;
  const int parts = (n_kv + pp.share - 1) / pp.share + 1;
  return parts <= 5 ? 9 : 5;
}

void launch_ln_modulate(const LaunchCtx& c, const LnModParams& p) {
  std::string* err_ = c.err_;
  REQUIRE(p.D % 256 == 0, TFX_ERR_INVALID, "LayerNorm width %d must be a multiple of 256", p.D);
  const int rows = p.rows - p.row_begin;
  if (rows <= 0) return;
  ProfScope ps(c, KF_LN);
  const int blocks = (rows + 3) / 4;
  switch (p.D / 256) {
    case 1: CUDA_TRY(launch_ex(ln_modulate_kernel<1>, dim3(blocks), dim3(128), 0, c, 1, p)); break;
    case 2: CUDA_TRY(launch_ex(ln_modulate_kernel<2>, dim3(blocks), dim3(128), 0, c, 1, p)); break;
    case 4: CUDA_TRY(launch_ex(ln_modulate_kernel<4>, dim3(blocks), dim3(128), 0, c, 1, p)); break;
    case 8: CUDA_TRY(launch_ex(ln_modulate_kernel<8>, dim3(blocks), dim3(128), 0, c, 1, p)); break;
    case 12: CUDA_TRY(launch_ex(ln_modulate_kernel<12>, dim3(blocks), dim3(128), 0, c, 1, p)); break;
    case 16: CUDA_TRY(launch_ex(ln_modulate_kernel<16>, dim3(blocks), dim3(128), 0, c, 1, p)); break;
    default: REQUIRE(false, TFX_ERR_INVALID, "LayerNorm width %d unsupported", p.D);
  }
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}

void launch_gemv(const LaunchCtx& c, const bf16* x, int B, int K, const bf16* W, const bf16* bias, long long N, bf16* out, int flags,
                 const int* skip = nullptr) {
  std::string* err_ = c.err_;
  REQUIRE(K % 8 == 0, TFX_ERR_INVALID, "GEMV K=%d must be a multiple of 8", K);
  REQUIRE(B >= 1 && B <= kGemvMaxB, TFX_ERR_INVALID, "batch %d unsupported (1..%d)", B, kGemvMaxB);
  const size_t smem = (size_t)B * K * 2;
  REQUIRE(smem <= 48 * 1024, TFX_ERR_INVALID, "GEMV input %zu bytes exceeds 48 KiB", smem);
  ProfScope ps(c, KF_GEMV);
  long long blocks = (N + 7) / 8;
  const long long cap = (long long)num_sms(c.device) * 8;
  if (blocks > cap) blocks = cap;
  gemv_kernel<<<(unsigned)blocks, 256, smem, c.stream>>>(x, B, K, W, bias, N, out, flags, skip);
  CUDA_TRY(cudaGetLastError());
  ++*c.counter;
}

struct Weight {
  const bf16* ptr = nullptr;
  long long rows = 0, cols = 0;
};

}  // namespace

// ==================================================================================================== model
struct tfx_model {
  tfx_config cfg;
  int device = 0;
  std::string err;
  std::string* err_ = &err;
  long long launches = 0;
  long long graph_nodes = 0;
  int gemm_cta_group = 2;  // 2-CTA 256x256 tiles: the configuration every published number was measured with
  int gemm_mcast = 0;  // 0: plain kernels; 2|4: pairs per cluster sharing A by TMA multicast
  int attn_variant = 0;  // 0: per shape, 5 or 9 (pick_attention_variant); 4, 5: schedule 3 (attention3.cuh) whole-P / split-P hand-over, one CTA per (head, 256 query rows);
                         // 8: schedule 4 (attention4.cuh): schedule 3 split-P + the last partial wave cut into KV shares;
                         // 9: schedule 5 (attention5.cuh): persistent CTAs, items overlapped, remainder cut into KV shares
  int gemm_narrow_tiles = 1;  // allow 224-wide tiles where they cut wave quantisation (option "gemm_narrow_tiles")
  int gemm_k_snake = 0;  // GemmParams::k_snake on the banded wide-K GEMMs: -3 .. -6 GB of DRAM traffic per cfg3 step, -0.3 .. -0.5 % step time, but a
                         // row's fp32 summation order then depends on where its tile sits, i.e. on the batch it is in -- off by default (profiles/r2h_band.md)
  int gemm_m_band = -1;  // tile order of the wide-K GEMMs (ff down, single proj_out), whose A operand outgrows the L2 at N >= 4608:
                         // 0 = M-fastest over all M tiles, b > 0 = bands of b M tiles (GemmParams::m_band), -1 = per shape (band_for)
  int gemm_l2_hints = 0;  // bit 0: A (activation) loads evict_last, bit 1: B (weight) loads evict_first (option "gemm_l2_hints")
  int attn_emu = 2;  // column pairs per 8 whose exponentials run on the FMA pipe (packed polynomial) instead of MUFU:
                     // 2 measured best (+4..8 %), 0 = all MUFU
  int mod_cache_slots = 64;  // drop-in path: modulation vectors of the last 64 distinct (t, guidance, pooled) triples stay
                             // on the device (pointwise.cuh mod_cache_*); 0 = recompute every forward
  int use_graph = 1;
  int use_pdl = 1;  // programmatic dependent launch between the kernels of a step (measured -1.5 % step time)
  int profile = 0;
  Profiler prof;
  double prof_us[KF_COUNT] = {0, 0, 0, 0, 0};
  long long prof_n[KF_COUNT] = {0, 0, 0, 0, 0};
  bool finalized = false;
  std::map<std::string, Weight> w;

  int D = 0, H = 0, dh = 0;
  long long mod_rows = 0;

  // per-request state
  int B = 0, S = 0, T = 0, N = 0;
  int pB = 0, pS = 0, pT = 0;  // shape of the last tfx_prepare: an option that drops the workspace re-prepares lazily
  std::vector<void*> allocs;
  bf16 *hidden = nullptr, *nbuf = nullptr, *cat = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *mod = nullptr;
  bf16 *x_in = nullptr, *enc_in = nullptr, *pooled_in = nullptr, *t_in = nullptr, *ids_txt = nullptr, *ids_img = nullptr;
  bf16 *tproj = nullptr, *h1 = nullptr, *temb = nullptr, *out_buf = nullptr, *lat_in = nullptr, *lat_out = nullptr;
  float* g_in = nullptr;
  float* dt_dev = nullptr;
  float2* rope = nullptr;
  // activation-side TMA descriptors, [0] text rows, [1] image rows
  enum AKind { A_NBUF = 0, A_ATTN = 1, A_MLP = 2, A_CAT = 3, A_X = 4, A_ENC = 5, A_KINDS = 6 };
  CUtensorMap mA[2][A_KINDS][2];  // [0: 128-row boxes | 1: 128/gemm_mcast-row boxes for the multicast kernels][kind][group]
  CUtensorMap mQ, mK, mV;  // [B*H, N, dh] with 128-row boxes
  // unfused LoRA (side path): a packed matrix "<m>.w" with adapters registered as "<m>.la" [64, K] (the lora_A rows of the
  // modules packed into it, zero padded) and "<m>.lb" [N, 64] ((alpha/r) * lora_B of each module in its rows x its columns)
  // runs T = bf16(x la^T) into `tbuf`, then its GEMM with one extension k-block (T, lb): x W^T + T lb^T in one accumulator
  bf16* tbuf = nullptr;      // [B*N, 64]
  CUtensorMap mT[2];         // tbuf rows of the text / image group as the A extension
  bf16* side_zero = nullptr;  // zeros: stands in for la / lb / the T bias on the stream of a launch that has no adapter
  long long side_zero_elems = 0;
  bool has_side = false;
  Attn4Workspace attn_ws;  // partial (O, m, l) of the KV shares of schedule 4
  std::map<std::string, CUtensorMap> mB;  // weight-side descriptors, keyed "<weight>#<cta_group>#<block_n>"

  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t graphs[4] = {nullptr, nullptr, nullptr, nullptr};  // [fused_euler * 2 + scheduled]
  long long graph_kernels[4] = {0, 0, 0, 0};
  // schedule-wide modulation table (tfx_set_schedule): [sched_steps * B, mod_rows]
  bf16* mod_table = nullptr;
  int sched_steps = 0;
  bf16 *sched_g = nullptr, *sched_pooled = nullptr, *sched_t = nullptr;
  std::vector<char> sched_pass_done;  // one flag per GEMV pass (kGemvMaxB rows of the table): passes run when first needed
  // modulation cache of the unscheduled (drop-in) path
  ModCacheParams mc;
  bool mc_ready = false;

  const Weight& W(const std::string& name) {
    auto it = w.find(name);
    REQUIRE(it != w.end(), TFX_ERR_MISSING, "weight '%s' was never set", name.c_str());
    return it->second;
  }
  const CUtensorMap& WB(const std::string& name, int block_n, int cg) {
    const std::string key = name + "#" + std::to_string(cg) + "#" + std::to_string(block_n);
    auto it = mB.find(key);
    if (it == mB.end()) {
      const Weight& t = W(name);
      REQUIRE(t.cols % kGemmBlockK == 0, TFX_ERR_INVALID, "weight '%s' has K=%lld, not a multiple of %d", name.c_str(), t.cols, kGemmBlockK);
      it = mB.emplace(key, make_map_2d(err_, t.ptr, t.rows, t.cols, t.cols, block_n / cg)).first;
    }
    return it->second;
  }
  // descriptor over a side-path matrix ([rows, cols] bf16, or the zero block read as that shape)
  const CUtensorMap& side_map(const std::string& name, long long rows, long long cols, int block_n, int cg) {
    const std::string key = name + "#" + std::to_string(rows) + "x" + std::to_string(cols) + "#" + std::to_string(cg) + "#" + std::to_string(block_n);
    auto it = mB.find(key);
    if (it == mB.end()) {
      const bf16* ptr = side_zero;
      if (name != "side_zero") {
        const Weight& t = W(name);
        REQUIRE(t.rows == rows && t.cols == cols, TFX_ERR_INVALID, "side matrix '%s' is [%lld,%lld], expected [%lld,%lld]", name.c_str(),
                t.rows, t.cols, rows, cols);
        ptr = t.ptr;
      } else {
        REQUIRE(rows * cols <= side_zero_elems, TFX_ERR_STATE, "zero block too small for [%lld,%lld]", rows, cols);
      }
      it = mB.emplace(key, make_map_2d(err_, ptr, rows, cols, cols, block_n / cg)).first;
    }
    return it->second;
  }
  // tile width for a two-stream (text rows + image rows) GEMM of width Nn
  int block_n_for(int Nn) const {
    const int tile_m = 128 * gemm_cta_group;
    const long long mt = ((long long)B * T + tile_m - 1) / tile_m + ((long long)B * S + tile_m - 1) / tile_m;
    const int bn = pick_block_n(mt, Nn, num_sms(device) / gemm_cta_group, gemm_narrow_tiles != 0);
    return (gemm_mcast >= 2 && bn == 192) ? 256 : bn;
  }
  // Wide-K GEMM [rows, K] x [Nn, K]^T: when A is larger than half the L2, walk it in bands of M tiles sized so that one band x all N
  // tiles is one wave -- each wave then streams band x 256 rows of A once and the whole weight, instead of the whole of A.
  // Measured in the step (profiles/r2h_band.md): cfg3 72.05 -> 71.4 ms, cfg5 142.0 -> 138.6 ms, DRAM traffic 81.6 -> 74.5 GB.
  int band_for(long long rows, long long K, int Nn, int bn) const {
    if (gemm_m_band <= -100) return gemm_m_band + 100;  // experiment: N bands of (-100 - value) tiles
    if (gemm_m_band >= 0) return gemm_m_band;
    if (rows * K * 2 <= (64LL << 20)) return 0;
    const int units = num_sms(device) / gemm_cta_group, nt = (Nn + bn - 1) / bn;
    return std::max(1, units / nt);
  }
  void drop_graphs() {
    for (int i = 0; i < 4; ++i)
      if (graphs[i]) { cudaGraphExecDestroy(graphs[i]); graphs[i] = nullptr; }
  }
  void free_workspace() {
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    drop_graphs();
    if (mod_table) { cudaFree(mod_table); mod_table = nullptr; }
    if (sched_g) { cudaFree(sched_g); sched_g = nullptr; }
    if (sched_pooled) { cudaFree(sched_pooled); sched_pooled = nullptr; }
    if (sched_t) { cudaFree(sched_t); sched_t = nullptr; }
    sched_steps = 0;
    mc_ready = false;
    B = S = T = N = 0;
  }
  template <typename Tp>
  Tp* alloc(size_t count) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, count * sizeof(Tp) + 256));
    allocs.push_back(p);
    return reinterpret_cast<Tp*>(p);
  }
  // modulation-vector offsets inside one sample's [mod_rows] vector
  long long mod_double(int i, int stream_c, int chunk) const { return ((long long)i * 12 + (stream_c ? 6 : 0) + chunk) * D; }
  long long mod_single(int j, int chunk) const { return ((long long)cfg.num_layers * 12 + (long long)j * 3 + chunk) * D; }
  long long mod_final(int chunk) const { return ((long long)cfg.num_layers * 12 + (long long)cfg.num_single_layers * 3 + chunk) * D; }

  // One GEMM of the step: A operand `kind` (text rows = group 0, image rows = group 1; `ga` picks which group's
  // descriptor feeds a single-group launch), weights w0/w1, tile width bn.
  void gemm(const LaunchCtx& c, int bn, AKind kind, const std::string& w0, const std::string& w1, const GemmParams& p, int ga = -1) {
    const int g0 = ga < 0 ? 0 : ga, g1 = ga < 0 ? 1 : ga;
    const int nt = (p.N + bn - 1) / bn;
    if (has_side && ga < 0) {
      const std::string m0 = w0.substr(0, w0.size() - 2), m1 = w1.substr(0, w1.size() - 2);  // strip ".w"
      const bool s0 = w.count(m0 + ".la") != 0, s1 = w.count(m1 + ".la") != 0;
      if (s0 || s1) {
        const int cg = gemm_cta_group;
        const std::string z = "side_zero";
        // T = bf16(x la^T): [rows, 64] per group (a group without adapters multiplies by zeros: its extension adds nothing)
        GemmParams pt;
        memset(&pt, 0, sizeof pt);
        pt.N = 64; pt.K = p.K; pt.num_groups = 2; pt.n_split = 64; pt.mode0 = pt.mode1 = EPI_STORE;
        for (int g = 0; g < 2; ++g) {
          pt.g[g].M = p.g[g].M; pt.g[g].rows_per_sample = p.g[g].rows_per_sample;
          pt.g[g].bias = side_zero; pt.g[g].out = tbuf + (g ? (long long)p.g[0].M * 64 : 0); pt.g[g].ldo = 64;
        }
        // single-CTA 128 x 64 tiles: the output is one tile wide, so the M tiles are all the parallelism there is
        launch_gemm(c, 1, 64, mA[0][kind][0], mA[0][kind][1], side_map(s0 ? m0 + ".la" : z, 64, p.K, 64, 1),
                    side_map(s1 ? m1 + ".la" : z, 64, p.K, 64, 1), pt);
        GemmParams pe = p;
        pe.k_ext = kGemmBlockK;
        GemmExt x{&mT[0], &mT[1], &side_map(s0 ? m0 + ".lb" : z, p.N, 64, bn, cg), &side_map(s1 ? m1 + ".lb" : z, p.N, 64, bn, cg)};
        launch_gemm(c, cg, bn, mA[0][kind][0], mA[0][kind][1], WB(w0, bn, cg), WB(w1, bn, cg), pe, &x);
        return;
      }
    }
    if (gemm_mcast >= 2 && nt >= gemm_mcast) {
      launch_gemm_mc(c, gemm_mcast, bn, mA[1][kind][g0], mA[1][kind][g1], WB(w0, bn, 2), WB(w1, bn, 2), p);
    } else {
      launch_gemm(c, gemm_cta_group, bn, mA[0][kind][g0], mA[0][kind][g1], WB(w0, bn, gemm_cta_group), WB(w1, bn, gemm_cta_group), p);
    }
  }
  void build_weight_maps();
  void prepare(int B_, int S_, int T_);
  void enqueue_modulation(const LaunchCtx& c, const bf16* t_rows, const float* g_rows, const bf16* pooled_rows, int rows, bf16* mod_out,
                          const int* skip = nullptr);
  void reset_modulation_caches();
  void enqueue_forward(const LaunchCtx& c, bool fused_euler, bool want_noise_pred, bool scheduled);
  void run(bool fused_euler, bool want_noise_pred, bool scheduled);
};

void tfx_model::build_weight_maps() { mB.clear(); }

void tfx_model::prepare(int B_, int S_, int T_) {
  REQUIRE(finalized, TFX_ERR_STATE, "tfx_prepare before tfx_finalize_weights");
  REQUIRE(B_ >= 1 && B_ <= kGemvMaxB && S_ >= 1 && T_ >= 1, TFX_ERR_INVALID, "unsupported problem B=%d S=%d T=%d", B_, S_, T_);
  if (B_ == B && S_ == S && T_ == T) return;
  free_workspace();
  B = B_; S = S_; T = T_; N = S + T;
  pB = B_; pS = S_; pT = T_;
  const long long R = (long long)B * N;
  hidden = alloc<bf16>(R * D);
  nbuf = alloc<bf16>(R * D);
  cat = alloc<bf16>(R * 5 * D);
  q = alloc<bf16>(R * D);
  k = alloc<bf16>(R * D);
  v = alloc<bf16>(R * D);
  mod = alloc<bf16>((long long)B * mod_rows);
  x_in = alloc<bf16>((long long)B * S * cfg.in_channels);
  enc_in = alloc<bf16>((long long)B * T * cfg.joint_attention_dim);
  pooled_in = alloc<bf16>((long long)B * cfg.pooled_projection_dim);
  t_in = alloc<bf16>(B);
  g_in = alloc<float>(B);
  dt_dev = alloc<float>(4);
  ids_txt = alloc<bf16>((long long)T * 3);
  ids_img = alloc<bf16>((long long)S * 3);
  tproj = alloc<bf16>((long long)kGemvMaxB * 256);
  h1 = alloc<bf16>((long long)kGemvMaxB * D);
  temb = alloc<bf16>((long long)kGemvMaxB * D);
  out_buf = alloc<bf16>((long long)B * S * cfg.out_channels);
  lat_in = alloc<bf16>((long long)B * S * cfg.out_channels);
  lat_out = alloc<bf16>((long long)B * S * cfg.out_channels);
  rope = alloc<float2>((long long)N * (dh / 2));
  tbuf = alloc<bf16>(R * 64);
  CUDA_TRY(cudaMemset(cat, 0, R * 5 * D * sizeof(bf16)));
  mc_ready = false;
  if (mod_cache_slots > 0) {
    const int P = cfg.pooled_projection_dim;
    memset(&mc, 0, sizeof mc);
    mc.t = t_in; mc.g = cfg.guidance_embeds ? g_in : nullptr; mc.pooled = pooled_in;
    mc.B = B; mc.P = P; mc.slots = mod_cache_slots;
    mc.key_t = alloc<uint16_t>((size_t)mod_cache_slots * B);
    mc.key_g = alloc<uint32_t>((size_t)mod_cache_slots * B);
    mc.key_p = alloc<uint16_t>((size_t)mod_cache_slots * B * P);
    mc.state = alloc<int>(8);
    mc.mod = mod; mc.mod_elems = (long long)B * mod_rows;
    mc.table = alloc<bf16>((size_t)mod_cache_slots * B * mod_rows);
    CUDA_TRY(cudaMemset(mc.state, 0, 8 * sizeof(int)));
    mc_ready = true;
  }

  const long long rt = (long long)B * T, ri = (long long)B * S;
  const long long row0[2] = {0, rt}, rows[2] = {rt, ri};
  for (int vi = 0; vi < 2; ++vi) {
    const int box = (vi == 0 || gemm_mcast < 2) ? 128 : 128 / gemm_mcast;
    for (int g = 0; g < 2; ++g) {
      mA[vi][A_NBUF][g] = make_map_2d(err_, nbuf + row0[g] * D, rows[g], D, D, box);
      mA[vi][A_ATTN][g] = make_map_2d(err_, cat + row0[g] * 5 * D, rows[g], D, 5LL * D, box);
      mA[vi][A_MLP][g] = make_map_2d(err_, cat + row0[g] * 5 * D + D, rows[g], 4LL * D, 5LL * D, box);
      mA[vi][A_CAT][g] = make_map_2d(err_, cat + row0[g] * 5 * D, rows[g], 5LL * D, 5LL * D, box);
      mA[vi][A_X][g] = make_map_2d(err_, x_in, ri, cfg.in_channels, cfg.in_channels, box);
      mA[vi][A_ENC][g] = make_map_2d(err_, enc_in, rt, cfg.joint_attention_dim, cfg.joint_attention_dim, box);
    }
  }
  mT[0] = make_map_2d(err_, tbuf, rt, 64, 64, 128);
  mT[1] = make_map_2d(err_, tbuf + rt * 64, ri, 64, 64, 128);
  mQ = make_map_3d(err_, q, (long long)B * H, N, dh);
  mK = make_map_3d(err_, k, (long long)B * H, N, dh);
  mV = make_map_3d(err_, v, (long long)B * H, N, dh);
  attn_ws.ws_o = reinterpret_cast<float*>(alloc<uint8_t>(Attn4Workspace::o_bytes(dh)));
  attn_ws.ws_ml = reinterpret_cast<float*>(alloc<uint8_t>(Attn4Workspace::ml_bytes()));
  attn_ws.counters = reinterpret_cast<int*>(alloc<uint8_t>(Attn4Workspace::counter_bytes()));
  CUDA_TRY(cudaMemset(attn_ws.counters, 0, Attn4Workspace::counter_bytes()));
}

// The launch sequence of one FluxTransformer2DModel.forward (+ optional fused Euler update).
// temb = time_text_embed(timestep, guidance, pooled) (transformer_flux.py:1088-1098, embeddings.py:1327-1339) for `rows`
// (<= 8) independent rows, then every adaLN `linear(silu(temb))` in one pass over the [mod_rows, D] matrix.
void tfx_model::enqueue_modulation(const LaunchCtx& c, const bf16* t_rows, const float* g_rows, const bf16* pooled_rows, int rows,
                                   bf16* mod_out, const int* skip) {
  timestep_embed_kernel<<<rows, 128, 0, c.stream>>>(t_rows, 0, rows, tproj, skip);
  ++*c.counter;
  launch_gemv(c, tproj, rows, 256, W("t_embed.l1.w").ptr, W("t_embed.l1.b").ptr, D, h1, GEMV_POST_SILU, skip);
  launch_gemv(c, h1, rows, D, W("t_embed.l2.w").ptr, W("t_embed.l2.b").ptr, D, temb, 0, skip);
  if (cfg.guidance_embeds) {
    timestep_embed_kernel<<<rows, 128, 0, c.stream>>>(g_rows, 1, rows, tproj, skip);
    ++*c.counter;
    launch_gemv(c, tproj, rows, 256, W("g_embed.l1.w").ptr, W("g_embed.l1.b").ptr, D, h1, GEMV_POST_SILU, skip);
    launch_gemv(c, h1, rows, D, W("g_embed.l2.w").ptr, W("g_embed.l2.b").ptr, D, temb, GEMV_ADD_TO_OUT, skip);
  }
  launch_gemv(c, pooled_rows, rows, cfg.pooled_projection_dim, W("p_embed.l1.w").ptr, W("p_embed.l1.b").ptr, D, h1, GEMV_POST_SILU, skip);
  launch_gemv(c, h1, rows, D, W("p_embed.l2.w").ptr, W("p_embed.l2.b").ptr, D, temb, GEMV_ADD_TO_OUT, skip);
  CUDA_TRY(cudaGetLastError());
  launch_gemv(c, temb, rows, D, W("mod.w").ptr, W("mod.b").ptr, mod_rows, mod_out, GEMV_PRE_SILU, skip);
}

// Weights changed in place (LoRA hot-swap): nothing cached from the old weights may be served again.
void tfx_model::reset_modulation_caches() {
  if (mc_ready) CUDA_TRY(cudaMemsetAsync(mc.state, 0, 8 * sizeof(int), stream));
  std::fill(sched_pass_done.begin(), sched_pass_done.end(), 0);
}

void tfx_model::enqueue_forward(const LaunchCtx& c, bool fused_euler, bool want_noise_pred, bool scheduled) {
  const long long rt = (long long)B * T, ri = (long long)B * S;
  const int L = cfg.num_layers, Ls = cfg.num_single_layers;
  char nm[96];
  auto name = [&](const char* fmt, int i, const char* suffix) {
    snprintf(nm, sizeof nm, fmt, i);
    return std::string(nm) + suffix;
  };

  // --- positional table (pos_embed(cat(txt_ids, img_ids)), transformer_flux.py:1114-1115)
  NvtxRange nvtx_fwd("enqueue_forward");
  {
    RopeParams rp;
    rp.txt_ids = ids_txt; rp.img_ids = ids_img; rp.T = T; rp.S = S;
    rp.axes[0] = cfg.axes_dims_rope[0]; rp.axes[1] = cfg.axes_dims_rope[1]; rp.axes[2] = cfg.axes_dims_rope[2];
    rp.half_dim = dh / 2; rp.out = rope;
    const int total = N * (dh / 2);
    CUDA_TRY(launch_ex(rope_table_kernel, dim3((total + 255) / 256), dim3(256), 0, c, 1, rp));
    ++*c.counter;
  }
  // --- temb + every adaLN `linear(silu(temb))` of the step; with a schedule set they were computed for all steps at
  //     once (tfx_set_schedule) and this step's rows were copied into `mod` before the launch
  if (!scheduled) {
    // drop-in path: the pipeline passes (t, guidance, pooled) per call; a triple seen before (the same schedule step of an
    // earlier image) is served from the device-side cache, a new one is computed and stored -- one static graph either way
    const bool cached = mc_ready && mod_cache_slots > 0;
    if (cached) {
      ProfScope ps(c, KF_MISC);
      mod_cache_lookup_kernel<<<1, 256, 0, c.stream>>>(mc);
      ++*c.counter;
    }
    enqueue_modulation(c, t_in, g_in, pooled_in, B, mod, cached ? mc.state : nullptr);
    if (cached) {
      ProfScope ps(c, KF_MISC);
      mod_cache_commit_kernel<<<num_sms(device) * 2, 256, 0, c.stream>>>(mc);
      CUDA_TRY(cudaGetLastError());
      ++*c.counter;
    }
  }

  auto base_params = [&](int Nn, int Kk) {
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = Nn; p.K = Kk; p.num_groups = 2; p.n_split = Nn; p.mode0 = EPI_STORE; p.mode1 = EPI_STORE;
    p.D = D; p.head_dim = dh; p.num_heads = H; p.n_joint = N; p.q = q; p.k = k; p.v = v; p.rope = rope;
    p.rms_eps = 1e-6f; p.dt_ptr = dt_dev;
    p.debug_flags = gemm_l2_hints;
    p.g[0].M = (int)rt; p.g[0].rows_per_sample = T; p.g[0].pos_offset = 0;
    p.g[1].M = (int)ri; p.g[1].rows_per_sample = S; p.g[1].pos_offset = T;
    return p;
  };
  // --- embedders (transformer_flux.py:1086,1099): text rows then image rows of `hidden`
  {
    GemmParams p = base_params(D, cfg.joint_attention_dim);
    p.num_groups = 1;
    p.g[0].bias = W("context_embedder.b").ptr; p.g[0].out = hidden; p.g[0].ldo = D;
    gemm(c, 256, A_ENC, "context_embedder.w", "context_embedder.w", p, 0);
    GemmParams px = base_params(D, cfg.in_channels);
    px.num_groups = 1;
    px.g[0] = px.g[1];
    px.g[0].bias = W("x_embedder.b").ptr; px.g[0].out = hidden + rt * D; px.g[0].ldo = D;
    gemm(c, 256, A_X, "x_embedder.w", "x_embedder.w", px, 0);
  }
  AttnParams ap;
  ap.B = B; ap.H = H; ap.N = N; ap.T = T; ap.S = S;
  ap.scale_log2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  ap.out = cat; ap.ld_out = 5LL * D;

  const int av = attn_variant ? attn_variant : pick_attention_variant(ap, num_sms(device));

  LnModParams lp;
  lp.x = hidden; lp.y = nbuf; lp.rows = (int)(rt + ri); lp.D = D; lp.row_begin = 0; lp.rows0 = (int)rt;
  lp.rows_per0 = T; lp.rows_per1 = S; lp.mod = mod; lp.mod_stride = mod_rows; lp.eps = 1e-6f;

  const int bn_d = block_n_for(D), bn_4d = block_n_for(4 * D);  // tile widths chosen against wave quantisation
  bf16* hid_g[2] = {hidden, hidden + rt * D};
  bf16* cat_g[2] = {cat, cat + rt * 5 * D};

  // --- 19 x FluxTransformerBlock (transformer_flux.py:794-841); chunk order shift,scale,gate (msa) shift,scale,gate (mlp)
  for (int i = 0; i < L; ++i) {
    NvtxRange nvtx_blk(name("double block %d", i, "").c_str());
    const char* sfx[2] = {"_c", "_x"};
    lp.shift0 = mod_double(i, 1, 0); lp.scale0 = mod_double(i, 1, 1);
    lp.shift1 = mod_double(i, 0, 0); lp.scale1 = mod_double(i, 0, 1);
    launch_ln_modulate(c, lp);
    {
      GemmParams p = base_params(3 * D, D);
      p.mode0 = EPI_QKV;
      for (int g = 0; g < 2; ++g) {
        p.g[g].bias = W(name("d%d.qkv", i, sfx[g]) + ".b").ptr;
        p.g[g].rms_q = W(name("d%d.rms_q", i, sfx[g])).ptr;
        p.g[g].rms_k = W(name("d%d.rms_k", i, sfx[g])).ptr;
      }
      gemm(c, 256, A_NBUF, name("d%d.qkv_c", i, ".w"), name("d%d.qkv_x", i, ".w"), p);
    }
    if (av == 9) launch_attention5(c, dh, attn_emu, mQ, mK, mV, ap, attn_ws);
    else if (av == 8) launch_attention4(c, dh, attn_emu, mQ, mK, mV, ap, attn_ws);
    else launch_attention3(c, dh, attn_emu, av != 4, false, mQ, mK, mV, ap);
    {
      GemmParams p = base_params(D, D);
      p.mode0 = EPI_GATE_RES;
      for (int g = 0; g < 2; ++g) {
        p.g[g].bias = W(name("d%d.out", i, sfx[g]) + ".b").ptr;
        p.g[g].out = hid_g[g]; p.g[g].ldo = D; p.g[g].res = hid_g[g]; p.g[g].ldr = D;
        p.g[g].gate = mod + mod_double(i, g == 0, 2); p.g[g].gate_stride = mod_rows;
      }
      gemm(c, bn_d, A_ATTN, name("d%d.out_c", i, ".w"), name("d%d.out_x", i, ".w"), p);
    }
    lp.shift0 = mod_double(i, 1, 3); lp.scale0 = mod_double(i, 1, 4);
    lp.shift1 = mod_double(i, 0, 3); lp.scale1 = mod_double(i, 0, 4);
    launch_ln_modulate(c, lp);
    {
      GemmParams p = base_params(4 * D, D);
      p.mode0 = EPI_GELU;
      for (int g = 0; g < 2; ++g) {
        p.g[g].bias = W(name("d%d.ff1", i, sfx[g]) + ".b").ptr;
        p.g[g].out = cat_g[g] + D; p.g[g].ldo = 5LL * D;
      }
      gemm(c, bn_4d, A_NBUF, name("d%d.ff1_c", i, ".w"), name("d%d.ff1_x", i, ".w"), p);
    }
    {
      GemmParams p = base_params(D, 4 * D);
      p.mode0 = EPI_GATE_RES;
      for (int g = 0; g < 2; ++g) {
        p.g[g].bias = W(name("d%d.ff2", i, sfx[g]) + ".b").ptr;
        p.g[g].out = hid_g[g]; p.g[g].ldo = D; p.g[g].res = hid_g[g]; p.g[g].ldr = D;
        p.g[g].gate = mod + mod_double(i, g == 0, 5); p.g[g].gate_stride = mod_rows;
      }
      p.m_band = band_for(rt + ri, 4LL * D, D, bn_d);
      p.k_snake = (p.m_band != 0) ? gemm_k_snake : 0;
      gemm(c, bn_d, A_MLP, name("d%d.ff2_c", i, ".w"), name("d%d.ff2_x", i, ".w"), p);
    }
  }
  // --- 38 x FluxSingleTransformerBlock (transformer_flux.py:715-739) on the joint [text;image] rows (the cat of
  //     :1160 is the row layout of `hidden` itself); chunk order shift, scale, gate
  for (int j = 0; j < Ls; ++j) {
    NvtxRange nvtx_blk(name("single block %d", j, "").c_str());
    lp.shift0 = lp.shift1 = mod_single(j, 0);
    lp.scale0 = lp.scale1 = mod_single(j, 1);
    launch_ln_modulate(c, lp);
    {
      GemmParams p = base_params(7 * D, D);
      p.n_split = 3 * D; p.mode0 = EPI_QKV; p.mode1 = EPI_GELU; p.col_offset1 = D;
      for (int g = 0; g < 2; ++g) {
        p.g[g].bias = W(name("s%d.qkvmlp", j, ".b")).ptr;
        p.g[g].rms_q = W(name("s%d.rms_q", j, "")).ptr;
        p.g[g].rms_k = W(name("s%d.rms_k", j, "")).ptr;
        p.g[g].out = cat_g[g]; p.g[g].ldo = 5LL * D;
      }
      const std::string wn = name("s%d.qkvmlp", j, ".w");
      gemm(c, 256, A_NBUF, wn, wn, p);
    }
    if (av == 9) launch_attention5(c, dh, attn_emu, mQ, mK, mV, ap, attn_ws);
    else if (av == 8) launch_attention4(c, dh, attn_emu, mQ, mK, mV, ap, attn_ws);
    else launch_attention3(c, dh, attn_emu, av != 4, false, mQ, mK, mV, ap);
    {
      GemmParams p = base_params(D, 5 * D);
      p.mode0 = EPI_GATE_RES;
      for (int g = 0; g < 2; ++g) {
        p.g[g].bias = W(name("s%d.out", j, ".b")).ptr;
        p.g[g].out = hid_g[g]; p.g[g].ldo = D; p.g[g].res = hid_g[g]; p.g[g].ldr = D;
        p.g[g].gate = mod + mod_single(j, 2); p.g[g].gate_stride = mod_rows;
      }
      const std::string wn = name("s%d.out", j, ".w");
      p.m_band = band_for(rt + ri, 5LL * D, D, bn_d);
      p.k_snake = (p.m_band != 0) ? gemm_k_snake : 0;
      gemm(c, bn_d, A_CAT, wn, wn, p);
    }
  }
  // --- norm_out (AdaLayerNormContinuous: chunk order scale, shift) + proj_out on the image rows (:1200-1203)
  lp.row_begin = (int)rt;
  lp.shift1 = mod_final(1); lp.scale1 = mod_final(0);
  launch_ln_modulate(c, lp);
  {
    const int C = cfg.out_channels;
    GemmParams p = base_params(C, D);
    p.num_groups = 1;
    p.g[0] = p.g[1];
    p.g[0].bias = W("proj_out.b").ptr;
    if (fused_euler) {
      p.mode0 = EPI_EULER;
      p.g[0].out = want_noise_pred ? out_buf : nullptr; p.g[0].ldo = C;
      p.g[0].res = lat_in; p.g[0].ldr = C; p.g[0].out2 = lat_out;
    } else {
      p.g[0].out = out_buf; p.g[0].ldo = C;
    }
    gemm(c, 256, A_NBUF, "proj_out.w", "proj_out.w", p, 1);
  }
}

void tfx_model::run(bool fused_euler, bool want_noise_pred, bool scheduled) {
  LaunchCtx c{stream, device, &launches, err_};
  c.pdl = use_pdl != 0;
  const int gi = (fused_euler ? 2 : 0) + (scheduled ? 1 : 0);
  if (profile) {
    c.prof = &prof;
    enqueue_forward(c, fused_euler, want_noise_pred, scheduled);
    double us[KF_COUNT];
    long long n[KF_COUNT];
    prof.collect(us, n);
    for (int i = 0; i < KF_COUNT; ++i) { prof_us[i] += us[i]; prof_n[i] += n[i]; }
    return;
  }
  if (!use_graph) {
    enqueue_forward(c, fused_euler, want_noise_pred, scheduled);
    return;
  }
  if (!graphs[gi]) {
    long long scratch = 0;
    LaunchCtx cc{stream, device, &scratch, err_};
    cc.pdl = use_pdl != 0;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_forward(cc, fused_euler, true, scheduled);
    } catch (...) {
      cudaStreamEndCapture(stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_TRY(cudaStreamEndCapture(stream, &graph));
    CUDA_TRY(cudaGraphInstantiate(&graphs[gi], graph, 0));
    cudaGraphDestroy(graph);
    graph_kernels[gi] = scratch;
    graph_nodes = scratch;
  }
  CUDA_TRY(cudaGraphLaunch(graphs[gi], stream));
  launches += graph_kernels[gi];
}

// ==================================================================================================== C ABI
#define API_BEGIN(h)                       \
  std::string* err_ = (h) ? &(h)->err : nullptr; \
  (void)err_;                              \
  try {
#define API_END                                                    \
  }                                                                \
  catch (const Fail& f) { return f.code; }                         \
  catch (const std::exception& e) { return set_error(err_, TFX_ERR_INVALID, "%s", e.what()); } \
  return TFX_OK;

extern "C" {

const char* tfx_last_error(tfx_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int tfx_create(const tfx_config* cfg, int32_t device, tfx_handle* out) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(cfg && out, TFX_ERR_INVALID, "null argument");
    REQUIRE(cfg->attention_head_dim == 64 || cfg->attention_head_dim == 128, TFX_ERR_INVALID,
            "attention_head_dim %d unsupported (64 or 128)", cfg->attention_head_dim);
    const int D = cfg->attention_head_dim * cfg->num_attention_heads;
    REQUIRE(D % 256 == 0, TFX_ERR_INVALID, "inner dim %d must be a multiple of 256", D);
    REQUIRE(cfg->axes_dims_rope[0] + cfg->axes_dims_rope[1] + cfg->axes_dims_rope[2] == cfg->attention_head_dim,
            TFX_ERR_INVALID, "axes_dims_rope must sum to attention_head_dim");
    REQUIRE(cfg->in_channels % 64 == 0 && cfg->joint_attention_dim % 64 == 0 && cfg->pooled_projection_dim % 8 == 0,
            TFX_ERR_INVALID, "in_channels / joint_attention_dim must be multiples of 64, pooled_projection_dim of 8");
    REQUIRE(cfg->out_channels % 8 == 0 && cfg->out_channels <= 256, TFX_ERR_INVALID, "out_channels %d unsupported", cfg->out_channels);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, TFX_ERR_INVALID, "device %d not present (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, TFX_ERR_INVALID, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    configure_kernels(err_);
    tfx_model* m = new tfx_model();
    m->cfg = *cfg;
    m->device = device;
    m->D = D;
    m->H = cfg->num_attention_heads;
    m->dh = cfg->attention_head_dim;
    m->mod_rows = ((long long)cfg->num_layers * 12 + (long long)cfg->num_single_layers * 3 + 2) * D;
    cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&m->ev_out, cudaEventDisableTiming);
    *out = m;
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

void tfx_destroy(tfx_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->free_workspace();
  if (h->side_zero) cudaFree(h->side_zero);
  cudaEventDestroy(h->ev_in);
  cudaEventDestroy(h->ev_out);
  cudaStreamDestroy(h->stream);
  delete h;
}

int tfx_set_option(tfx_handle h, const char* key, int64_t value) {
  API_BEGIN(h)
  REQUIRE(h && key, TFX_ERR_INVALID, "null argument");
  std::string k(key);
  if (k == "gemm_cta_group") {
    REQUIRE(value == 1 || value == 2, TFX_ERR_INVALID, "gemm_cta_group must be 1 or 2");
    h->gemm_cta_group = (int)value;
  } else if (k == "gemm_mcast") {
    REQUIRE(value == 0 || value == 2 || value == 4, TFX_ERR_INVALID, "gemm_mcast must be 0, 2 or 4");
    h->gemm_mcast = (int)value;
    h->free_workspace();  // A-side descriptors depend on it: the next tfx_prepare rebuilds them
  } else if (k == "attn_variant") {
    REQUIRE(value == 0 || value == 4 || value == 5 || value == 8 || value == 9, TFX_ERR_INVALID, "attn_variant must be 0, 4, 5, 8 or 9");
    h->attn_variant = (int)value;
  } else if (k == "attn_emu") {
    REQUIRE(value == 0 || (value >= 2 && value <= 4), TFX_ERR_INVALID, "attn_emu must be 0, 2, 3 or 4");
    h->attn_emu = (int)value;
  } else if (k == "gemm_narrow_tiles") {
    h->gemm_narrow_tiles = value != 0;  // weight-side descriptors are keyed by tile width: nothing to rebuild
  } else if (k == "gemm_k_snake") {
    REQUIRE(value == 0 || value == 1, TFX_ERR_INVALID, "gemm_k_snake must be 0 or 1");
    h->gemm_k_snake = (int)value;
    h->drop_graphs();
  } else if (k == "gemm_m_band") {
    REQUIRE(value >= -164 && value <= 64, TFX_ERR_INVALID, "gemm_m_band must be -1 (per shape), 0..64 (M bands) or -100 - n (N bands of n tiles)");
    h->gemm_m_band = (int)value;
    h->drop_graphs();
  } else if (k == "gemm_l2_hints") {
    REQUIRE(value >= 0 && value <= 3, TFX_ERR_INVALID, "gemm_l2_hints must be 0..3");
    h->gemm_l2_hints = (int)value;
  } else if (k == "mod_cache_slots") {
    REQUIRE(value >= 0 && value <= 4096, TFX_ERR_INVALID, "mod_cache_slots must be 0..4096");
    h->mod_cache_slots = (int)value;
    h->free_workspace();  // the cache is part of the workspace: the next tfx_prepare sizes it
  } else if (k == "mod_cache_reset") {
    h->reset_modulation_caches();
  } else if (k == "use_graph") {
    h->use_graph = value != 0;
  } else if (k == "use_pdl") {
    h->use_pdl = value != 0;
  } else if (k == "profile") {
    h->profile = value != 0;
    for (int i = 0; i < KF_COUNT; ++i) { h->prof_us[i] = 0; h->prof_n[i] = 0; }
  } else {
    REQUIRE(false, TFX_ERR_INVALID, "unknown option '%s'", key);
  }
  h->drop_graphs();
  API_END
}

int tfx_get_counter(tfx_handle h, const char* key, int64_t* value) {
  API_BEGIN(h)
  REQUIRE(h && key && value, TFX_ERR_INVALID, "null argument");
  std::string k(key);
  if (k == "launches") *value = h->launches;
  else if (k == "graph_nodes") *value = h->graph_nodes;
  else if (k == "mod_cache_hits" || k == "mod_cache_valid") {
    int st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h->mc_ready) {
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      CUDA_TRY(cudaMemcpy(st, h->mc.state, sizeof st, cudaMemcpyDeviceToHost));
    }
    *value = (k == "mod_cache_hits") ? st[3] : st[2];
  }
  else if (k.compare(0, 8, "prof_us_") == 0 || k.compare(0, 7, "prof_n_") == 0) {
    const bool is_us = k.compare(0, 8, "prof_us_") == 0;
    const std::string fam = k.substr(is_us ? 8 : 7);
    static const char* names[KF_COUNT] = {"gemm", "attn", "ln", "gemv", "misc"};
    int idx = -1;
    for (int i = 0; i < KF_COUNT; ++i) if (fam == names[i]) idx = i;
    REQUIRE(idx >= 0, TFX_ERR_INVALID, "unknown kernel family '%s'", fam.c_str());
    *value = is_us ? (long long)(h->prof_us[idx] + 0.5) : h->prof_n[idx];
  }
  else REQUIRE(false, TFX_ERR_INVALID, "unknown counter '%s'", key);
  API_END
}

int tfx_set_weight(tfx_handle h, const char* name, const void* dev_ptr, int64_t rows, int64_t cols) {
  API_BEGIN(h)
  REQUIRE(h && name && dev_ptr, TFX_ERR_INVALID, "null argument");
  REQUIRE(rows > 0 && cols > 0, TFX_ERR_INVALID, "weight '%s' has empty shape", name);
  REQUIRE((reinterpret_cast<uintptr_t>(dev_ptr) & 15) == 0, TFX_ERR_INVALID, "weight '%s' is not 16-byte aligned", name);
  Weight t;
  t.ptr = reinterpret_cast<const bf16*>(dev_ptr);
  t.rows = rows;
  t.cols = cols;
  h->w[name] = t;
  h->finalized = false;
  API_END
}

int tfx_finalize_weights(tfx_handle h) {
  API_BEGIN(h)
  REQUIRE(h, TFX_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  const tfx_config& c = h->cfg;
  const long long D = h->D;
  auto need = [&](const std::string& n, long long rows, long long cols) {
    const Weight& t = h->W(n);
    REQUIRE(t.rows == rows && t.cols == cols, TFX_ERR_INVALID, "weight '%s' is [%lld,%lld], expected [%lld,%lld]", n.c_str(),
            t.rows, t.cols, rows, cols);
  };
  auto lin = [&](const std::string& n, long long o, long long i) { need(n + ".w", o, i); need(n + ".b", 1, o); };
  lin("x_embedder", D, c.in_channels);
  lin("context_embedder", D, c.joint_attention_dim);
  lin("t_embed.l1", D, 256); lin("t_embed.l2", D, D);
  if (c.guidance_embeds) { lin("g_embed.l1", D, 256); lin("g_embed.l2", D, D); }
  lin("p_embed.l1", D, c.pooled_projection_dim); lin("p_embed.l2", D, D);
  lin("mod", h->mod_rows, D);
  lin("proj_out", c.out_channels, D);
  char b[64];
  for (int i = 0; i < c.num_layers; ++i) {
    for (const char* s : {"_x", "_c"}) {
      snprintf(b, sizeof b, "d%d.", i);
      std::string p(b);
      lin(p + "qkv" + s, 3 * D, D); lin(p + "out" + s, D, D); lin(p + "ff1" + s, 4 * D, D); lin(p + "ff2" + s, D, 4 * D);
      need(p + "rms_q" + s, 1, h->dh); need(p + "rms_k" + s, 1, h->dh);
    }
  }
  for (int j = 0; j < c.num_single_layers; ++j) {
    snprintf(b, sizeof b, "s%d.", j);
    std::string p(b);
    lin(p + "qkvmlp", 7 * D, D); lin(p + "out", D, 5 * D);
    need(p + "rms_q", 1, h->dh); need(p + "rms_k", 1, h->dh);
  }
  // unfused-LoRA side matrices: "<m>.la" [64, K] and "<m>.lb" [N, 64] next to a packed "<m>.w" [N, K], always as a pair
  h->has_side = false;
  for (const auto& kv : h->w) {
    const std::string& n = kv.first;
    if (n.size() < 3) continue;
    const std::string sfx = n.substr(n.size() - 3), m = n.substr(0, n.size() - 3);
    if (sfx != ".la" && sfx != ".lb") continue;
    REQUIRE(h->w.count(m + ".w") && h->w.count(m + ".la") && h->w.count(m + ".lb"), TFX_ERR_MISSING,
            "side adapter '%s' needs '%s.w', '%s.la' and '%s.lb'", n.c_str(), m.c_str(), m.c_str(), m.c_str());
    const Weight& base = h->W(m + ".w");
    if (sfx == ".la") need(n, 64, base.cols); else need(n, base.rows, 64);
    h->has_side = true;
  }
  if (h->has_side && !h->side_zero) {
    h->side_zero_elems = 64LL * 7 * D;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->side_zero), (size_t)h->side_zero_elems * 2));
    CUDA_TRY(cudaMemset(h->side_zero, 0, (size_t)h->side_zero_elems * 2));
  }
  h->build_weight_maps();
  h->drop_graphs();  // the launch sequence depends on which matrices carry side adapters
  h->finalized = true;
  API_END
}

int tfx_unset_weight(tfx_handle h, const char* name) {
  API_BEGIN(h)
  REQUIRE(h && name, TFX_ERR_INVALID, "null argument");
  h->w.erase(name);
  h->finalized = false;
  API_END
}

int tfx_prepare(tfx_handle h, int32_t B, int32_t S, int32_t T) {
  API_BEGIN(h)
  REQUIRE(h, TFX_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  h->prepare(B, S, T);
  API_END
}

static void stage_common(tfx_model* h, std::string* err_, const void* enc, const void* pooled, const void* t, const void* g,
                         const void* img_ids, const void* txt_ids, cudaStream_t user, bool scheduled = false) {
  if (h->B == 0 && h->pB > 0) h->prepare(h->pB, h->pS, h->pT);  // an option dropped the workspace since
  REQUIRE(h->B > 0, TFX_ERR_STATE, "tfx_prepare has not been called");
  REQUIRE(enc && img_ids && txt_ids && (scheduled || (pooled && t)), TFX_ERR_INVALID, "null input pointer");
  REQUIRE(scheduled || !h->cfg.guidance_embeds || g, TFX_ERR_INVALID, "guidance is required when guidance_embeds is set");
  CUDA_TRY(cudaEventRecord(h->ev_in, user));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_in, 0));
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaMemcpyAsync(h->enc_in, enc, (size_t)h->B * h->T * h->cfg.joint_attention_dim * 2, cudaMemcpyDeviceToDevice, s));
  if (!scheduled) {
    CUDA_TRY(cudaMemcpyAsync(h->pooled_in, pooled, (size_t)h->B * h->cfg.pooled_projection_dim * 2, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->t_in, t, (size_t)h->B * 2, cudaMemcpyDeviceToDevice, s));
    if (h->cfg.guidance_embeds) CUDA_TRY(cudaMemcpyAsync(h->g_in, g, (size_t)h->B * 4, cudaMemcpyDeviceToDevice, s));
  }
  CUDA_TRY(cudaMemcpyAsync(h->ids_img, img_ids, (size_t)h->S * 3 * 2, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->ids_txt, txt_ids, (size_t)h->T * 3 * 2, cudaMemcpyDeviceToDevice, s));
}

static void finish(tfx_model* h, std::string* err_, cudaStream_t user) {
  CUDA_TRY(cudaEventRecord(h->ev_out, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(user, h->ev_out, 0));
}

int tfx_forward(tfx_handle h, const void* hidden_states, const void* encoder_hidden_states, const void* pooled,
                const void* timestep_bf16, const void* guidance_f32, const void* img_ids, const void* txt_ids,
                void* out_sample, void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_forward");
  REQUIRE(h && hidden_states && out_sample, TFX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  stage_common(h, err_, encoder_hidden_states, pooled, timestep_bf16, guidance_f32, img_ids, txt_ids, user);
  CUDA_TRY(cudaMemcpyAsync(h->x_in, hidden_states, (size_t)h->B * h->S * h->cfg.in_channels * 2, cudaMemcpyDeviceToDevice, h->stream));
  h->run(false, true, false);
  CUDA_TRY(cudaMemcpyAsync(out_sample, h->out_buf, (size_t)h->B * h->S * h->cfg.out_channels * 2, cudaMemcpyDeviceToDevice, h->stream));
  finish(h, err_, user);
  API_END
}

int tfx_step(tfx_handle h, const void* latents_in, const void* cond, const void* encoder_hidden_states, const void* pooled,
             const void* timestep_bf16, const void* guidance_f32, const void* img_ids, const void* txt_ids, float sigma,
             float sigma_next, void* latents_out, void* noise_pred_out, void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_step");
  REQUIRE(h && latents_in && cond && latents_out, TFX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  stage_common(h, err_, encoder_hidden_states, pooled, timestep_bf16, guidance_f32, img_ids, txt_ids, user);
  const int Cl = h->cfg.out_channels, Cc = h->cfg.in_channels - h->cfg.out_channels, Ci = h->cfg.in_channels;
  const size_t rows = (size_t)h->B * h->S;
  cudaStream_t s = h->stream;
  // hidden_states = cat(latents, cond, dim=2) (pipeline_flux_fill.py:2085) as two strided copies into the staging tile
  CUDA_TRY(cudaMemcpy2DAsync(h->x_in, (size_t)Ci * 2, latents_in, (size_t)Cl * 2, (size_t)Cl * 2, rows, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpy2DAsync(h->x_in + Cl, (size_t)Ci * 2, cond, (size_t)Cc * 2, (size_t)Cc * 2, rows, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->lat_in, latents_in, rows * Cl * 2, cudaMemcpyDeviceToDevice, s));
  const float dt = __bfloat162float(__float2bfloat16_rn(sigma_next - sigma));
  set_float_kernel<<<1, 1, 0, s>>>(h->dt_dev, dt);
  ++h->launches;
  h->run(true, noise_pred_out != nullptr, false);
  CUDA_TRY(cudaMemcpyAsync(latents_out, h->lat_out, rows * Cl * 2, cudaMemcpyDeviceToDevice, s));
  if (noise_pred_out) CUDA_TRY(cudaMemcpyAsync(noise_pred_out, h->out_buf, rows * Cl * 2, cudaMemcpyDeviceToDevice, s));
  finish(h, err_, user);
  API_END
}

// Step-invariant work hoisted out of the loop: temb depends only on (t_i, guidance, pooled),
// all known once set_timesteps has run, so the modulation vectors of every step are produced here with ceil(n*B/8)
// passes over the 6.5 GB adaLN matrix instead of one pass per step.
int tfx_set_schedule(tfx_handle h, const void* timesteps_bf16, int32_t n_steps, const void* guidance_f32, const void* pooled,
                     void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_set_schedule");
  REQUIRE(h && timesteps_bf16 && pooled && n_steps > 0, TFX_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->B == 0 && h->pB > 0) h->prepare(h->pB, h->pS, h->pT);
  REQUIRE(h->B > 0, TFX_ERR_STATE, "tfx_prepare has not been called");
  REQUIRE(!h->cfg.guidance_embeds || guidance_f32, TFX_ERR_INVALID, "guidance is required when guidance_embeds is set");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  const int B = h->B, P = h->cfg.pooled_projection_dim;
  const long long rows = (long long)n_steps * B;
  if (h->sched_steps != n_steps) {
    if (h->mod_table) cudaFree(h->mod_table);
    if (h->sched_g) cudaFree(h->sched_g);
    if (h->sched_pooled) cudaFree(h->sched_pooled);
    if (h->sched_t) cudaFree(h->sched_t);
    h->mod_table = nullptr; h->sched_g = nullptr; h->sched_pooled = nullptr; h->sched_t = nullptr; h->sched_steps = 0;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->mod_table), (size_t)rows * h->mod_rows * 2));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->sched_g), (size_t)rows * 4 + 256));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->sched_pooled), (size_t)rows * P * 2 + 256));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->sched_t), (size_t)rows * 2 + 256));
    h->sched_steps = n_steps;
  }
  CUDA_TRY(cudaEventRecord(h->ev_in, user));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_in, 0));
  cudaStream_t s = h->stream;
  // row r = step * B + b uses timestep[step][b], guidance[b], pooled[b]
  for (int st = 0; st < n_steps; ++st) {
    if (h->cfg.guidance_embeds)
      CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<float*>(h->sched_g) + (size_t)st * B, guidance_f32, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->sched_pooled + (size_t)st * B * P, pooled, (size_t)B * P * 2, cudaMemcpyDeviceToDevice, s));
  }
  // The table is filled lazily, one GEMV pass (kGemvMaxB rows = steps x samples) at a time, when tfx_step_scheduled first
  // needs a row of it: a run that stops after k steps pays for ceil(k*B / 8) passes over the 6.5 GB modulation matrix,
  // not for the whole schedule.  Same kernels on the same rows, so the values are identical either way.
  CUDA_TRY(cudaMemcpyAsync(h->sched_t, timesteps_bf16, (size_t)rows * 2, cudaMemcpyDeviceToDevice, s));
  h->sched_pass_done.assign((size_t)((rows + kGemvMaxB - 1) / kGemvMaxB), 0);
  finish(h, err_, user);
  API_END
}

int tfx_step_scheduled(tfx_handle h, int32_t step_index, const void* latents_in, const void* cond,
                       const void* encoder_hidden_states, const void* img_ids, const void* txt_ids, float sigma, float sigma_next,
                       void* latents_out, void* noise_pred_out, void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_step_scheduled");
  REQUIRE(h && latents_in && cond && latents_out, TFX_ERR_INVALID, "null argument");
  REQUIRE(h->mod_table && step_index >= 0 && step_index < h->sched_steps, TFX_ERR_STATE,
          "step %d outside the schedule set by tfx_set_schedule (%d steps)", step_index, h->sched_steps);
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  stage_common(h, err_, encoder_hidden_states, nullptr, nullptr, nullptr, img_ids, txt_ids, user, true);
  const int Cl = h->cfg.out_channels, Cc = h->cfg.in_channels - h->cfg.out_channels, Ci = h->cfg.in_channels;
  const size_t rows = (size_t)h->B * h->S;
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaMemcpy2DAsync(h->x_in, (size_t)Ci * 2, latents_in, (size_t)Cl * 2, (size_t)Cl * 2, rows, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpy2DAsync(h->x_in + Cl, (size_t)Ci * 2, cond, (size_t)Cc * 2, (size_t)Cc * 2, rows, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->lat_in, latents_in, rows * Cl * 2, cudaMemcpyDeviceToDevice, s));
  {
    const long long total = (long long)h->sched_steps * h->B, P = h->cfg.pooled_projection_dim;
    const long long first = (long long)step_index * h->B / kGemvMaxB, last = ((long long)(step_index + 1) * h->B - 1) / kGemvMaxB;
    LaunchCtx c{s, h->device, &h->launches, err_};
    for (long long g = first; g <= last; ++g) {
      if (h->sched_pass_done[(size_t)g]) continue;
      const long long r0 = g * kGemvMaxB;
      const int nr = (int)((total - r0 < kGemvMaxB) ? total - r0 : kGemvMaxB);
      h->enqueue_modulation(c, h->sched_t + r0, reinterpret_cast<const float*>(h->sched_g) + r0, h->sched_pooled + r0 * P, nr,
                            h->mod_table + r0 * h->mod_rows);
      h->sched_pass_done[(size_t)g] = 1;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(h->mod, h->mod_table + (size_t)step_index * h->B * h->mod_rows, (size_t)h->B * h->mod_rows * 2,
                           cudaMemcpyDeviceToDevice, s));
  const float dt = __bfloat162float(__float2bfloat16_rn(sigma_next - sigma));
  set_float_kernel<<<1, 1, 0, s>>>(h->dt_dev, dt);
  ++h->launches;
  h->run(true, noise_pred_out != nullptr, true);
  CUDA_TRY(cudaMemcpyAsync(latents_out, h->lat_out, rows * Cl * 2, cudaMemcpyDeviceToDevice, s));
  if (noise_pred_out) CUDA_TRY(cudaMemcpyAsync(noise_pred_out, h->out_buf, rows * Cl * 2, cudaMemcpyDeviceToDevice, s));
  finish(h, err_, user);
  API_END
}

int tfx_euler_step(const void* model_output, const void* sample, void* prev_sample, int64_t n, float sigma, float sigma_next,
                   void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(model_output && sample && prev_sample && n >= 0, TFX_ERR_INVALID, "bad argument");
    if (n == 0) return TFX_OK;
    const float dt = __bfloat162float(__float2bfloat16_rn(sigma_next - sigma));
    euler_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(model_output), reinterpret_cast<const bf16*>(sample), reinterpret_cast<bf16*>(prev_sample), n, dt);
    CUDA_TRY(cudaGetLastError());
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_overshoot_step(const void* model_output, const void* sample, const void* noise_f32, void* prev_sample,
                       void* predicted_x1_f32, int64_t n, float t_overshoot_minus_t, float a, float b, float sigma, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(model_output && sample && noise_f32 && prev_sample && n >= 0, TFX_ERR_INVALID, "bad argument");
    if (n == 0) return TFX_OK;
    const float coef = __bfloat162float(__float2bfloat16_rn(t_overshoot_minus_t));
    const float sg = __bfloat162float(__float2bfloat16_rn(sigma));
    overshoot_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(model_output), reinterpret_cast<const bf16*>(sample), reinterpret_cast<const float*>(noise_f32),
        reinterpret_cast<bf16*>(prev_sample), reinterpret_cast<float*>(predicted_x1_f32), n, coef, a, b, sg);
    CUDA_TRY(cudaGetLastError());
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

// ------------------------------------------------------------------------------------------------ op-level entry points
static long long g_op_launches = 0;

int tfx_op_linear(const void* A, int64_t lda, const void* Wt, const void* bias, void* out, int64_t ldo, int32_t M, int32_t N,
                  int32_t K, int32_t mode, const void* gate, const void* res, int32_t cta_group, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(A && Wt && bias && out, TFX_ERR_INVALID, "null argument");
    REQUIRE(mode >= 0 && mode <= 2, TFX_ERR_INVALID, "mode must be 0..2");
    const int pn = cta_group >= 20 ? cta_group - 20 : 0;  // 22 / 24: multicast kernel with 2 / 4 pairs per cluster
    REQUIRE(cta_group == 1 || cta_group == 2 || pn == 1 || pn == 2 || pn == 4, TFX_ERR_INVALID, "cta_group must be 1, 2, 21, 22 or 24");
    REQUIRE(mode != EPI_GATE_RES || (gate && res), TFX_ERR_INVALID, "gate/res required for mode 2");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    configure_kernels(err_);
    if (M == 0) return TFX_OK;
    const int cg = pn ? 2 : cta_group;
    const char* force = getenv("TFX_OP_LINEAR_BLOCK_N");  // tests pin a tile width through this
    int bn = force ? atoi(force) : pick_block_n((M + 128 * cg - 1) / (128 * cg), N, num_sms(dev) / cg, true);
    if (pn && bn == 192) bn = 256;
    REQUIRE(bn == 256 || bn == 224 || bn == 192, TFX_ERR_INVALID, "TFX_OP_LINEAR_BLOCK_N must be 256, 224 or 192");
    CUtensorMap ma = make_map_2d(err_, A, M, K, lda, pn ? 128 / pn : 128);
    CUtensorMap mb = make_map_2d(err_, Wt, N, K, K, bn / cg);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.K = K; p.num_groups = 1; p.n_split = N; p.mode0 = mode; p.mode1 = mode;
    p.g[0].M = M; p.g[0].rows_per_sample = M; p.g[0].bias = reinterpret_cast<const bf16*>(bias);
    p.g[0].out = reinterpret_cast<bf16*>(out); p.g[0].ldo = ldo;
    p.g[0].res = reinterpret_cast<const bf16*>(res); p.g[0].ldr = ldo;
    p.g[0].gate = reinterpret_cast<const bf16*>(gate); p.g[0].gate_stride = 0;
    if (const char* e = getenv("TFX_GEMM_DEBUG_FLAGS")) p.debug_flags = atoi(e);
    LaunchCtx c{reinterpret_cast<cudaStream_t>(stream), dev, &g_op_launches, err_};
    if (pn) launch_gemm_mc(c, pn, bn, ma, ma, mb, mb, p);
    else launch_gemm(c, cta_group, bn, ma, ma, mb, mb, p);
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_linear_lora(const void* A, int64_t lda, const void* Wt, const void* bias, const void* la, const void* lb, void* t_scratch,
                       void* out, int64_t ldo, int32_t M, int32_t N, int32_t K, int32_t mode, const void* gate, const void* res,
                       int32_t cta_group, int32_t m_band, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(A && Wt && bias && out, TFX_ERR_INVALID, "null argument");
    REQUIRE((la == nullptr) == (lb == nullptr) && (!la || t_scratch), TFX_ERR_INVALID, "la, lb and t_scratch go together");
    REQUIRE(mode >= 0 && mode <= 2, TFX_ERR_INVALID, "mode must be 0..2");
    REQUIRE(cta_group == 1 || cta_group == 2, TFX_ERR_INVALID, "cta_group must be 1 or 2");
    REQUIRE(mode != EPI_GATE_RES || (gate && res), TFX_ERR_INVALID, "gate/res required for mode 2");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    configure_kernels(err_);
    if (M == 0) return TFX_OK;
    const int cg = cta_group;
    const char* force = getenv("TFX_OP_LINEAR_BLOCK_N");
    const int bn = force ? atoi(force) : pick_block_n((M + 128 * cg - 1) / (128 * cg), N, num_sms(dev) / cg, true);
    REQUIRE(bn == 256 || bn == 224 || bn == 192, TFX_ERR_INVALID, "TFX_OP_LINEAR_BLOCK_N must be 256, 224 or 192");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    LaunchCtx c{st, dev, &g_op_launches, err_};
    CUtensorMap ma = make_map_2d(err_, A, M, K, lda, 128);
    CUtensorMap mb = make_map_2d(err_, Wt, N, K, K, bn / cg);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.K = K; p.num_groups = 1; p.n_split = N; p.mode0 = mode; p.mode1 = mode;
    p.k_snake = m_band >= 1000 || m_band <= -1000;  // b +- 1000: band b with the k direction alternating per band (GemmParams::k_snake)
    p.m_band = m_band >= 1000 ? m_band - 1000 : m_band <= -1000 ? m_band + 1000 : m_band;
    p.g[0].M = M; p.g[0].rows_per_sample = M; p.g[0].bias = reinterpret_cast<const bf16*>(bias);
    p.g[0].out = reinterpret_cast<bf16*>(out); p.g[0].ldo = ldo;
    p.g[0].res = reinterpret_cast<const bf16*>(res); p.g[0].ldr = ldo;
    p.g[0].gate = reinterpret_cast<const bf16*>(gate); p.g[0].gate_stride = 0;
    if (!la) {
      launch_gemm(c, cg, bn, ma, ma, mb, mb, p);
      return TFX_OK;
    }
    static std::map<int, bf16*> zero_on;  // per device: 64 zeros, the bias of the T GEMM
    bf16*& zero = zero_on[dev];
    if (!zero) {
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&zero), 256));
      CUDA_TRY(cudaMemset(zero, 0, 256));
    }
    CUtensorMap mla = make_map_2d(err_, la, 64, K, K, 64);
    GemmParams pt;
    memset(&pt, 0, sizeof pt);
    pt.N = 64; pt.K = K; pt.num_groups = 1; pt.n_split = 64; pt.mode0 = pt.mode1 = EPI_STORE;
    pt.g[0].M = M; pt.g[0].rows_per_sample = M; pt.g[0].bias = zero;
    pt.g[0].out = reinterpret_cast<bf16*>(t_scratch); pt.g[0].ldo = 64;
    launch_gemm(c, 1, 64, ma, ma, mla, mla, pt);
    CUtensorMap mt = make_map_2d(err_, t_scratch, M, 64, 64, 128);
    CUtensorMap mlb = make_map_2d(err_, lb, N, 64, 64, bn / cg);
    p.k_ext = kGemmBlockK;
    GemmExt x{&mt, &mt, &mlb, &mlb};
    launch_gemm(c, cg, bn, ma, ma, mb, mb, p, &x);
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_linear_qkv(const void* A, int64_t lda, const void* Wt, const void* bias, const void* rms_q, const void* rms_k,
                      const void* rope_f32, void* q, void* k, void* v, int32_t M, int32_t K, int32_t H, int32_t head_dim,
                      int32_t rows_per_sample, int32_t pos_offset, int32_t n_joint, int32_t cta_group, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(A && Wt && bias && rms_q && rms_k && rope_f32 && q && k && v, TFX_ERR_INVALID, "null argument");
    REQUIRE(cta_group == 1 || cta_group == 2, TFX_ERR_INVALID, "cta_group must be 1 or 2");
    REQUIRE(head_dim == 64 || head_dim == 128, TFX_ERR_INVALID, "head_dim %d unsupported (64 or 128)", head_dim);
    REQUIRE((H * head_dim) % 256 == 0, TFX_ERR_INVALID, "H * head_dim = %d must be a multiple of 256", H * head_dim);
    REQUIRE(rows_per_sample > 0 && pos_offset >= 0 && pos_offset + rows_per_sample <= n_joint, TFX_ERR_INVALID,
            "rows_per_sample %d at offset %d does not fit n_joint %d", rows_per_sample, pos_offset, n_joint);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    configure_kernels(err_);
    if (M == 0) return TFX_OK;
    const int D = H * head_dim, N = 3 * D;
    CUtensorMap ma = make_map_2d(err_, A, M, K, lda, 128);
    CUtensorMap mb = make_map_2d(err_, Wt, N, K, K, 256 / cta_group);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.K = K; p.num_groups = 1; p.n_split = N; p.mode0 = EPI_QKV; p.mode1 = EPI_QKV;
    p.D = D; p.head_dim = head_dim; p.num_heads = H; p.n_joint = n_joint;
    p.q = reinterpret_cast<bf16*>(q); p.k = reinterpret_cast<bf16*>(k); p.v = reinterpret_cast<bf16*>(v);
    p.rope = reinterpret_cast<const float2*>(rope_f32); p.rms_eps = 1e-6f;
    p.g[0].M = M; p.g[0].rows_per_sample = rows_per_sample; p.g[0].pos_offset = pos_offset;
    p.g[0].bias = reinterpret_cast<const bf16*>(bias);
    p.g[0].rms_q = reinterpret_cast<const bf16*>(rms_q); p.g[0].rms_k = reinterpret_cast<const bf16*>(rms_k);
    LaunchCtx c{reinterpret_cast<cudaStream_t>(stream), dev, &g_op_launches, err_};
    launch_gemm(c, cta_group, 256, ma, ma, mb, mb, p);
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_linear_euler(const void* A, int64_t lda, const void* Wt, const void* bias, const void* latents_in, const void* dt_f32_dev,
                        void* noise_pred_out, void* latents_out, int32_t M, int32_t N, int32_t K, int32_t cta_group, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(A && Wt && bias && latents_in && dt_f32_dev && latents_out, TFX_ERR_INVALID, "null argument");
    REQUIRE(cta_group == 1 || cta_group == 2, TFX_ERR_INVALID, "cta_group must be 1 or 2");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    configure_kernels(err_);
    if (M == 0) return TFX_OK;
    CUtensorMap ma = make_map_2d(err_, A, M, K, lda, 128);
    CUtensorMap mb = make_map_2d(err_, Wt, N, K, K, 256 / cta_group);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.K = K; p.num_groups = 1; p.n_split = N; p.mode0 = EPI_EULER; p.mode1 = EPI_EULER;
    p.dt_ptr = reinterpret_cast<const float*>(dt_f32_dev);
    p.g[0].M = M; p.g[0].rows_per_sample = M; p.g[0].bias = reinterpret_cast<const bf16*>(bias);
    p.g[0].out = reinterpret_cast<bf16*>(noise_pred_out); p.g[0].ldo = N;
    p.g[0].res = reinterpret_cast<const bf16*>(latents_in); p.g[0].ldr = N;
    p.g[0].out2 = reinterpret_cast<bf16*>(latents_out);
    LaunchCtx c{reinterpret_cast<cudaStream_t>(stream), dev, &g_op_launches, err_};
    launch_gemm(c, cta_group, 256, ma, ma, mb, mb, p);
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

static void* g_attn_trace = nullptr;
static void* g_attn_cta_trace = nullptr;
int tfx_debug_set_attention_trace(void* dev_ptr) {
  g_attn_trace = dev_ptr;
  return TFX_OK;
}
int tfx_debug_set_attention_cta_trace(void* dev_ptr) {
  g_attn_cta_trace = dev_ptr;
  return TFX_OK;
}

int tfx_op_attention(const void* q, const void* k, const void* v, void* out, int64_t ld_out, int32_t B, int32_t H, int32_t T,
                     int32_t S, int32_t head_dim, int32_t q_tiles, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(q && k && v && out, TFX_ERR_INVALID, "null argument");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    configure_kernels(err_);
    const int N = T + S;
    if (N == 0 || B == 0) return TFX_OK;
    CUtensorMap mq = make_map_3d(err_, q, (long long)B * H, N, head_dim);
    CUtensorMap mk = make_map_3d(err_, k, (long long)B * H, N, head_dim);
    CUtensorMap mv = make_map_3d(err_, v, (long long)B * H, N, head_dim);
    AttnParams p;
    p.B = B; p.H = H; p.N = N; p.T = T; p.S = S;
    p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
    p.out = reinterpret_cast<bf16*>(out); p.ld_out = ld_out;
    LaunchCtx c{reinterpret_cast<cudaStream_t>(stream), dev, &g_op_launches, err_};
    const int sched = q_tiles % 10, emu = (q_tiles / 10) % 10;
    const bool trace_build = (q_tiles / 100) % 10 != 0;
    if (sched == 9 || sched == 7) {  // schedule 4 (stream) | schedule 5 (persistent stream): + 10 * emu
      static std::map<int, Attn4Workspace> ws_on;  // per device, sized for the largest head_dim
      Attn4Workspace& ws = ws_on[dev];
      if (!ws.ws_o) {
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ws.ws_o), Attn4Workspace::o_bytes(128)));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ws.ws_ml), Attn4Workspace::ml_bytes()));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ws.counters), Attn4Workspace::counter_bytes()));
        CUDA_TRY(cudaMemset(ws.counters, 0, Attn4Workspace::counter_bytes()));
      }
      p.cta_trace = reinterpret_cast<long long*>(g_attn_cta_trace);
      if (sched == 9) launch_attention4(c, head_dim, emu, mq, mk, mv, p, ws);
      else launch_attention5(c, head_dim, emu, mq, mk, mv, p, ws, trace_build);
    } else if (sched == 5 || sched == 6) {  // schedule 3: 5 = whole-P hand-over, 6 = split; + 10 * emu; + 100 trace
      p.trace = reinterpret_cast<long long*>(g_attn_trace);
      p.cta_trace = reinterpret_cast<long long*>(g_attn_cta_trace);
      launch_attention3(c, head_dim, emu, sched == 6, trace_build, mq, mk, mv, p);
#ifdef TFX_ATTN8
    } else if (sched == 8) {
      launch_attention8(c, head_dim, emu, mq, mk, mv, p);
#endif
#ifdef TFX_ATTN11
    } else if (sched == 2) {
      launch_attention11(c, head_dim, emu, mq, mk, mv, p);
#endif
#ifdef TFX_ATTN10
    } else if (sched == 3) {
      launch_attention10(c, head_dim, emu, mq, mk, mv, p);
#endif
#ifdef TFX_ATTN9
    } else if (sched == 4) {
      if (trace_build) p.trace = reinterpret_cast<long long*>(g_attn_trace);
      launch_attention9(c, head_dim, emu, mq, mk, mv, p);
#endif
    } else {
      REQUIRE(false, TFX_ERR_INVALID, "attention schedule code %d unknown (5 | 6 | 7 | 9, + 10 * emu)", q_tiles);
    }
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_ln_modulate(const void* x, void* y, int32_t rows, int32_t D, int32_t rows_per_sample, const void* mod,
                       int64_t mod_stride, int64_t shift_off, int64_t scale_off, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(x && y && mod && rows_per_sample > 0, TFX_ERR_INVALID, "bad argument");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    LnModParams p;
    p.x = reinterpret_cast<const bf16*>(x); p.y = reinterpret_cast<bf16*>(y); p.rows = rows; p.D = D; p.row_begin = 0;
    p.rows0 = rows; p.rows_per0 = rows_per_sample; p.rows_per1 = rows_per_sample;
    p.mod = reinterpret_cast<const bf16*>(mod); p.mod_stride = mod_stride;
    p.shift0 = p.shift1 = shift_off; p.scale0 = p.scale1 = scale_off; p.eps = 1e-6f;
    LaunchCtx c{reinterpret_cast<cudaStream_t>(stream), dev, &g_op_launches, err_};
    launch_ln_modulate(c, p);
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_gemv(const void* x, int32_t B, int32_t K, const void* Wt, const void* bias, int64_t N, void* out, int32_t flags,
                void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(x && Wt && bias && out, TFX_ERR_INVALID, "null argument");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    LaunchCtx c{reinterpret_cast<cudaStream_t>(stream), dev, &g_op_launches, err_};
    launch_gemv(c, reinterpret_cast<const bf16*>(x), B, K, reinterpret_cast<const bf16*>(Wt), reinterpret_cast<const bf16*>(bias), N,
                reinterpret_cast<bf16*>(out), flags);
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_rope_table(const void* txt_ids, const void* img_ids, int32_t T, int32_t S, const int32_t* axes_dims, void* out_f32,
                      void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(txt_ids && img_ids && axes_dims && out_f32, TFX_ERR_INVALID, "null argument");
    RopeParams rp;
    rp.txt_ids = reinterpret_cast<const bf16*>(txt_ids); rp.img_ids = reinterpret_cast<const bf16*>(img_ids);
    rp.T = T; rp.S = S;
    rp.axes[0] = axes_dims[0]; rp.axes[1] = axes_dims[1]; rp.axes[2] = axes_dims[2];
    rp.half_dim = (axes_dims[0] + axes_dims[1] + axes_dims[2]) / 2;
    rp.out = reinterpret_cast<float2*>(out_f32);
    const int total = (T + S) * rp.half_dim;
    if (total == 0) return TFX_OK;
    rope_table_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rp);
    CUDA_TRY(cudaGetLastError());
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_timestep_embed(const void* t, int32_t is_f32, int32_t B, void* out, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(t && out && B > 0, TFX_ERR_INVALID, "bad argument");
    timestep_embed_kernel<<<B, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, is_f32, B, reinterpret_cast<bf16*>(out));
    CUDA_TRY(cudaGetLastError());
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

static int cond_blocks(long long total) {
  long long b = (total + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));  // grid-stride, at most 16 blocks per SM
}

int tfx_op_pack_latents(const void* src, int32_t src_is_f32, void* dst, int64_t dst_ld, int64_t dst_off, int32_t B, int32_t C,
                        int32_t h, int32_t w, int32_t affine, float shift, float scale, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(src && dst && B > 0 && C > 0, TFX_ERR_INVALID, "bad argument");
    REQUIRE(h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, TFX_ERR_INVALID, "latent height %d and width %d must be even (2x2 packing)", h, w);
    REQUIRE(dst_ld >= dst_off + 4LL * C && dst_ld % 2 == 0 && dst_off % 2 == 0, TFX_ERR_INVALID, "packed row stride %lld too small for offset %lld + %d channels",
            (long long)dst_ld, (long long)dst_off, 4 * C);
    const long long total = (long long)B * (h / 2) * (w / 2) * C * 2;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (src_is_f32)
      pack_latents_kernel<float><<<cond_blocks(total), 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(dst), dst_ld, dst_off, B, C, h, w, affine, shift, scale);
    else
      pack_latents_kernel<bf16><<<cond_blocks(total), 256, 0, st>>>(reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(dst), dst_ld, dst_off, B, C, h, w, affine, shift, scale);
    CUDA_TRY(cudaGetLastError());
    ++g_op_launches;
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_unpack_latents(const void* src, int64_t src_ld, void* dst, int32_t B, int32_t C, int32_t h, int32_t w, int32_t affine,
                          float shift, float scale, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(src && dst && B > 0 && C > 0, TFX_ERR_INVALID, "bad argument");
    REQUIRE(h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, TFX_ERR_INVALID, "latent height %d and width %d must be even (2x2 packing)", h, w);
    REQUIRE(src_ld >= 4LL * C && src_ld % 2 == 0, TFX_ERR_INVALID, "packed row stride %lld too small for %d channels", (long long)src_ld, 4 * C);
    REQUIRE(!affine || scale != 0.f, TFX_ERR_INVALID, "scaling_factor must not be 0");
    const long long total = (long long)B * C * h * (w / 2);
    unpack_latents_kernel<<<cond_blocks(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(src), src_ld, reinterpret_cast<bf16*>(dst), B, C, h, w, affine, shift, scale);
    CUDA_TRY(cudaGetLastError());
    ++g_op_launches;
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_pack_mask(const void* mask, int32_t mask_is_f32, void* dst, int64_t dst_ld, int64_t dst_off, int32_t B, int32_t h,
                     int32_t w, int32_t vae_scale_factor, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(mask && dst && B > 0, TFX_ERR_INVALID, "bad argument");
    REQUIRE(h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, TFX_ERR_INVALID, "latent height %d and width %d must be even (2x2 packing)", h, w);
    REQUIRE(vae_scale_factor > 0 && vae_scale_factor <= 16, TFX_ERR_INVALID, "vae_scale_factor %d unsupported", vae_scale_factor);
    const int ch = vae_scale_factor * vae_scale_factor;
    REQUIRE(dst_ld >= dst_off + 4LL * ch && dst_ld % 2 == 0 && dst_off % 2 == 0, TFX_ERR_INVALID, "packed row stride %lld too small for offset %lld + %d mask channels",
            (long long)dst_ld, (long long)dst_off, 4 * ch);
    const long long total = (long long)B * (h / 2) * (w / 2) * ch * 2;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (mask_is_f32)
      pack_mask_kernel<float><<<cond_blocks(total), 256, 0, st>>>(reinterpret_cast<const float*>(mask), reinterpret_cast<bf16*>(dst), dst_ld, dst_off, B, h, w, vae_scale_factor);
    else
      pack_mask_kernel<bf16><<<cond_blocks(total), 256, 0, st>>>(reinterpret_cast<const bf16*>(mask), reinterpret_cast<bf16*>(dst), dst_ld, dst_off, B, h, w, vae_scale_factor);
    CUDA_TRY(cudaGetLastError());
    ++g_op_launches;
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

int tfx_op_umma_probe(const void* A, const void* Bm, void* D_f32, int32_t n_dim, int32_t k_dim, int32_t b_mn_major,
                      int32_t a_from_tmem, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep_bytes, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(A && Bm && D_f32, TFX_ERR_INVALID, "null argument");
    REQUIRE(k_dim % 64 == 0 && k_dim >= 64 && k_dim <= 256 && n_dim % 64 == 0 && n_dim >= 64 && n_dim <= 256, TFX_ERR_INVALID,
            "probe supports n,k in {64,128,192,256}");
    configure_kernels(err_);
    CUtensorMap ma = make_map_2d(err_, A, 128, k_dim, k_dim, 128);
    CUtensorMap mb = b_mn_major ? make_map_2d(err_, Bm, k_dim, n_dim, n_dim, k_dim) : make_map_2d(err_, Bm, n_dim, k_dim, k_dim, n_dim);
    ProbeParams p;
    p.A = reinterpret_cast<const bf16*>(A); p.D = reinterpret_cast<float*>(D_f32); p.n = n_dim; p.k = k_dim;
    p.b_mn_major = b_mn_major; p.a_from_tmem = a_from_tmem; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_kstep_bytes = b_kstep_bytes;
    umma_probe_kernel<<<1, 128, 200 * 1024, reinterpret_cast<cudaStream_t>(stream)>>>(ma, mb, p);
    CUDA_TRY(cudaGetLastError());
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

}  // extern "C"

#include "vae_host.inl"
#include "textenc_host.inl"
