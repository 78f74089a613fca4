// The prompt encoders behind the C ABI (include/textflux_b200.h, tfx_textenc_*): launch sequences of transformers' T5EncoderModel
// (modeling_t5.py: T5Stack -> T5Block -> T5LayerSelfAttention / T5LayerFF) and CLIPTextModel (modeling_clip.py: CLIPTextTransformer),
// as FluxFillPipeline._get_t5_prompt_embeds / _get_clip_prompt_embeds call them (pipeline_flux_fill.py:1411-1503).  Textually
// included by tfx_api.cu; kernels in textenc.cuh, every Linear is the tcgen05 GEMM.
struct tfx_textenc {
  tfx_textenc_config cfg;
  int device = 0;
  std::string err;
  std::string* err_ = &err;
  long long launches = 0;
  std::map<std::string, Weight> w;
  // workspace for R = B * T rows
  long long rows_cap = 0;
  bf16 *hidden = nullptr, *normed = nullptr, *qkv = nullptr, *attn = nullptr, *ff = nullptr;
  bf16 *ones = nullptr, *zeros = nullptr;  // [max(3 * inner, d_ff, d_model)]
  cudaStream_t stream = nullptr;
  int inner() const { return cfg.num_heads * 64; }

  const Weight& Wt(const std::string& name, long long rows, long long cols) {
    auto it = w.find(name);
    REQUIRE(it != w.end(), TFX_ERR_MISSING, "text-encoder weight '%s' was never set", name.c_str());
    REQUIRE(it->second.rows == rows && it->second.cols == cols, TFX_ERR_INVALID, "text-encoder weight '%s' is [%lld,%lld], expected [%lld,%lld]",
            name.c_str(), it->second.rows, it->second.cols, rows, cols);
    return it->second;
  }
  const bf16* bias_or_zero(const std::string& name, long long n) {
    if (cfg.kind == TFX_TEXTENC_T5) return zeros;  // T5's Linear layers have no bias
    return Wt(name, 1, n).ptr;
  }
  void release() {
    for (bf16** q : {&hidden, &normed, &qkv, &attn, &ff}) { if (*q) cudaFree(*q); *q = nullptr; }
    rows_cap = 0;
  }
  void reserve(long long rows) {
    if (rows <= rows_cap) return;
    release();
    const long long D = cfg.d_model, I = inner(), F = cfg.d_ff, slack = 64;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&hidden), (size_t)rows * D * 2 + 256));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&normed), (size_t)rows * D * 2 + 256));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&qkv), (size_t)(rows + slack) * 3 * I * 2 + 256));
    CUDA_TRY(cudaMemset(qkv, 0, (size_t)(rows + slack) * 3 * I * 2));  // rows past the last sample are read (and masked): keep them finite
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&attn), (size_t)rows * I * 2 + 256));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ff), (size_t)rows * F * 2 + 256));
    rows_cap = rows;
  }
  void linear(const bf16* x, long long rows, int K, bf16* y, int N, const std::string& wname, const bf16* bias, int mode, const bf16* res) {
    const Weight& wt = Wt(wname, N, K);
    const int bn = N >= 256 ? 256 : (N >= 128 ? 128 : 64), cg = bn == 128 ? 1 : 2;  // 128-wide outputs: single-CTA tiles (vae_host.inl)
    CUtensorMap ma = make_map_2d(err_, x, rows, K, K, 128);
    CUtensorMap mb = make_map_2d(err_, wt.ptr, N, K, K, bn / cg);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.K = K; p.num_groups = 1; p.n_split = N; p.mode0 = p.mode1 = mode;
    p.g[0].M = (int)rows; p.g[0].rows_per_sample = (int)rows + 1; p.g[0].bias = bias; p.g[0].out = y; p.g[0].ldo = N;
    p.g[0].res = res; p.g[0].ldr = N; p.g[0].gate = ones; p.g[0].gate_stride = 0;
    LaunchCtx c{stream, device, &launches, err_};
    launch_gemm(c, cg, bn, ma, ma, mb, mb, p);
  }
  void norm(const bf16* x, bf16* y, long long rows, const std::string& name) {
    const int D = cfg.d_model;
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    if (cfg.kind == TFX_TEXTENC_T5)
      norm_rows_kernel<true><<<blocks, 256, 0, stream>>>(x, Wt(name + ".w", 1, D).ptr, nullptr, y, (int)rows, D, cfg.eps);
    else
      norm_rows_kernel<false><<<blocks, 256, 0, stream>>>(x, Wt(name + ".w", 1, D).ptr, Wt(name + ".b", 1, D).ptr, y, (int)rows, D, cfg.eps);
    CUDA_TRY(cudaGetLastError());
    ++launches;
  }

  void encode(const int32_t* ids, int B, int T, const int32_t* rel_lut, bf16* last_hidden, const int32_t* pooled_index, bf16* pooled) {
    const long long R = (long long)B * T;
    const int D = cfg.d_model, I = inner(), F = cfg.d_ff, H = cfg.num_heads, Tp = (T + 15) / 16 * 16;
    const bool t5 = cfg.kind == TFX_TEXTENC_T5;
    REQUIRE(T >= 1 && Tp <= 512, TFX_ERR_INVALID, "sequence length %d unsupported (1..512)", T);
    REQUIRE(!t5 || rel_lut, TFX_ERR_INVALID, "the T5 encoder needs the relative-position bucket table");
    REQUIRE(t5 || T <= cfg.max_positions, TFX_ERR_INVALID, "sequence length %d exceeds max_position_embeddings %d", T, cfg.max_positions);
    reserve(R);
    const size_t smem = small_attn_smem(Tp);
    embed_rows_kernel<<<(unsigned)std::min<long long>((R * (D / 8) + 255) / 256, 148 * 16), 256, 0, stream>>>(
        ids, Wt("embed", cfg.vocab_size, D).ptr, t5 ? nullptr : Wt("pos", cfg.max_positions, D).ptr, hidden, (int)R, T, D, cfg.vocab_size);
    ++launches;
    char nm[64];
    for (int l = 0; l < cfg.num_layers; ++l) {
      snprintf(nm, sizeof nm, "l%d.", l);
      const std::string L(nm);
      // ---- self-attention sub-layer: x + o(attention(q, k, v of norm(x)))
      norm(hidden, normed, R, L + "ln1");
      linear(normed, R, D, qkv, 3 * I, L + "qkv.w", bias_or_zero(L + "qkv.b", 3 * I), EPI_STORE, nullptr);
      SmallAttnParams ap;
      ap.q = qkv; ap.k = qkv + I; ap.v = qkv + 2 * I; ap.ld = 3LL * I; ap.out = attn; ap.ld_out = I;
      ap.B = B; ap.H = H; ap.T = T; ap.Tp = Tp; ap.scale = 0.125f;  // head_dim^-0.5 (CLIP); T5 folds the scale into its weights
      ap.rel_table = t5 ? Wt("rel_bias", cfg.rel_buckets, H).ptr : nullptr; ap.rel_lut = rel_lut;
      const dim3 grid((T + kSmallAttnRows - 1) / kSmallAttnRows, H, B);
      if (t5) small_attention_kernel<false><<<grid, 128, smem, stream>>>(ap);
      else small_attention_kernel<true><<<grid, 128, smem, stream>>>(ap);
      CUDA_TRY(cudaGetLastError());
      ++launches;
      linear(attn, R, I, hidden, D, L + "o.w", bias_or_zero(L + "o.b", D), EPI_GATE_RES, hidden);
      // ---- feed-forward sub-layer
      norm(hidden, normed, R, L + "ln2");
      if (t5) {  // T5DenseGatedActDense: wo(gelu_new(wi_0 x) * wi_1 x)
        linear(normed, R, D, ff, F, L + "wi0.w", zeros, EPI_GELU, nullptr);
        linear(normed, R, D, ff, F, L + "wi1.w", zeros, EPI_MUL, ff);
        linear(ff, R, F, hidden, D, L + "wo.w", zeros, EPI_GATE_RES, hidden);
      } else {  // CLIPMLP: fc2(quick_gelu(fc1 x))
        linear(normed, R, D, ff, F, L + "fc1.w", Wt(L + "fc1.b", 1, F).ptr, EPI_QUICK_GELU, nullptr);
        linear(ff, R, F, hidden, D, L + "fc2.w", Wt(L + "fc2.b", 1, D).ptr, EPI_GATE_RES, hidden);
      }
    }
    norm(hidden, last_hidden, R, "final_ln");
    if (pooled) {  // CLIPTextTransformer: last_hidden_state[b, eos position of sample b]
      REQUIRE(pooled_index, TFX_ERR_INVALID, "pooled output needs the per-sample token index");
      std::vector<int32_t> idx(B);
      CUDA_TRY(cudaMemcpyAsync(idx.data(), pooled_index, (size_t)B * 4, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      for (int b = 0; b < B; ++b) {
        REQUIRE(idx[b] >= 0 && idx[b] < T, TFX_ERR_INVALID, "pooled index %d of sample %d outside the sequence", idx[b], b);
        CUDA_TRY(cudaMemcpyAsync(pooled + (long long)b * D, last_hidden + ((long long)b * T + idx[b]) * D, (size_t)D * 2, cudaMemcpyDeviceToDevice, stream));
      }
    }
  }
};

extern "C" {

int tfx_textenc_create(const tfx_textenc_config* cfg, int32_t device, tfx_textenc_handle* out) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(cfg && out, TFX_ERR_INVALID, "null argument");
    REQUIRE(cfg->kind == TFX_TEXTENC_T5 || cfg->kind == TFX_TEXTENC_CLIP, TFX_ERR_INVALID, "kind must be TFX_TEXTENC_T5 or TFX_TEXTENC_CLIP");
    REQUIRE(cfg->d_model % 256 == 0 && cfg->d_model >= 256, TFX_ERR_INVALID, "d_model %d must be a multiple of 256", cfg->d_model);
    REQUIRE(cfg->d_kv == 64, TFX_ERR_INVALID, "head dimension %d unsupported (64: T5 v1.1 and CLIP-L)", cfg->d_kv);
    REQUIRE(cfg->num_heads >= 1 && cfg->num_layers >= 1 && cfg->d_ff % 64 == 0 && cfg->vocab_size >= 1, TFX_ERR_INVALID, "bad encoder dimensions");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, TFX_ERR_INVALID, "device %d not present (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, TFX_ERR_INVALID, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    configure_kernels(err_);
    CUDA_TRY(cudaFuncSetAttribute(small_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_attn_smem(512)));
    CUDA_TRY(cudaFuncSetAttribute(small_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_attn_smem(512)));
    tfx_textenc* m = new tfx_textenc();
    m->cfg = *cfg;
    m->device = device;
    try {
      CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
      const size_t n = (size_t)std::max(std::max(3 * m->inner(), cfg->d_ff), cfg->d_model);
      std::vector<bf16> one(n, __float2bfloat16(1.0f));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->ones), n * 2));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->zeros), n * 2));
      CUDA_TRY(cudaMemcpy(m->ones, one.data(), n * 2, cudaMemcpyHostToDevice));
      CUDA_TRY(cudaMemset(m->zeros, 0, n * 2));
    } catch (...) {
      tfx_textenc_destroy(m);  // frees whatever was allocated before the failure
      throw;
    }
    *out = m;
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

void tfx_textenc_destroy(tfx_textenc_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->release();
  cudaFree(h->ones);
  cudaFree(h->zeros);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* tfx_textenc_last_error(tfx_textenc_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int tfx_textenc_set_weight(tfx_textenc_handle h, const char* name, const void* dev_ptr, int64_t rows, int64_t cols) {
  API_BEGIN(h)
  REQUIRE(h && name && dev_ptr && rows > 0 && cols > 0, TFX_ERR_INVALID, "bad argument");
  REQUIRE((reinterpret_cast<uintptr_t>(dev_ptr) & 15) == 0, TFX_ERR_INVALID, "weight '%s' is not 16-byte aligned", name);
  Weight t;
  t.ptr = reinterpret_cast<const bf16*>(dev_ptr); t.rows = rows; t.cols = cols;
  h->w[name] = t;
  API_END
}

int tfx_textenc_get_counter(tfx_textenc_handle h, const char* key, int64_t* value) {
  API_BEGIN(h)
  REQUIRE(h && key && value, TFX_ERR_INVALID, "null argument");
  REQUIRE(std::string(key) == "launches", TFX_ERR_INVALID, "unknown counter '%s'", key);
  *value = h->launches;
  API_END
}

int tfx_textenc_encode(tfx_textenc_handle h, const int32_t* input_ids, int32_t B, int32_t T, const int32_t* rel_bucket_lut,
                       void* last_hidden_state, const int32_t* pooled_index, void* pooled_out, void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_textenc_encode");
  REQUIRE(h && input_ids && last_hidden_state && B >= 1, TFX_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(e0, user));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, e0, 0));
  try {
    h->encode(input_ids, B, T, rel_bucket_lut, reinterpret_cast<bf16*>(last_hidden_state), pooled_index, reinterpret_cast<bf16*>(pooled_out));
  } catch (...) { cudaEventDestroy(e0); cudaEventDestroy(e1); throw; }
  CUDA_TRY(cudaEventRecord(e1, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(user, e1, 0));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  API_END
}

}  // extern "C"
