// AutoencoderKL encode / decode behind the C ABI (include/textflux_b200.h, tfx_vae_*): the host-side launch sequence of
// models/autoencoders/vae.py Encoder.forward / Decoder.forward over NHWC bf16 activations.  Textually included by tfx_api.cu
// (shares its launch helpers); kernels in gemm.cuh (implicit-GEMM convolution mode) and vae.cuh.
namespace {

CUtensorMap make_map_nd(std::string* err_, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box) {
  CUtensorMap m;
  memset(&m, 0, sizeof m);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, TFX_ERR_INVALID, "TMA operand must be 16-byte aligned (%p)", ptr);
  CUresult r = get_encode_fn(err_)(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, TFX_ERR_CUDA, "cuTensorMapEncodeTiled(rank %d) failed: %d", rank, (int)r);
  return m;
}

// NHWC image [B, H, W, C] as a [C, W, H, B] tensor with a 64-channel x 16 x 8-pixel box: the A tile of a stride-1 convolution tap
CUtensorMap make_map_conv_s1(std::string* err_, const bf16* x, long long B, long long H, long long W, long long C) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kConvPatchW, (cuuint32_t)kConvPatchH, 1};
  return make_map_nd(err_, x, 4, dims, str, box);
}
// ONE NHWC image [H, W, C] (H, W even) as [C, 2, W/2, 2, H/2]: pixel (2y + py, 2x + px) = coordinate (c, px, x, py, y); a box of
// 16 x 8 (x, y) pairs at fixed parities is the A tile of a stride-2 convolution tap
CUtensorMap make_map_conv_s2(std::string* err_, const bf16* x, long long H, long long W, long long C) {
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)(W / 2), 2, (cuuint64_t)(H / 2)};
  cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2};
  cuuint32_t box[5] = {64, 1, (cuuint32_t)kConvPatchW, 1, (cuuint32_t)kConvPatchH};
  return make_map_nd(err_, x, 5, dims, str, box);
}

int ew_blocks(long long total, int device) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms(device) * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

struct tfx_vae {
  tfx_vae_config cfg;
  int device = 0;
  std::string err;
  std::string* err_ = &err;
  long long launches = 0;
  std::map<std::string, Weight> w;
  bool finalized = false;
  // workspace (grown on demand)
  bf16* buf[4] = {nullptr, nullptr, nullptr, nullptr};
  long long buf_elems = 0;
  bf16 *aq = nullptr, *ak = nullptr, *avt = nullptr, *ao = nullptr, *ap = nullptr;  // mid-block attention: q, k [N, C]; v^T [C, N]; o [N, C]; P chunk
  float* as = nullptr;                                                                // score chunk [rows, N] fp32
  long long attn_tokens = 0, attn_chunk = 0;
  float *gn_partial = nullptr, *gn_stats = nullptr;
  bf16 *ones = nullptr, *zeros = nullptr;  // [512]-vectors: gate of the residual epilogue, bias of the bias-free GEMMs
  cudaStream_t stream = nullptr;

  const Weight& Wt(const std::string& name) {
    auto it = w.find(name);
    REQUIRE(it != w.end(), TFX_ERR_MISSING, "VAE weight '%s' was never set", name.c_str());
    return it->second;
  }
  const Weight& Wt(const std::string& name, long long rows, long long cols) {
    const Weight& t = Wt(name);
    REQUIRE(t.rows == rows && t.cols == cols, TFX_ERR_INVALID, "VAE weight '%s' is [%lld,%lld], expected [%lld,%lld]", name.c_str(), t.rows,
            t.cols, rows, cols);
    return t;
  }
  void release() {
    for (auto& b : buf) { if (b) cudaFree(b); b = nullptr; }
    for (bf16** q : {&aq, &ak, &avt, &ao, &ap}) { if (*q) cudaFree(*q); *q = nullptr; }
    if (as) cudaFree(as); as = nullptr;
    buf_elems = 0; attn_tokens = 0;
  }
  void reserve(long long elems, long long tokens, int C) {
    if (elems > buf_elems) {
      for (auto& b : buf) { if (b) cudaFree(b); b = nullptr; }
      for (auto& b : buf) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&b), (size_t)elems * 2 + 256));
      buf_elems = elems;
    }
    if (cfg.mid_block_add_attention && tokens > attn_tokens) {
      for (bf16** q : {&aq, &ak, &avt, &ao, &ap}) { if (*q) cudaFree(*q); *q = nullptr; }
      if (as) cudaFree(as); as = nullptr;
      attn_chunk = std::min<long long>(tokens, 8192);
      const long long tp = (tokens + 7) / 8 * 8;
      for (bf16** q : {&aq, &ak, &avt, &ao}) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(q), (size_t)tp * C * 2 + 256));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ap), (size_t)attn_chunk * tp * 2 + 256));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&as), (size_t)attn_chunk * tp * 4 + 256));
      attn_tokens = tokens;
    }
  }
  LaunchCtx ctx() { return LaunchCtx{stream, device, &launches, err_}; }

  // ---- layers -------------------------------------------------------------------------------------------------------------
  static int tile_n(int Cout) { return Cout >= 256 ? 256 : (Cout >= 128 ? 128 : 64); }
  // An SS-mode MMA re-reads its 128 x 16 A operand (4 KB) from shared memory at ~64 B/cycle whatever N is (profiles/r2s_attn_dbuf.md):
  // a CTA pair on a 256 x 128 tile does 32 cycles of math per 64-cycle read (ncu: tensor pipe 35-38 % on the Cout = 128 layers), a
  // single CTA on 128 x 128 does 64 -- but then fetches 32 KB of operands per k-block from the L2, twice what the L2 feeds an SM.
  // Measured at 1024 x 1152 (profiles/r2t_final.md): 128-wide outputs 8 % faster on single-CTA tiles, 64-wide ones 30 % slower.
  static int cta_group_for(int bn) { return bn == 128 ? 1 : 2; }

  // y[B, Ho, Wo, Cout] = conv3x3(x[B, H, W, Cin]) + bias (+ res): stride 1 pad 1, or stride 2 with Downsample2D's (0,1,0,1) padding
  void conv3x3(const bf16* x, int B, int H, int W, int Cin, bf16* y, int Cout, const std::string& name, bool stride2, const bf16* res) {
    REQUIRE(Cin % 64 == 0, TFX_ERR_INVALID, "conv '%s': %d input channels (must be a multiple of 64 after padding)", name.c_str(), Cin);
    const Weight& wt = Wt(name + ".weight", Cout, 9LL * Cin);
    const Weight& bs = Wt(name + ".bias", 1, Cout);
    const int Ho = stride2 ? H / 2 : H, Wo = stride2 ? W / 2 : W;
    REQUIRE(!stride2 || (H % 2 == 0 && W % 2 == 0), TFX_ERR_INVALID, "stride-2 convolution needs even height and width (%d x %d)", H, W);
    const int bn = tile_n(Cout), cg = cta_group_for(bn);
    CUtensorMap mb = make_map_2d(err_, wt.ptr, Cout, 9LL * Cin, 9LL * Cin, bn / cg);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = Cout; p.K = 9 * Cin; p.num_groups = 1; p.n_split = Cout;
    p.mode0 = p.mode1 = res ? EPI_GATE_RES : EPI_STORE;
    p.conv.mode = stride2 ? 2 : 1; p.conv.H = Ho; p.conv.W = Wo;
    p.conv.tiles_x = (Wo + kConvPatchW - 1) / kConvPatchW; p.conv.tiles_y = (Ho + kConvPatchH - 1) / kConvPatchH;
    p.conv.cin_blocks = Cin / 64;
    p.g[0].bias = bs.ptr; p.g[0].ldo = Cout; p.g[0].ldr = Cout; p.g[0].gate = ones; p.g[0].gate_stride = 0;
    LaunchCtx c = ctx();
    const int images_per_launch = stride2 ? 1 : B;  // the stride-2 view is a map of one image
    for (int b0 = 0; b0 < B; b0 += images_per_launch) {
      const long long in_off = (long long)b0 * H * W * Cin, out_off = (long long)b0 * Ho * Wo * Cout;
      CUtensorMap ma = stride2 ? make_map_conv_s2(err_, x + in_off, H, W, Cin) : make_map_conv_s1(err_, x, B, H, W, Cin);
      p.conv.n_patches = images_per_launch * p.conv.tiles_x * p.conv.tiles_y;
      p.g[0].M = images_per_launch * Ho * Wo; p.g[0].rows_per_sample = p.g[0].M + 1;
      p.g[0].out = y + out_off; p.g[0].res = res ? res + out_off : nullptr;
      launch_gemm(c, cg, bn, ma, ma, mb, mb, p);
    }
  }

  // y[rows, N] = epilogue(x[rows, K] W^T + bias): 1x1 convolutions and the attention projections on NHWC rows
  void linear(const bf16* x, long long lda, long long rows, int K, bf16* y, long long ldo, int N, const bf16* wt, const bf16* bias, int mode,
              const bf16* res, long long ldw = 0) {
    const int bn = tile_n(N), cg = cta_group_for(bn);
    CUtensorMap ma = make_map_2d(err_, x, rows, K, lda, 128);
    CUtensorMap mb = make_map_2d(err_, wt, N, K, ldw ? ldw : K, bn / cg);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.K = K; p.num_groups = 1; p.n_split = N; p.mode0 = p.mode1 = mode;
    p.g[0].M = (int)rows; p.g[0].rows_per_sample = (int)rows + 1; p.g[0].bias = bias; p.g[0].out = y; p.g[0].ldo = ldo;
    p.g[0].res = res; p.g[0].ldr = ldo; p.g[0].gate = ones; p.g[0].gate_stride = 0;
    LaunchCtx c = ctx();
    launch_gemm(c, cg, bn, ma, ma, mb, mb, p);
  }

  void group_norm(const bf16* x, bf16* y, int B, long long HW, int C, const std::string& name, bool act) {
    const int G = cfg.norm_num_groups;
    REQUIRE(C % 64 == 0 && C <= 2048 && kGnThreads % (C / 8) == 0 && C % G == 0 && G <= kGnThreads, TFX_ERR_INVALID,
            "GroupNorm '%s': %d channels in %d groups unsupported", name.c_str(), C, G);
    GnParams p;
    p.x = x; p.y = y; p.B = B; p.C = C; p.G = G; p.HW = HW;
    p.gamma = Wt(name + ".weight", 1, C).ptr; p.beta = Wt(name + ".bias", 1, C).ptr;
    p.eps = 1e-6f; p.silu = act ? 1 : 0; p.partial = gn_partial; p.stats = gn_stats;
    const long long want = (HW + 63) / 64;  // at least 64 pixels per chunk
    p.nchunk = (int)std::max<long long>(1, std::min<long long>(want, std::min<long long>(2048, 8LL * num_sms(device) / std::max(1, B))));
    REQUIRE((long long)B * p.nchunk * G * 2 <= (1 << 20) && (long long)B * G * 2 <= 8192, TFX_ERR_INVALID, "GroupNorm workspace too small");
    gn_partial_kernel<<<dim3(p.nchunk, B), kGnThreads, 0, stream>>>(p);
    gn_finalize_kernel<<<dim3(G, B), 128, 0, stream>>>(p);
    const long long vec_total = HW * (C / 8);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((vec_total + kGnThreads - 1) / kGnThreads, 8LL * num_sms(device)));
    gn_apply_kernel<<<dim3(blocks, B), kGnThreads, 0, stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    launches += 3;
  }

  // ResnetBlock2D.forward (models/resnet.py) with temb = None: x -> x_or_shortcut(x) + conv2(silu(norm2(conv1(silu(norm1(x))))))
  // `cur` indexes buf[]; returns the index holding the result
  int resnet(int cur, int B, int H, int W, int Cin, int Cout, const std::string& pre) {
    const long long HW = (long long)H * W;
    int f[3], n = 0;
    for (int i = 0; i < 4; ++i) if (i != cur) f[n++] = i;
    group_norm(buf[cur], buf[f[0]], B, HW, Cin, pre + ".norm1", true);
    conv3x3(buf[f[0]], B, H, W, Cin, buf[f[1]], Cout, pre + ".conv1", false, nullptr);
    group_norm(buf[f[1]], buf[f[0]], B, HW, Cout, pre + ".norm2", true);
    int skip = cur;
    if (Cin != Cout) {  // conv_shortcut: 1x1 convolution = a GEMM on the pixel rows
      linear(buf[cur], Cin, (long long)B * HW, Cin, buf[f[2]], Cout, Cout, Wt(pre + ".conv_shortcut.weight", Cout, Cin).ptr,
             Wt(pre + ".conv_shortcut.bias", 1, Cout).ptr, EPI_STORE, nullptr);
      skip = f[2];
    }
    conv3x3(buf[f[0]], B, H, W, Cout, buf[skip], Cout, pre + ".conv2", false, buf[skip]);  // in place over the skip operand
    return skip;
  }

  // UNetMidBlock2D's Attention (one head of C channels, AttnProcessor2_0): x + to_out(softmax(q k^T / sqrt(C)) v), q/k/v = Linear(GroupNorm(x))
  int mid_attention(int cur, int B, int H, int W, int C, const std::string& pre) {
    const long long N = (long long)H * W;
    int f[3], n = 0;
    for (int i = 0; i < 4; ++i) if (i != cur) f[n++] = i;
    const long long Np = (N + 7) / 8 * 8;  // row stride of the [*, N] matrices (16-byte rows for TMA; the pad columns are never read)
    group_norm(buf[cur], buf[f[0]], B, N, C, pre + ".group_norm", false);
    const float scale_log2 = (1.0f / sqrtf((float)C)) * 1.4426950408889634f;
    for (int b = 0; b < B; ++b) {
      const bf16* xn = buf[f[0]] + (long long)b * N * C;
      linear(xn, C, N, C, aq, C, C, Wt(pre + ".to_q.weight", C, C).ptr, Wt(pre + ".to_q.bias", 1, C).ptr, EPI_STORE, nullptr);
      linear(xn, C, N, C, ak, C, C, Wt(pre + ".to_k.weight", C, C).ptr, Wt(pre + ".to_k.bias", 1, C).ptr, EPI_STORE, nullptr);
      // v^T [C, N] = W_v xn^T (operands swapped so that v arrives K-major for the P v GEMM); its bias is added after P v instead:
      // rows of P sum to 1, so P (v + 1 b^T) = P v + b^T
      linear(Wt(pre + ".to_v.weight", C, C).ptr, C, C, C, avt, Np, (int)N, xn, zeros_n(N), EPI_STORE, nullptr);
      for (long long r0 = 0; r0 < N; r0 += attn_chunk) {
        const long long rows = std::min(attn_chunk, N - r0);
        linear(aq + r0 * C, C, rows, C, reinterpret_cast<bf16*>(as), Np, (int)N, ak, zeros_n(N), EPI_STORE_F32, nullptr);
        softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(as, Np, ap, Np, (int)N, scale_log2);
        ++launches;
        linear(ap, Np, rows, (int)N, ao + r0 * C, C, C, avt, Wt(pre + ".to_v.bias", 1, C).ptr, EPI_STORE, nullptr, Np);
      }
      // to_out[0] + residual (residual_connection = True, rescale_output_factor = 1), in place over x
      bf16* xb = buf[cur] + (long long)b * N * C;
      linear(ao, C, N, C, xb, C, C, Wt(pre + ".to_out.0.weight", C, C).ptr, Wt(pre + ".to_out.0.bias", 1, C).ptr, EPI_GATE_RES, xb);
    }
    CUDA_TRY(cudaGetLastError());
    return cur;
  }
  // a zero bias of n elements (the score and v^T GEMMs have none): `zeros` is sized for the largest token count at reserve()
  bf16* zeros_big = nullptr;
  long long zeros_big_elems = 0;
  const bf16* zeros_n(long long n) {
    if (n <= 512) return zeros;
    if (n > zeros_big_elems) {
      if (zeros_big) cudaFree(zeros_big);
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&zeros_big), (size_t)n * 2 + 256));
      CUDA_TRY(cudaMemsetAsync(zeros_big, 0, (size_t)n * 2, stream));
      zeros_big_elems = n;
    }
    return zeros_big;
  }

  int mid_block(int cur, int B, int H, int W, int C, const std::string& pre) {
    cur = resnet(cur, B, H, W, C, C, pre + ".resnets.0");
    if (cfg.mid_block_add_attention) cur = mid_attention(cur, B, H, W, C, pre + ".attentions.0");
    return resnet(cur, B, H, W, C, C, pre + ".resnets.1");
  }

  long long max_elems(int B, int H, int W) const {  // largest activation of either direction at image size H x W
    long long m = (long long)B * H * W * 64;
    int h = H, ww = W;
    for (int i = 0; i < cfg.num_blocks; ++i) {
      const long long c = std::max(cfg.block_out_channels[i], i + 1 < cfg.num_blocks ? cfg.block_out_channels[i + 1] : 0);
      m = std::max(m, (long long)B * h * ww * std::max<long long>(c, 64));
      if (i + 1 < cfg.num_blocks) { h /= 2; ww /= 2; }
    }
    return m;
  }

  // Encoder.forward: image [B, in_channels, H, W] (fp32 | bf16) -> moments [B, 2 * latent_channels, H / f, W / f] bf16
  void encode(const void* image, bool is_f32, int B, int H, int W, bf16* moments) {
    const int nb = cfg.num_blocks, f = 1 << (nb - 1);
    REQUIRE(H % f == 0 && W % f == 0 && H >= f && W >= f, TFX_ERR_INVALID, "image %d x %d is not a multiple of the VAE scale factor %d", H, W, f);
    const int Cl = cfg.block_out_channels[nb - 1];
    reserve(max_elems(B, H, W), (long long)(H / f) * (W / f), Cl);
    const long long HW = (long long)H * W;
    if (is_f32) nchw_to_nhwc_kernel<float><<<ew_blocks((long long)B * HW * 8, device), 256, 0, stream>>>(reinterpret_cast<const float*>(image), buf[0], B, cfg.in_channels, HW, 64);
    else nchw_to_nhwc_kernel<bf16><<<ew_blocks((long long)B * HW * 8, device), 256, 0, stream>>>(reinterpret_cast<const bf16*>(image), buf[0], B, cfg.in_channels, HW, 64);
    ++launches;
    int cur = 1, h = H, ww = W, C = cfg.block_out_channels[0];
    conv3x3(buf[0], B, h, ww, 64, buf[cur], C, "encoder.conv_in", false, nullptr);
    char nm[96];
    for (int i = 0; i < nb; ++i) {
      const int Co = cfg.block_out_channels[i];
      for (int j = 0; j < cfg.layers_per_block; ++j) {
        snprintf(nm, sizeof nm, "encoder.down_blocks.%d.resnets.%d", i, j);
        cur = resnet(cur, B, h, ww, C, Co, nm);
        C = Co;
      }
      if (i + 1 < nb) {
        snprintf(nm, sizeof nm, "encoder.down_blocks.%d.downsamplers.0.conv", i);
        const int nxt = (cur + 1) % 4;
        conv3x3(buf[cur], B, h, ww, C, buf[nxt], C, nm, true, nullptr);
        cur = nxt; h /= 2; ww /= 2;
      }
    }
    cur = mid_block(cur, B, h, ww, C, "encoder.mid_block");
    const int a = (cur + 1) % 4, o = (cur + 2) % 4, Cm = 2 * cfg.latent_channels;
    group_norm(buf[cur], buf[a], B, (long long)h * ww, C, "encoder.conv_norm_out", true);
    conv3x3(buf[a], B, h, ww, C, buf[o], Cm, "encoder.conv_out", false, nullptr);
    nhwc_to_nchw_kernel<<<ew_blocks((long long)B * Cm * h * ww, device), 256, 0, stream>>>(buf[o], Cm, moments, B, Cm, (long long)h * ww);
    ++launches;
    CUDA_TRY(cudaGetLastError());
  }

  // Decoder.forward: z [B, latent_channels, h, w] bf16 -> image [B, out_channels, h * f, w * f] bf16
  void decode(const bf16* z, int B, int h, int ww, bf16* image) {
    const int nb = cfg.num_blocks, f = 1 << (nb - 1);
    int C = cfg.block_out_channels[nb - 1];
    reserve(max_elems(B, h * f, ww * f), (long long)h * ww, C);
    nchw_to_nhwc_kernel<bf16><<<ew_blocks((long long)B * h * ww * 8, device), 256, 0, stream>>>(z, buf[0], B, cfg.latent_channels, (long long)h * ww, 64);
    ++launches;
    int cur = 1;
    conv3x3(buf[0], B, h, ww, 64, buf[cur], C, "decoder.conv_in", false, nullptr);
    cur = mid_block(cur, B, h, ww, C, "decoder.mid_block");
    char nm[96];
    for (int i = 0; i < nb; ++i) {
      const int Co = cfg.block_out_channels[nb - 1 - i];
      for (int j = 0; j < cfg.layers_per_block + 1; ++j) {
        snprintf(nm, sizeof nm, "decoder.up_blocks.%d.resnets.%d", i, j);
        cur = resnet(cur, B, h, ww, C, Co, nm);
        C = Co;
      }
      if (i + 1 < nb) {  // Upsample2D: nearest 2x, then a 3x3 convolution
        snprintf(nm, sizeof nm, "decoder.up_blocks.%d.upsamplers.0.conv", i);
        const int up = (cur + 1) % 4, nxt = (cur + 2) % 4;
        upsample_nearest2x_kernel<<<ew_blocks((long long)B * 4 * h * ww * (C / 8), device), 256, 0, stream>>>(buf[cur], buf[up], B, h, ww, C);
        ++launches;
        h *= 2; ww *= 2;
        conv3x3(buf[up], B, h, ww, C, buf[nxt], C, nm, false, nullptr);
        cur = nxt;
      }
    }
    const int a = (cur + 1) % 4, o = (cur + 2) % 4, Co = cfg.out_channels;
    group_norm(buf[cur], buf[a], B, (long long)h * ww, C, "decoder.conv_norm_out", true);
    conv3x3(buf[a], B, h, ww, C, buf[o], Co, "decoder.conv_out", false, nullptr);
    nhwc_to_nchw_kernel<<<ew_blocks((long long)B * Co * h * ww, device), 256, 0, stream>>>(buf[o], Co, image, B, Co, (long long)h * ww);
    ++launches;
    CUDA_TRY(cudaGetLastError());
  }
};

extern "C" {

int tfx_vae_create(const tfx_vae_config* cfg, int32_t device, tfx_vae_handle* out) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(cfg && out, TFX_ERR_INVALID, "null argument");
    REQUIRE(cfg->num_blocks >= 1 && cfg->num_blocks <= 8, TFX_ERR_INVALID, "num_blocks %d unsupported (1..8)", cfg->num_blocks);
    REQUIRE(cfg->in_channels >= 1 && cfg->in_channels <= 64 && cfg->latent_channels >= 1 && cfg->latent_channels <= 32 &&
                cfg->out_channels >= 1 && cfg->out_channels <= 64,
            TFX_ERR_INVALID, "in / out channels must be <= 64 and latent_channels <= 32");
    for (int i = 0; i < cfg->num_blocks; ++i)
      REQUIRE(cfg->block_out_channels[i] % 64 == 0 && cfg->block_out_channels[i] >= 64 && cfg->block_out_channels[i] <= 2048 &&
                  kGnThreads % (cfg->block_out_channels[i] / 8) == 0 && cfg->block_out_channels[i] % cfg->norm_num_groups == 0,
              TFX_ERR_INVALID, "block_out_channels[%d] = %d unsupported (64, 128, 256, 512, 1024 or 2048, divisible by norm_num_groups)", i,
              cfg->block_out_channels[i]);
    REQUIRE(cfg->layers_per_block >= 1 && cfg->norm_num_groups >= 1 && cfg->norm_num_groups <= 256, TFX_ERR_INVALID, "bad layers_per_block / norm_num_groups");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, TFX_ERR_INVALID, "device %d not present (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, TFX_ERR_INVALID, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    configure_kernels(err_);
    tfx_vae* m = new tfx_vae();
    m->cfg = *cfg;
    m->device = device;
    try {
      CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
      std::vector<bf16> one(2048, __float2bfloat16(1.0f));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->ones), 4096));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->zeros), 4096));
      CUDA_TRY(cudaMemcpy(m->ones, one.data(), 4096, cudaMemcpyHostToDevice));
      CUDA_TRY(cudaMemset(m->zeros, 0, 4096));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->gn_partial), (size_t)(1 << 20) * 4));
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&m->gn_stats), 8192 * 4));
    } catch (...) {
      tfx_vae_destroy(m);  // frees whatever was allocated before the failure
      throw;
    }
    *out = m;
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

void tfx_vae_destroy(tfx_vae_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->release();
  for (void* p : {(void*)h->ones, (void*)h->zeros, (void*)h->gn_partial, (void*)h->gn_stats, (void*)h->zeros_big})
    if (p) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* tfx_vae_last_error(tfx_vae_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int tfx_vae_set_weight(tfx_vae_handle h, const char* name, const void* dev_ptr, int64_t rows, int64_t cols) {
  API_BEGIN(h)
  REQUIRE(h && name && dev_ptr && rows > 0 && cols > 0, TFX_ERR_INVALID, "bad argument");
  REQUIRE((reinterpret_cast<uintptr_t>(dev_ptr) & 15) == 0, TFX_ERR_INVALID, "weight '%s' is not 16-byte aligned", name);
  Weight t;
  t.ptr = reinterpret_cast<const bf16*>(dev_ptr); t.rows = rows; t.cols = cols;
  h->w[name] = t;
  API_END
}

int tfx_vae_get_counter(tfx_vae_handle h, const char* key, int64_t* value) {
  API_BEGIN(h)
  REQUIRE(h && key && value, TFX_ERR_INVALID, "null argument");
  REQUIRE(std::string(key) == "launches", TFX_ERR_INVALID, "unknown counter '%s'", key);
  *value = h->launches;
  API_END
}

int tfx_vae_encode(tfx_vae_handle h, const void* image, int32_t image_is_f32, int32_t B, int32_t H, int32_t W, void* moments_out, void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_vae_encode");
  REQUIRE(h && image && moments_out && B >= 1, TFX_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(e0, user));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, e0, 0));
  try {
    h->encode(image, image_is_f32 != 0, B, H, W, reinterpret_cast<bf16*>(moments_out));
  } catch (...) { cudaEventDestroy(e0); cudaEventDestroy(e1); throw; }
  CUDA_TRY(cudaEventRecord(e1, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(user, e1, 0));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  API_END
}

int tfx_vae_decode(tfx_vae_handle h, const void* latents, int32_t B, int32_t lh, int32_t lw, void* image_out, void* stream) {
  API_BEGIN(h)
  NvtxRange nvtx_("tfx_vae_decode");
  REQUIRE(h && latents && image_out && B >= 1 && lh >= 1 && lw >= 1, TFX_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(e0, user));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, e0, 0));
  try {
    h->decode(reinterpret_cast<const bf16*>(latents), B, lh, lw, reinterpret_cast<bf16*>(image_out));
  } catch (...) { cudaEventDestroy(e0); cudaEventDestroy(e1); throw; }
  CUDA_TRY(cudaEventRecord(e1, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(user, e1, 0));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  API_END
}

int tfx_op_gaussian_sample(const void* moments, const void* noise, void* out, int32_t B, int32_t latent_channels, int64_t hw, void* stream) {
  std::string* err_ = nullptr;
  try {
    REQUIRE(moments && noise && out && B >= 1 && latent_channels >= 1 && hw >= 1, TFX_ERR_INVALID, "bad argument");
    const long long total = (long long)B * latent_channels * hw;
    gaussian_sample_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(moments), reinterpret_cast<const bf16*>(noise), reinterpret_cast<bf16*>(out), B, latent_channels, hw);
    CUDA_TRY(cudaGetLastError());
  } catch (const Fail& f) {
    return f.code;
  }
  return TFX_OK;
}

}  // extern "C"
