/* textflux_b200 C ABI: the drop-in boundary for the TextFlux / FLUX-Fill denoising hot path on B200 (sm_100a).
 *
 * The reference has no native code and no FFI; the boundary it exposes for this path is two Python calls inside
 * FluxFillPipeline.__call__ (reference: diffusers/src/diffusers/pipelines/flux/pipeline_flux_fill.py):
 *     :2084-2094   noise_pred = self.transformer(hidden_states=..., timestep=..., guidance=..., ...)[0]
 *     :2098        latents    = self.scheduler.step(noise_pred, t, latents, return_dict=False)[0]
 * Each entry point below names the reference interface it replaces.  Plain pointers and sizes only; every device
 * pointer is bf16 unless stated; `stream` is a cudaStream_t passed as void*.  All functions return 0 on success and a
 * non-zero tfx_status otherwise (never throw); tfx_last_error() gives the message.  A handle is bound to one device
 * and is not re-entrant (the reference is single-threaded per process, one stream).
 */
#ifndef TEXTFLUX_B200_H_
#define TEXTFLUX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfx_model* tfx_handle;

typedef enum {
  TFX_OK = 0,
  TFX_ERR_INVALID = 1, /* bad argument / unsupported shape (the reference raises ValueError) */
  TFX_ERR_STATE = 2,   /* call order violated (weights not finalized, prepare() missing) */
  TFX_ERR_CUDA = 3,    /* CUDA runtime / driver failure */
  TFX_ERR_MISSING = 4  /* a weight the config requires was never set */
} tfx_status;

/* Mirrors FluxTransformer2DModel.__init__'s @register_to_config arguments
 * (models/transformers/transformer_flux.py:865-879). */
typedef struct {
  int32_t in_channels;
  int32_t out_channels;
  int32_t num_layers;
  int32_t num_single_layers;
  int32_t attention_head_dim;
  int32_t num_attention_heads;
  int32_t joint_attention_dim;
  int32_t pooled_projection_dim;
  int32_t guidance_embeds;
  int32_t axes_dims_rope[3];
} tfx_config;

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
/* replaces FluxTransformer2DModel.__init__ (transformer_flux.py:865-921) */
int tfx_create(const tfx_config* cfg, int32_t device, tfx_handle* out);
void tfx_destroy(tfx_handle h);
/* message of the last failure on `h` (or of the last failing call without a handle when h == NULL) */
const char* tfx_last_error(tfx_handle h);
/* "gemm_narrow_tiles" (0|1: allow 224-wide GEMM tiles against wave quantisation), "gemm_m_band" (-1..64: tile order of the wide-K GEMMs, -1 = per shape (default), 0 = M-fastest, b = bands of b M tiles, -100 - n = N bands of n tiles), "gemm_k_snake" (0|1, default 0: on the banded GEMMs the tiles of every second band walk their k-blocks back to front so that consecutive bands meet in the L2), "gemm_l2_hints" (0..3: bit 0 activations evict_last, bit 1 weights evict_first in the GEMM's TMA loads), "gemm_cta_group" (1|2), "gemm_mcast" (0|2|4: CTA pairs per cluster sharing A by TMA multicast), "attn_variant" (0 = per shape, schedule 3 or 5 (default); 4 | 5 = schedule 3 whole-P / split-P, 8 = schedule 4 stream, 9 = schedule 5 persistent stream), "attn_emu" (0|2|3|4), "use_graph" (0|1), "mod_cache_slots" (0..4096: device-side cache of modulation vectors keyed on the (timestep, guidance, pooled) bits tfx_forward / tfx_step receive; 0 = recompute every call), "mod_cache_reset" (drop every cached modulation vector: call after rewriting weights in place), "use_pdl" (0|1: programmatic dependent launch between a step's kernels), "profile" (0|1: eager launches, one CUDA-event
 * pair per kernel, summed per family; resets the sums) */
int tfx_set_option(tfx_handle h, const char* key, int64_t value);
/* "launches": kernels launched by this handle since creation; "graph_nodes": kernel nodes in the captured step;
 * "mod_cache_hits" / "mod_cache_valid": forwards served from the modulation cache / slots in use;
 * "prof_us_<fam>" / "prof_n_<fam>" with fam in gemm|attn|ln|gemv|misc: device microseconds / launches in profile mode */
int tfx_get_counter(tfx_handle h, const char* key, int64_t* value);

/* ---- weights (replaces load_state_dict on the reference module; names listed in INTEGRATION.md) -------------- */
/* Registers a bf16 device tensor [rows, cols] (row-major, 16-byte aligned) under a packed-layout name.
 * The caller keeps the memory alive for the handle's lifetime. */
int tfx_set_weight(tfx_handle h, const char* name, const void* dev_ptr, int64_t rows, int64_t cols);
/* Unfused LoRA, the form the reference runs at inference (run_inference_lora.py:52-65; PEFT Linear.forward:
 * base(x) + scaling * lora_B(lora_A(x))): next to a packed matrix "<m>.w" [N, K] register "<m>.la" [64, K] (the lora_A rows of
 * the modules packed into it, zero padded to 64) and "<m>.lb" [N, 64] (scaling * lora_B of each module in that module's rows
 * and rank columns).  The GEMM of "<m>" then accumulates x W^T + bf16(x la^T) lb^T in one fp32 accumulator before its epilogue;
 * the base weight is not modified.  tfx_unset_weight removes a registered tensor (both ".la" and ".lb" to drop an adapter);
 * call tfx_finalize_weights again after either. */
int tfx_unset_weight(tfx_handle h, const char* name);
int tfx_finalize_weights(tfx_handle h);

/* ---- per-request set-up --------------------------------------------------------------------------------------- */
/* Sizes workspaces and TMA descriptors for batch B, S image tokens, T text tokens per sample. */
int tfx_prepare(tfx_handle h, int32_t B, int32_t S, int32_t T);

/* ---- the hot path ------------------------------------------------------------------------------------------- */
/* replaces FluxTransformer2DModel.forward (transformer_flux.py:1028-1212) as called at pipeline_flux_fill.py:2084.
 *   hidden_states [B,S,in_channels]   encoder_hidden_states [B,T,joint_attention_dim]   pooled [B,pooled_dim]
 *   timestep [B] bf16 (= t/1000, as the pipeline passes it)     guidance [B] fp32 or NULL when !guidance_embeds
 *   img_ids [S,3] bf16   txt_ids [T,3] bf16      out_sample [B,S,out_channels]
 */
int tfx_forward(tfx_handle h, const void* hidden_states, const void* encoder_hidden_states, const void* pooled,
                const void* timestep_bf16, const void* guidance_f32, const void* img_ids, const void* txt_ids,
                void* out_sample, void* stream);

/* replaces FlowMatchEulerDiscreteScheduler.step (schedulers/scheduling_flow_match_euler_discrete.py:265-338):
 *   prev = bf16(fp32(sample) + bf16(bf16(sigma_next - sigma) * model_output)),  n elements. */
int tfx_euler_step(const void* model_output, const void* sample, void* prev_sample, int64_t n, float sigma,
                   float sigma_next, void* stream);

/* replaces StochasticRFOvershotDiscreteScheduler.step, the TextFlux default sampler (demo.py:15, scripts/batch_eval.sh)
 * in its attn_map = None form (schedulers/scheduling_stochastic_rf_discrete_overshot.py:300-366):
 *   x_o = fp32(sample) + bf16(bf16(t_o - t) * (-v));  prev = bf16(x_o * a + noise * b);  x1 = fp32(sample) - bf16(bf16(sigma) * v)
 * The host passes the scalars (t_o - t, a, b, sigma) and the fp32 noise drawn by torch's generator exactly where the
 * reference draws it; predicted_x1_f32 may be NULL. */
int tfx_overshoot_step(const void* model_output, const void* sample, const void* noise_f32, void* prev_sample,
                       void* predicted_x1_f32, int64_t n, float t_overshoot_minus_t, float a, float b, float sigma, void* stream);

/* One whole sampling step = loop body of pipeline_flux_fill.py:2082-2098 in one call: cat(latents, cond) -> forward
 * -> Euler update fused into the last GEMM's store.  latents_in/out [B,S,out_channels] (may alias),
 * cond [B,S,in_channels-out_channels], noise_pred_out optional (NULL to skip). */
int tfx_step(tfx_handle h, const void* latents_in, const void* cond, const void* encoder_hidden_states,
             const void* pooled, const void* timestep_bf16, const void* guidance_f32, const void* img_ids,
             const void* txt_ids, float sigma, float sigma_next, void* latents_out, void* noise_pred_out, void* stream);

/* Hoists the step-invariant part of the loop: computes temb and all adaLN modulation vectors for every step of a
 * schedule at once (timesteps [n_steps, B] bf16 = t/1000 as the pipeline would pass them per step;
 * embeddings.py:1327-1339 + normalization.py:167,200,363).  Then tfx_step_scheduled(i) is tfx_step for step i without
 * the per-step pass over the 6.5 GB modulation matrix (the table is filled lazily, 8 rows = steps x samples per pass, when
 * a step first needs it).  Results are bit-identical to tfx_step. */
int tfx_set_schedule(tfx_handle h, const void* timesteps_bf16, int32_t n_steps, const void* guidance_f32, const void* pooled,
                     void* stream);
int tfx_step_scheduled(tfx_handle h, int32_t step_index, const void* latents_in, const void* cond,
                       const void* encoder_hidden_states, const void* img_ids, const void* txt_ids, float sigma,
                       float sigma_next, void* latents_out, void* noise_pred_out, void* stream);

/* ---- single kernels, exported for parity tests against the oracle --------------------------------------------- */
/* Y = epilogue(A[M,K] W[N,K]^T + bias); mode: 0 store, 1 gelu-tanh, 2 out = res + gate*(.) (gate [N], res [M,N]);
 * cta_group: 1 | 2 plain kernels, 22 | 24 multicast kernel with 2 | 4 CTA pairs per cluster */
int tfx_op_linear(const void* A, int64_t lda, const void* W, const void* bias, void* out, int64_t ldo, int32_t M,
                  int32_t N, int32_t K, int32_t mode, const void* gate, const void* res, int32_t cta_group, void* stream);
/* tfx_op_linear with the unfused-LoRA side path: Y = epilogue(A W^T + bf16(A la^T) lb^T + bias), la [64, K], lb [N, 64];
 * t_scratch [M, 64] bf16 receives bf16(A la^T).  m_band: GEMM tile order (0 = M-fastest; b > 0 = bands of b M tiles; b < 0 = N bands of -b tiles; b + 1000 / b - 1000 = the same with
 * the k direction alternating per band, the model option "gemm_k_snake"). */
int tfx_op_linear_lora(const void* A, int64_t lda, const void* W, const void* bias, const void* la, const void* lb, void* t_scratch,
                       void* out, int64_t ldo, int32_t M, int32_t N, int32_t K, int32_t mode, const void* gate, const void* res,
                       int32_t cta_group, int32_t m_band, void* stream);
/* The fused QKV projection of one stream, as the engine launches it for attention_processor.py:1987-2037: Y = A W^T + bias
 * with W [3*H*dh, K] = to_q;to_k;to_v, then per head RMSNorm(q), RMSNorm(k) (normalization.py:532-549, weights rms_q/rms_k
 * [dh]) and apply_rotary_emb (embeddings.py:879-925) with the (cos,sin) table rope_f32 [n_joint, dh/2, 2], scattered
 * head-major: row m of A is token pos_offset + m % rows_per_sample of sample m / rows_per_sample in q,k,v [*,H,n_joint,dh]. */
int tfx_op_linear_qkv(const void* A, int64_t lda, const void* W, const void* bias, const void* rms_q, const void* rms_k,
                      const void* rope_f32, void* q, void* k, void* v, int32_t M, int32_t K, int32_t H, int32_t head_dim,
                      int32_t rows_per_sample, int32_t pos_offset, int32_t n_joint, int32_t cta_group, void* stream);
/* proj_out with the scheduler step fused on its store (transformer_flux.py:1203 + scheduling_flow_match_euler_discrete.py:322-330):
 * v = bf16(A W^T + bias) -> noise_pred_out [M,N] (NULL to skip); latents_out = bf16(latents_in + bf16(dt * v)) with
 * dt_f32_dev a DEVICE scalar holding float(bf16(sigma_next - sigma)). */
int tfx_op_linear_euler(const void* A, int64_t lda, const void* W, const void* bias, const void* latents_in, const void* dt_f32_dev,
                        void* noise_pred_out, void* latents_out, int32_t M, int32_t N, int32_t K, int32_t cta_group, void* stream);
/* q,k,v [B,H,N,dh] -> out rows in the engine's [B*T text rows ; B*S image rows] order, row stride ld_out.
 * schedule = code % 10: 5 | 6 = schedule 3 (attention3.cuh: one CTA per 256 query rows of a head; 6 hands P to the tensor
 * pipe in two halves), 9 = schedule 4 (attention4.cuh: schedule 3/6 plus the last partial wave cut into KV shares),
 * 7 = schedule 5 (attention5.cuh: one persistent CTA per SM, items overlapped, remainder cut into KV shares); + 10 * emu (exponentials per 8 column pairs evaluated by the FMA-pipe polynomial: 0, 2, 3, 4);
 * + 100 with schedule 6 records clock stamps (tfx_debug_set_attention_trace) */
int tfx_op_attention(const void* q, const void* k, const void* v, void* out, int64_t ld_out, int32_t B, int32_t H,
                     int32_t T, int32_t S, int32_t head_dim, int32_t schedule_code, void* stream);
/* debug: device buffer [n_kv_tiles * 2 * 8] int64 receiving clock64 stamps of CTA (0,0,0) of the schedule-3 attention
 * kernel when tfx_op_attention is called with code >= 100 (tools/attn_trace.py); NULL switches it off */
int tfx_debug_set_attention_trace(void* dev_ptr);
/* debug: device buffer [num CTAs * 8] int64 receiving per-CTA globaltimer stamps (entry, set-up done, Q + first K landed,
 * first scores seen, last P handed over, last PV retired, output stored, SM id) of the same trace build
 * (tools/attn_timeline.py); NULL switches it off */
int tfx_debug_set_attention_cta_trace(void* dev_ptr);
/* y = LN(x)*(1+scale)+shift per row; mod [B, mod_stride]; rows = B*rows_per_sample */
int tfx_op_ln_modulate(const void* x, void* y, int32_t rows, int32_t D, int32_t rows_per_sample, const void* mod,
                       int64_t mod_stride, int64_t shift_off, int64_t scale_off, void* stream);
/* y[b,n] = post(sum_k pre(x[b,k]) W[n,k] + bias[n]); flags: 1 silu(x), 2 silu(y), 4 out += y */
int tfx_op_gemv(const void* x, int32_t B, int32_t K, const void* W, const void* bias, int64_t N, void* out,
                int32_t flags, void* stream);
/* (cos,sin) table [T+S, dh/2] float2 from bf16 ids */
int tfx_op_rope_table(const void* txt_ids, const void* img_ids, int32_t T, int32_t S, const int32_t* axes_dims,
                      void* out_f32, void* stream);
/* sinusoidal embedding [B,256] bf16 of bf16(bf16(t)*1000) */
int tfx_op_timestep_embed(const void* t, int32_t is_f32, int32_t B, void* out, void* stream);
/* ---- conditioning glue either side of the loop (pure index permutations, bit-exact) ------------ */
/* replaces FluxFillPipeline._pack_latents (pipeline_flux_fill.py:1743-1748), optionally preceded by the VAE latent
 * normalisation `(x - shift_factor) * scaling_factor` of prepare_mask_latents (:1536):
 *   dst[b, i*(w/2)+j, dst_off + c*4 + di*2 + dj] = f(src[b, c, 2i+di, 2j+dj]);  src [B,C,h,w] bf16 or fp32, dst bf16 rows of stride
 *   dst_ld elements (so latents, masked-image latents and mask can be packed straight into one [B,S,384] buffer) */
int tfx_op_pack_latents(const void* src, int32_t src_is_f32, void* dst, int64_t dst_ld, int64_t dst_off, int32_t B, int32_t C,
                        int32_t h, int32_t w, int32_t affine, float shift, float scale, void* stream);
/* replaces FluxFillPipeline._unpack_latents (:1752-1765), optionally followed by `x / scaling_factor + shift_factor` (:2127) */
int tfx_op_unpack_latents(const void* src, int64_t src_ld, void* dst, int32_t B, int32_t C, int32_t h, int32_t w, int32_t affine,
                          float shift, float scale, void* stream);
/* replaces the mask reshape + pack of prepare_mask_latents (:1563-1580): mask [B,1,h*vs,w*vs] (bf16 or fp32) ->
 *   dst[b, i*(w/2)+j, dst_off + (py*vs+px)*4 + di*2 + dj] = mask[b, 0, (2i+di)*vs+py, (2j+dj)*vs+px] */
int tfx_op_pack_mask(const void* mask, int32_t mask_is_f32, void* dst, int64_t dst_ld, int64_t dst_off, int32_t B, int32_t h,
                     int32_t w, int32_t vae_scale_factor, void* stream);
/* raw tcgen05 descriptor probe (bring-up / regression of the UMMA encodings); see tests/test_gpu_ops.py */
int tfx_op_umma_probe(const void* A, const void* Bm, void* D_f32, int32_t n_dim, int32_t k_dim, int32_t b_mn_major,
                      int32_t a_from_tmem, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep_bytes, void* stream);

/* ---- AutoencoderKL: vae.encode at pipeline_flux_fill.py:1528, vae.decode at :2128 ------------------- */
typedef struct tfx_vae* tfx_vae_handle;

/* Mirrors AutoencoderKL.__init__'s @register_to_config arguments (models/autoencoders/autoencoder_kl.py:76-97) for the
 * DownEncoderBlock2D / UpDecoderBlock2D family with act_fn = "silu" and no quant / post-quant convolutions (FLUX's VAE). */
typedef struct {
  int32_t in_channels;      /* 3 */
  int32_t out_channels;     /* 3 */
  int32_t latent_channels;  /* 16 */
  int32_t num_blocks;       /* len(block_out_channels), 4 */
  int32_t block_out_channels[8]; /* (128, 256, 512, 512) */
  int32_t layers_per_block; /* 2 */
  int32_t norm_num_groups;  /* 32 */
  int32_t mid_block_add_attention;
} tfx_vae_config;

/* replaces AutoencoderKL.__init__ */
int tfx_vae_create(const tfx_vae_config* cfg, int32_t device, tfx_vae_handle* out);
void tfx_vae_destroy(tfx_vae_handle h);
const char* tfx_vae_last_error(tfx_vae_handle h);
/* "launches": kernels launched by this handle */
int tfx_vae_get_counter(tfx_vae_handle h, const char* key, int64_t* value);
/* Registers a bf16 device tensor under its reference state-dict name (replaces load_state_dict).  Layouts: 3x3 convolutions
 * `<name>.weight` [Cout, 9 * Cin64] with column (ky * 3 + kx) * Cin64 + c, Cin64 = Cin rounded up to 64 (zero padded); 1x1
 * convolutions and Linear layers [Cout, Cin]; biases and GroupNorm weight / bias [1, C].  INTEGRATION.md lists the names. */
int tfx_vae_set_weight(tfx_vae_handle h, const char* name, const void* dev_ptr, int64_t rows, int64_t cols);
/* replaces AutoencoderKL.encode up to the distribution parameters (autoencoder_kl.py:240-272 -> Encoder.forward, vae.py:140-193):
 * image [B, in_channels, H, W] (fp32 or bf16, NCHW) -> moments [B, 2 * latent_channels, H / f, W / f] bf16 (mean ; logvar),
 * f = 2^(num_blocks - 1).  DiagonalGaussianDistribution.sample / .mode stay with the caller (they draw from torch's generator). */
int tfx_vae_encode(tfx_vae_handle h, const void* image, int32_t image_is_f32, int32_t B, int32_t H, int32_t W, void* moments_out,
                   void* stream);
/* replaces AutoencoderKL.decode (autoencoder_kl.py:274-324 -> Decoder.forward, vae.py:284-343): latents [B, latent_channels, h, w]
 * bf16 -> image [B, out_channels, h * f, w * f] bf16 */
int tfx_vae_decode(tfx_vae_handle h, const void* latents, int32_t B, int32_t h_latent, int32_t w_latent, void* image_out, void* stream);
/* replaces DiagonalGaussianDistribution.sample (models/autoencoders/vae.py:793-803) once the caller has drawn `noise` from torch's
 * generator where the reference draws it: out = bf16(mean + bf16(bf16(exp(bf16(0.5 * clamp(logvar, -30, 20)))) * noise)),
 * moments [B, 2 * latent_channels, hw] bf16 (mean ; logvar), noise / out [B, latent_channels, hw] bf16 */
int tfx_op_gaussian_sample(const void* moments, const void* noise, void* out, int32_t B, int32_t latent_channels, int64_t hw, void* stream);

/* ---- prompt encoders: text_encoder_2 = T5EncoderModel (T5-XXL v1.1), text_encoder = CLIPTextModel
 * (CLIP-L), called at pipeline_flux_fill.py:1411-1503 (_get_t5_prompt_embeds :1438, _get_clip_prompt_embeds :1483).  Both live in the
 * third-party `transformers` package (reference pin 4.43.3, requirements.txt:4): models/t5/modeling_t5.py, models/clip/modeling_clip.py.
 * Tokenisation stays with the caller (transformers' tokenizers on the host); these entry points start at input_ids. ------------------ */
typedef struct tfx_textenc* tfx_textenc_handle;
enum { TFX_TEXTENC_T5 = 0, TFX_TEXTENC_CLIP = 1 };

/* T5Config / CLIPTextConfig fields the encoders read */
typedef struct {
  int32_t kind;              /* TFX_TEXTENC_T5: encoder-only T5 v1.1 (gated-gelu, RMS norm, relative position bias, no biases);
                                TFX_TEXTENC_CLIP: CLIP text transformer (pre-LN, causal mask, quick_gelu) */
  int32_t vocab_size;        /* 32128 | 49408 */
  int32_t d_model;           /* 4096 | 768 (hidden_size) */
  int32_t d_kv;              /* 64 (head dimension; CLIP: hidden_size / num_attention_heads) */
  int32_t num_heads;         /* 64 | 12 */
  int32_t num_layers;        /* 24 | 12 */
  int32_t d_ff;              /* 10240 | 3072 (intermediate_size) */
  int32_t max_positions;     /* CLIP: max_position_embeddings (77) */
  int32_t rel_buckets;       /* T5: relative_attention_num_buckets (32) */
  int32_t rel_max_distance;  /* T5: relative_attention_max_distance (128) */
  float eps;                 /* layer_norm_epsilon 1e-6 | layer_norm_eps 1e-5 */
} tfx_textenc_config;

int tfx_textenc_create(const tfx_textenc_config* cfg, int32_t device, tfx_textenc_handle* out);
void tfx_textenc_destroy(tfx_textenc_handle h);
const char* tfx_textenc_last_error(tfx_textenc_handle h);
int tfx_textenc_get_counter(tfx_textenc_handle h, const char* key, int64_t* value);
/* bf16 device tensors under packed names: "embed" [vocab, D]; per layer i "l<i>.ln1.w", "l<i>.qkv.w" [3 H 64, D]
 * (q ; k ; v rows), "l<i>.o.w" [D, H 64], "l<i>.ln2.w"; T5: "rel_bias" [rel_buckets, H], "l<i>.wi0.w", "l<i>.wi1.w" [d_ff, D],
 * "l<i>.wo.w" [D, d_ff], "final_ln.w"; CLIP: "pos" [max_positions, D], ".b" next to every ".w", "l<i>.fc1.w" [d_ff, D], "l<i>.fc2.w" */
int tfx_textenc_set_weight(tfx_textenc_handle h, const char* name, const void* dev_ptr, int64_t rows, int64_t cols);
/* replaces T5EncoderModel.forward(input_ids)[0] / CLIPTextModel.forward(input_ids) -> (last_hidden_state, pooler_output):
 *   input_ids [B, T] int32 (device); last_hidden_state [B, T, D] bf16
 *   rel_bucket_lut (T5) [2T - 1] int32: bucket of relative position d = key - query at index d + T - 1 (T5Attention._relative_position_bucket,
 *     evaluated by the caller exactly as the reference evaluates it)
 *   pooled_index (CLIP) [B] int32 = position of each sample's EOS token (argmax rule of CLIPTextTransformer.forward), pooled_out [B, D]
 *     bf16 = last_hidden_state[b, pooled_index[b]]; both NULL to skip */
int tfx_textenc_encode(tfx_textenc_handle h, const int32_t* input_ids, int32_t B, int32_t T, const int32_t* rel_bucket_lut,
                       void* last_hidden_state, const int32_t* pooled_index, void* pooled_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TEXTFLUX_B200_H_ */
