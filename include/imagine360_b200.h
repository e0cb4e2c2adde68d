/* imagine360_b200 -- C ABI of the B200-native (sm_100a) kernels behind Imagine360's dual-branch
 * denoising hot path.
 *
 * The reference (3DTopia/Imagine360) is pure Python: it has no FFI of its own, its "operator
 * interface" for this path is the set of torch / xformers / kornia library calls inventoried in
 * SURVEY.md section 2.1.  Each entry point below replaces one of those call sites (cited as
 * reference file:line) and is what a binding from the reference's Python modules would call
 * (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator); the library
 *     borrows it for the duration of the call, never frees it and allocates no device memory itself;
 *   - activations are bf16, channels-last: images [N, H, W, C], tokens [rows, C]; weights bf16 [out, in];
 *   - `stream` is a cudaStream_t; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 on success, negative I360_ERR_* otherwise (never throws, never aborts).
 */
#ifndef IMAGINE360_B200_H
#define IMAGINE360_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define I360_OK 0
#define I360_ERR_ARG (-1)
#define I360_ERR_CUDA (-2)
#define I360_ERR_TMAP (-3)
#define I360_ERR_UNSUPPORTED (-4)

/* epilogue activations of i360_gemm_bf16 */
#define I360_ACT_NONE 0
#define I360_ACT_GEGLU 1 /* tile = [values | gates] -> values * gelu_erf(gates)  (diffusers/models/activations.py:93-122) */
#define I360_ACT_GELU 2  /* nn.GELU (animatediff/models/resampler.py:19) */
#define I360_ACT_SILU 3  /* TimestepEmbedding.act (diffusers/models/embeddings.py:210) */

/* D[M,N] = act(A[M,K] W[N,K]^T + bias[N] + rowvec[(row/rowvec_div), N]) + resid[M,N], then * out_scale.
 * Replaces nn.Linear / LoRACompatibleLinear (diffusers/models/lora.py:368) at every call site of the path:
 * attention projections (attention_processor.py:1244-1262), FeedForward (attention_lora.py:540-547),
 * proj_in/proj_out (animatediff/models/attention.py:271,:289), time embeddings, the adapter.
 * tcgen05 tensor cores, TMA-fed, persistent, fp32 accumulation in TMEM.  lda/ldw/ldd/ldr in elements. */
int i360_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* D, long long ldd, int M, int N,
                   int K, const void* bias, const void* resid, long long ldr, const float* rowvec, int rowvec_div,
                   int rowvec_ld, int act, float out_scale, void* stream);

/* The same GEMM (bias, residual; no activation) that ALSO writes row statistics of what it stores:
 * rowstats[(slot * M + r) * 2 + {0, 1}] = (sum, sum of squares) over the columns of row r handled by one (N tile, epilogue
 * group) slot, for slot < i360_gemm_rowstats_slots(M, N, K, resid != NULL).  No atomics: the consumer adds the slots in a
 * fixed order, so the result is deterministic.  Producer half of the LayerNorm fold below. */
int i360_gemm_rowstats_slots(int M, int N, int K, int has_resid);
int i360_gemm_rowstats_bf16(const void* A, long long lda, const void* W, long long ldw, void* D, long long ldd, int M, int N,
                            int K, const void* bias, const void* resid, long long ldr, float* rowstats, void* stream);

/* LayerNorm folded into the projection that consumes it: D = act(LN(A; gamma, beta, eps) W^T + bias (+ rowvec)).
 * Wf = W * gamma (bf16, [N, K]); u[n] = sum_k Wf[n,k]; c[n] = sum_k beta[k] W[n,k] + bias[n] (fp32).  The row statistics
 * of A come from the GEMM that produced it (rowstats / slots as written by i360_gemm_rowstats_bf16); the epilogue applies
 * rstd * acc - mean * rstd * u + c, so the normalised activations never travel through HBM and no LayerNorm pass runs.
 * K = LayerNorm width (multiple of 64).  rowvec fp32 [rowvec_mod, N] is added at row (r / rowvec_div) % rowvec_mod (temporal
 * PE term (LN(x) + pe_f) W^T, motion_module.py:350).  act: 0 or 1 (GEGLU, Wf / u / c packed like the weight).
 * Replaces nn.LayerNorm + nn.Linear (animatediff/models/attention.py:463-508; motion_module.py:247-259). */
int i360_gemm_ln_bf16(const void* A, long long lda, const void* Wf, long long ldw, void* D, long long ldd, int M, int N,
                      int K, const float* u, const float* c, float eps, const float* rowstats, int slots,
                      const float* rowvec, int rowvec_div, int rowvec_mod, int rowvec_ld, int act, void* stream);
/* Tile width i360_gemm_ln_bf16 uses for (N, K, act); 0 = unsupported (keep i360_layernorm + i360_gemm_bf16). */
int i360_gemm_ln_supported(int N, int K, int act);

/* Row-block size the GEGLU weight must be packed with ([block/2 value rows | block/2 gate rows] repeating). */
int i360_gemm_geglu_block(int n_total);

/* 3x3 / stride 1 / zero-pad 1 convolution as an im2col-free implicit GEMM over NHWC x[B,H,W,Cin], with up to
 * two fused 1x1 sources x2/x3 (ResnetBlock3D.conv_shortcut over the [hidden | skip] concat), per-image
 * rowvec (time embedding), residual, and `crop` halo columns dropped from the output (pano circular pad).
 * Replaces nn.Conv2d inside InflatedConv3d (animatediff/models/resnet.py:19-27, call sites :227,:246,:249)
 * and the VAE's convs (diffusers/models/resnet.py:454-496).  Wt: [Cout, 9*Cin + C2 + C3], taps (kh,kw,cin). */
int i360_conv3x3_bf16(const void* x, int B, int H, int W, int Cin, const void* x2, int C2, const void* x3, int C3,
                      const void* Wt, int Cout, void* D, int crop, const void* bias, const void* resid,
                      const float* rowvec, int rowvec_div, int rowvec_ld, float out_scale, void* stream);

/* Convolutions that also accumulate the GroupNorm statistics of their OUTPUT in the epilogue (32 groups of 4 / 8 / 16 channels,
 * one image per tile): stats[(image * 32 + group) * 2 + {0, 1}] = (sum, sum of squares), fp64, zeroed by the call.  The GroupNorm
 * reading the output skips its statistics pass (pass the buffer to i360_groupnorm_apply).  Replaces the first read of
 * nn.GroupNorm after a conv in the VAE's ResnetBlock2D / Upsample2D chain (diffusers/models/resnet.py:454-496,:108-143). */
int i360_conv3x3_gnstats_bf16(const void* x, int B, int H, int W, int Cin, const void* x2, int C2, const void* Wt, int Cout,
                              void* D, const void* bias, const void* resid, int groups, double* stats, void* stream);
int i360_conv_upsample2x_gnstats_bf16(const void* x, int B, int H, int W, int Cin, const void* Weff, int Cout, void* D,
                                      const void* bias, int groups, double* stats, void* stream);

/* conv -> GroupNorm for ANY group size: i360_conv3x3_bf16 (same arguments) whose epilogue also accumulates per-CHANNEL
 * statistics of the stored output, chan_stats[(image * Cout + channel) * 2 + {0, 1}] = (sum, sum of squares), fp64, zeroed by
 * the call; and the GroupNorm (+SiLU) that folds those channels into groups instead of running a statistics pass.  Replaces
 * the first read of InflatedGroupNorm after a conv: norm2 after conv1 in ResnetBlock3D (animatediff/models/resnet.py:243) and
 * Transformer3DModel.norm after the block's conv2 (animatediff/models/attention.py:262). */
int i360_conv3x3_chanstats_bf16(const void* x, int B, int H, int W, int Cin, const void* x2, int C2, const void* x3, int C3,
                                const void* Wt, int Cout, void* D, int crop, const void* bias, const void* resid,
                                const float* rowvec, int rowvec_div, int rowvec_ld, float out_scale, double* chan_stats,
                                void* stream);
int i360_groupnorm_apply_chanstats(const void* x, int C, int B, int H, int W, int groups, const double* chan_stats,
                                   const void* gamma, const void* beta, float eps, int do_silu, void* out, void* stream);

/* 3x3 / stride 2 convolution as an implicit GEMM over TMA boxes with traversal stride 2 (no im2col buffer).  pad_lo = 1:
 * symmetric zero pad 1 (Downsample3D, animatediff/models/resnet.py:117-140; with pad_pano(2) / unpad_pano(1) of
 * MVGenModel.py:305-314 as a materialised 2-column circular halo and crop = 1); pad_lo = 0: the VAE encoder's asymmetric
 * F.pad(0, 1, 0, 1) (diffusers/models/resnet.py:184).  x: NHWC [B, Hin, Win, Cin], D: NHWC [B, Hin/2, Win/2 - 2 crop, Cout],
 * Wt: [Cout, 9*Cin] taps (kh, kw, cin). */
int i360_conv3x3_s2_bf16(const void* x, int B, int Hin, int Win, int Cin, const void* Wt, int Cout, void* D, int pad_lo,
                         int crop, const void* bias, void* stream);

/* Nearest 2x upsample + 3x3 / pad 1 convolution in sub-pixel form: no upsampled tensor, 4/9 of the multiply-adds.  Each output
 * parity (a, b) = (row % 2, col % 2) is a 2x2-tap convolution of the LOW-resolution input with pre-summed weights
 * Weff[a*2+b][Cout][(dr*2+dc)*Cin + c] (a = 0: rows {w[0], w[1]+w[2]}, a = 1: {w[0]+w[1], w[2]}, same for columns; summed in
 * fp32 and rounded once to bf16).  x: NHWC [B, H, W, Cin] (W includes `crop` circular halo columns per side), D: NHWC
 * [B, 2H, 2(W - 2 crop), Cout].  Replaces Upsample3D.forward (F.interpolate nearest x2 + InflatedConv3d,
 * animatediff/models/resnet.py:86-114; pad_pano(1) / unpad_pano(2) of MVGenModel.py:449-456 as crop = 1) and the VAE's
 * Upsample2D (diffusers/models/resnet.py:108-143). */
int i360_conv_upsample2x_bf16(const void* x, int B, int H, int W, int Cin, const void* Weff, int Cout, void* D, int crop,
                              const void* bias, void* stream);

/* 1 when i360_conv3x3_bf16 runs this problem through its halo-tile variant (16 x 8 pixel tiles whose 18 x 10 halo is
 * loaded once per 64-channel block and shared by the nine taps), 0 for the tap-by-tap variant.  Informational. */
int i360_conv3x3_uses_halo(int B, int H, int W, int Cin, int has_resid, int has_rowvec, int has_extra_sources);
/* Selection rule of the halo variant: enabled / padded-work tolerance against the best free-form pixel box (default
 * 1.04) / allow fused 1x1 sources (default 0) / smallest image side (default 64); a negative argument keeps the current
 * value.  Tests and A/B runs only. */
void i360_conv3x3_halo_policy(int on, double tol, int allow_extra, int min_hw);

/* GroupNorm statistics / application over a virtual tensor = concat(x1, x2) along C, circularly padded by `pad`
 * columns.  Replaces InflatedGroupNorm / nn.GroupNorm (+ SiLU) (animatediff/models/resnet.py:9-17,:224-225,:243)
 * together with torch.cat (MVGenModel.py:399) and pad_pano (src/utils/pano.py:75-92).  stats: [B, groups, 2] fp64. */
int i360_groupnorm_stats(const void* x1, int C1, const void* x2, int C2, int B, int H, int Wsrc, int pad, int groups,
                         double* stats, void* stream);
int i360_groupnorm_apply(const void* x1, int C1, const void* x2, int C2, int B, int H, int Wsrc, int pad, int groups,
                         const double* stats, double stats_count, const void* gamma, const void* beta, float eps,
                         int do_silu, void* out, void* stream);

/* LayerNorm over the last dim; optional bf16 table added before (spherical PE of WarpAttn,
 * src/modules/transformer.py:155-161) and fp32 table added after (temporal PE, motion_module.py:350).
 * Replaces nn.LayerNorm (animatediff/models/attention.py:463,:474,:505; motion_module.py:248,:256). */
int i360_layernorm(const void* x, long long ldx, void* y, long long ldy, long long M, int C, const void* gamma,
                   const void* beta, float eps, const void* pre_add, int pre_div_a, int pre_mod_a, int pre_mul_a,
                   int pre_mod_b, const float* post_add, int post_div, int post_mod, void* stream);

/* A strided 4-D token view [channels | d1, d2, d3] (strides in elements, channel stride 1).
 * Batch item bi selects c2 = (bi % A) / Bdiv and c3 in [(bi / A) * mul, +ext3); a sequence is the d1 tokens of
 * those ext3 dim-3 entries. */
typedef struct I360TokenView {
  const void* ptr;
  int channels;
  int col0;
  int d1, d2, d3;
  long long s1, s2, s3;
  int A, Bdiv, mul;
  int ext3;
} I360TokenView;

/* O = softmax(Q K^T * scale + bias) V per (batch item, head); head_dim 32 or 64; optional dense bf16 bias
 * [bias_rows, bias_cols] shared by all batch items and heads; accumulate != 0: O_out = bf16(O_out + bf16(O)).
 * Replaces xformers.ops.memory_efficient_attention (diffusers/models/attention_processor.py:1264,:641;
 * src/modules/transformer.py:72) and F.scaled_dot_product_attention (attention_processor.py:1351). */
int i360_attention_bf16(const I360TokenView* q, const I360TokenView* k, const I360TokenView* v,
                        const I360TokenView* o, int heads, int head_dim, int batch, float scale, const void* bias,
                        int bias_rows, int bias_cols, int accumulate, void* stream);

/* The same attention with one bias block per (batch item, head): bias is [batch * heads * bias_item_rows, bias_cols]
 * bf16, block (bi * heads + h) = the [Nq, Nk] additive logits of that head.  Plain sequences, head_dim 64.
 * Replaces `attn = (q * scale) @ k^T; attn = add_decomposed_rel_pos(attn, ...); softmax; attn @ v` of
 * segment_anything's ImageEncoderViT (modeling/image_encoder.py Attention.forward, release 1.0; the package is a
 * dependency of the reference: requirements.txt:11, inference_dual_p2e.py:369-370; its encoder is run by
 * animatediff/pipelines/pipeline_animation_inference_dual.py:685-690,:708-713). */
int i360_attention_item_bias_bf16(const I360TokenView* q, const I360TokenView* k, const I360TokenView* v,
                                  const I360TokenView* o, int heads, int head_dim, int batch, float scale,
                                  const void* bias, int bias_item_rows, int bias_cols, void* stream);

/* SAM ViT decomposed relative position bias: bias[(item, head), q, k] = q . Rh[qh - kh + S - 1] + q . Rw[qw - kw + S - 1]
 * for sequences of S x S tokens (a 14 x 14 window or the 64 x 64 grid), q the unscaled query of the head at columns
 * col0 + head * hd of the [items * S*S, ldq] projection output; rel_h / rel_w: [2S - 1, hd] bf16; bias rows of ldb
 * elements (ldb % 8 == 0, columns past S*S zeroed).  Replaces get_rel_pos + add_decomposed_rel_pos (two einsums and a
 * broadcast add on the [B*heads, S^2, S^2] logits) of segment_anything modeling/image_encoder.py. */
int i360_relpos_bias_bf16(const void* q, long long ldq, int col0, int items, int heads, int hd, int S,
                          const void* rel_h, const void* rel_w, void* bias, int ldb, void* stream);

/* Fused text + image-prompt cross-attention: O = softmax(Q Kt^T * scale) Vt + softmax(Q Ki^T * scale) Vi, summed in
 * fp32 and rounded once.  q, o: [rows, heads*64] (row strides in elements); rows = n_ctx * rows_per_ctx, the rows of
 * clip element e (all of its frames) are [e * rows_per_ctx, +rows_per_ctx) and attend to text tokens
 * kv_text[e*nt .. e*nt+nt) and image tokens kv_ip[e*ni .. e*ni+ni); kv_* = [K | V] projections, 2*heads*64 columns.
 * Keys/values stay in shared memory while the element's queries stream through (HBM bound: one read of q, one write
 * of o).  Limits: head_dim 64, nt <= 96, ni <= 96, ceil(nt/64) + ceil(ni/64) <= 3.  Replaces both attention calls and the sum of
 * IPCrossAttention (animatediff/models/attention.py:119-148) on the xformers path
 * (diffusers/models/attention_processor.py:1264). */
int i360_cross_attention_text_ip_bf16(const void* q, long long ldq, void* o, long long ldo, long long rows, int n_ctx,
                                      const void* kv_text, long long ld_kvt, int nt, const void* kv_ip, long long ld_kvi,
                                      int ni, int heads, int head_dim, float scale, void* stream);

/* Attention over the frame axis (F <= 32) for every (clip b, pixel d, head), rows ordered (b, f, d).
 * Replaces the baddbmm/softmax/bmm math path of VersatileAttention (animatediff/models/motion_module.py:343-429,
 * diffusers/models/attention_processor.py:562-591) and of TemporalProjection.attn_temp (resampler.py:246,:259). */
int i360_temporal_attention_bf16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                 long long ldv, void* o, long long ldo, int B, int F, int D, int heads, int head_dim,
                                 float scale, void* stream);

/* Row softmax with fp32 math over materialised bf16 logits [M, N] (the VAE's single-head AttentionBlock,
 * diffusers/models/attention.py:353-367: head_dim 512 does not fit the fused kernel's TMEM budget). */
int i360_softmax_rows_bf16(const void* x, void* y, long long ld, long long M, int N, void* stream);

/* nearest x2 upsample, out [B, 2H, 2(W+2*pad_in), C] (Upsample3D, animatediff/models/resnet.py:86-110, with the
 * pano halo of MVGenModel.py:449-452 folded in). */
int i360_upsample2x_nhwc(const void* x, void* out, int B, int H, int W, int C, int pad_in, void* stream);

/* im2col for the stride-2 downsample convs (Downsample3D, resnet.py:117-140; circular = pano halo of
 * MVGenModel.py:305-314; pad_lo = 0 = the VAE encoder's asymmetric (0,1,0,1) pad, diffusers/models/resnet.py:181-190).
 * out [B*(H/2)*(W/2), 9*C]. */
int i360_im2col3x3_s2_nhwc(const void* x, void* out, int B, int H, int W, int C, int circular, int pad_lo,
                           void* stream);

/* out = a*x + bf16(b*y)  (add_noise_to_condition, src/models/MVGenModel.py:11-14). */
int i360_axpby_bf16(const void* x, const void* y, void* out, float a, float b, long long n, void* stream);

/* Fused classifier-free guidance + DDIM v-prediction step with the reference's bf16 rounding sequence
 * (pipeline_animation_inference_dual.py:789-800; diffusers/schedulers/scheduling_ddim.py:319-346). */
int i360_cfg_ddim_step_bf16(const void* latent, const void* pred_uncond, const void* pred_cond, void* out,
                            float guidance, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
                            float sqrt_1m_a_prev, long long n, void* stream);

/* mean over groups of 4 frames (F.avg_pool1d(kernel 4), animatediff/models/resampler.py:251,:264). */
int i360_avgpool_frames4_bf16(const void* x, void* out, int B, int F, long long DC, void* stream);

/* grid_sample(align_corners=True, zeros padding), bilinear or nearest, NCHW fp32 (kornia.geometry.transform.remap
 * as called by e2p.py:77 / p2e.py:70; init_noise pipeline_animation_inference_dual.py:376-379). */
int i360_grid_sample_f32(const float* img, const float* grid, float* out, int N, int C, int Hi, int Wi, int Ho, int Wo,
                         int nearest, void* stream);

/* ---- geometric pre-processing (the CPU step right before the denoising path) ---- */

/* OpenCV's fixed-point bicubic weight table for 8-bit remaps (imgproc/src/imgwarp.cpp: interpolateCubic,
 * initInterTab2D(INTER_CUBIC, fixpt)): out[(fy*32 + fx)*16 + ky*4 + kx], int16, scale 2^15.  Host only, no GPU. */
int i360_remap_cubic_table_i16(short* out_1024x16);

/* cv2.remap(src, mapx, mapy, INTER_CUBIC, borderMode=BORDER_WRAP) for uint8 RGB frames, bit-exact, batched:
 * src [n_img, H, W, 3]; maps [n_map, h, w] float32 (device).  paired = 0: every frame is sampled with every map
 * (outputs [n_img, n_map, ...]: process_equi, inference_dual_p2e.py:113-144, Equirec2Perspec.py:61); paired = 1:
 * frame i with map i (outputs [n_img, 1, ...]: get_anchor_target, video_mask.py:158-173; pers2pano_vid,
 * inference_dual_p2e.py:293-301, Perspec2Equirec.py:74).  keep: optional uint8 [n_map, h, w], pixels with keep == 0
 * are zeroed (`persp * mask`, Perspec2Equirec.py:78).  table: device copy of i360_remap_cubic_table_i16.
 * out_u8 [.., h, w, 3] and/or out_f32: f32_mode 1 = [.., 3, h, w] (u8 / 127.5) - 1; 2 = [.., 1, h, w] any(u8 > 0). */
int i360_remap_cubic_wrap_u8(const void* src, int n_img, int H, int W, const float* mapx, const float* mapy,
                             const void* keep, int n_map, int h, int w, int paired, const short* table, void* out_u8,
                             float* out_f32, int f32_mode, void* stream);

/* float32 frames -> uint8 [n, H, W, 3] with float32 arithmetic and C truncation like torch + numpy .astype(uint8):
 * mode 0: x * 255; 1: (x + 1) * 127.5 (inference_dual_p2e.py:122-129; video_mask.py:168-169);
 * 2: ((x + 1) / 2) * 255 (save_videos_grid with rescale, animatediff/utils/util.py:55-72; mode 0 = without).
 * Channel c of frame i starts at x + i * frame_stride + c * chan_stride (elements), rows are contiguous: covers both
 * [n, 3, H, W] frames and the [c, t, h, w] video layout the pipeline returns. */
int i360_frames_to_u8_nhwc(const float* x, long long frame_stride, long long chan_stride, void* out, long long n, int H,
                           int W, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IMAGINE360_B200_H */
