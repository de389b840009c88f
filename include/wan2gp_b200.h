/* wan2gp_b200 -- C ABI of the B200 (sm_100a) kernels behind the Wan DiT denoise / WanVAE decode hot path.
 *
 * The reference (deepbeepmeep/Wan2GP) has no FFI on this path: its "operator API" is Python duck typing
 * (SURVEY.md section 8b).  This header is the boundary BELOW the Python modules that mirror
 * models/wan/modules/model.py::WanModel and models/wan/modules/vae.py::WanVAE; each entry point names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding the reference side would add.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (PyTorch allocates); kernels never
 * allocate, never synchronise, and run on `stream` (a cudaStream_t passed as void*).  bf16 = raw 16-bit
 * brain-float.  Return value 0 = ok, < 0 = error (text from b200_last_error(), thread-local).  No
 * exceptions cross the ABI.  One host thread drives one device (the reference's single "generation"
 * worker thread, shared/utils/thread_utils.py:9-52).
 */
#ifndef WAN2GP_B200_H
#define WAN2GP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_DRIVER (-3)

#define B200_ACT_NONE 0
#define B200_ACT_GELU_TANH 1
#define B200_ACT_SILU 2
#define B200_ACT_GELU_ERF 3

const char* b200_last_error(void);
int b200_version(void);

/* size (floats) of the CFG-Zero* scratch buffer: {<c,u>, <u,u>, per-CTA partial sums, completion ticket}.  The reduction is
 * evaluated in a fixed order, so identical inputs give bit-identical alpha on every launch and every rank. */
#define B200_CFG_DOTS_FLOATS (2 + 2 * 1184 + 2)
/* number of kernels this library has launched in the calling process (bench.py reports it as gpu_launches) */
long long b200_launch_count(void);

/* D[M,N] = act(A[M,K] * B^T + bias) [* gate] ; A bf16 row-major (lda), B bf16 [N,K] row-major (ldb)
 * (nn.Linear weight layout) or, with b_mn_major=1, B bf16 [K,N] row-major.
 * out: bf16 (out_fp32=0) or fp32 (out_fp32=1); accumulate=1 (fp32 only): out += value (residual stream).
 * bias/gate: fp32 [N] or NULL; residual_bf16: bf16 [M,N] with row stride ldc, added after the gate, or NULL.
 * N % 8 == 0, K % 8 == 0, lda/ldb/ldc % 8 == 0 (partial tiles are handled by TMA zero fill + masked stores).
 * Replaces nn.Linear on the hot path: model.py:322/337 (q,k,v), :405 (o), :551-553/:694-710 (ffn),
 * :1133-1135 (text_embedding), :659/:710 (gated residual addcmul_), :668 (cross-attn residual). */
int b200_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                   long long ldc, const float* bias, const float* gate, const void* residual_bf16, int act,
                   int out_fp32, int accumulate, int b_mn_major, void* stream);

/* y[L,D] bf16 = LN(x[L,D] fp32) * (1 + scale[D]) + shift[D]   (affine=0: WanLayerNorm + modulation, model.py:634-638,
 * 686-691, 857-861)  or  LN(x) * scale + shift (affine=1: norm3, model.py:664).  D % 4 == 0, D <= 8192.
 * pre_round=1 rounds LN(x) to bf16 before the modulation (Hunyuan: hyvideo/modules/models.py:210-216, 289-291). */
int b200_ln_modulate(const float* x, const float* shift, const float* scale, int affine, int pre_round, void* y_bf16,
                     int L, int D, float eps, void* stream);

/* in place on bf16 rows x[L, D] (row stride ld): RoPE(x * rsqrt(mean(x^2)+eps) * w); cos/sin fp32 [L,128] or NULL.
 * WanRMSNorm over the full dim (model.py:152-175) + apply_rotary_emb (posemb_layers.py:251-269). D % 128 == 0.
 * per_head=1: statistics per 128-wide head with w[128] (Hunyuan RMSNorm.apply_, hyvideo/modules/norm_layers.py:62-70,
 * used at hyvideo/modules/models.py:226-228, 254-255). */
int b200_rmsnorm_rope(void* x_bf16, long long ld, const float* w, int L, int D, float eps, const float* cos_t,
                      const float* sin_t, int per_head, void* stream);

/* the same for the q and k column blocks of one fused q|k|v buffer in ONE launch (two norm weights, shared RoPE tables):
 * replaces norm_q / norm_k + the two apply_rotary_emb calls of WanSelfAttention.forward (model.py:343-344, 372-377) and of
 * MMDoubleStreamBlock (hyvideo/modules/models.py:226-228, 254-255). */
int b200_qk_rmsnorm_rope(void* q_bf16, void* k_bf16, long long ld, const float* wq, const float* wk, int L, int D, float eps,
                         const float* cos_t, const float* sin_t, int per_head, void* stream);

/* out[Lq, H*128] = softmax(q k^T / sqrt(128)) v per head; q/k/v/out bf16 with row strides ldq/ldk/ldv/ldo
 * (elements), head h at columns [128h, 128h+128).  Non-causal, no mask.
 * Replaces pay_attention -> sdpa_wrapper (shared/attention.py:208-225) at model.py:385 (self) and :265 (cross). */
int b200_attention_d128(const void* q, const void* k, const void* v, void* out, int Lq, int Lk, int H, long long ldq,
                        long long ldk, long long ldv, long long ldo, float scale, void* stream);

/* The same for nseq equally long sequences stacked along the rows (q / out [nseq*Lq, H*128], k / v [nseq*Lk, H*128]); sequence z
 * attends to its own keys only.  One launch for the cond / uncond forwards of a CFG pair, which the reference runs one after the other
 * (any2video.py:1625-1646; model.py:2030-2037 joint pass). */
int b200_attention_d128_batched(const void* q, const void* k, const void* v, void* out, int nseq, int Lq, int Lk, int H, long long ldq,
                                long long ldk, long long ldv, long long ldo, float scale, void* stream);

/* x fp32 -> bf16, n % 4 == 0 */
int b200_cast_f32_bf16(const float* x, void* y_bf16, long long n, void* stream);

/* Patch embedding Conv3d k=s=(1,2,2) (model.py:1131-1132, 1631, 1731) over cat(x0[C0], x1[C1]) [C,T,H,W] fp32
 * (x1 = i2v `y`, model.py:1597-1600; NULL/0 for t2v) -> out [L=T*H/p*W/p, D] fp32.  w fp32 [D, (C0+C1)*p*p].
 * patch p = 2 (Wan, Hunyuan 1.0) or 1 (Hunyuan 1.5 PatchEmbed, hyvideo/modules/embed_layers.py:9-60). */
int b200_patch_embed(const float* x0, int C0, const float* x1, int C1, const float* w, const float* bias, float* out,
                     int T, int H, int W, int D, int patch, void* stream);

/* unpatchify: y [L, p*p*C] fp32 -> out [C,T,H,W] fp32.  c_major=0: feature order (ph,pw,c) (Wan, model.py:2100-2126);
 * c_major=1: (c,ph,pw) (Hunyuan, hyvideo/modules/models.py:1235-1248). */
int b200_unpatchify(const float* y, float* out, int C, int T, int H, int W, int patch, int c_major, void* stream);

/* out[N] = act_out(sum_k act_in(x[k]) W[N,K] + b) fp32 GEMV (time_embedding / time_projection, model.py:1815-1818) */
int b200_gemv_f32(const float* x, const float* w, const float* b, float* out, int N, int K, int silu_in, int silu_out,
                  void* stream);
/* sinusoidal_embedding_1d (model.py:32-42) */
int b200_sinusoid(float t, float* out, int dim, void* stream);
/* same, the timestep read from device memory: a captured whole-step CUDA graph is replayed with a new t */
int b200_sinusoid_dev(const float* t_dev, float* out, int dim, void* stream);
/* out[c] = mean_r x[r, c], fp32 [rows, cols] (masked text mean, hyvideo/modules/token_refiner.py:221-226) */
int b200_col_mean_f32(const float* x, float* out, int rows, int cols, void* stream);
/* out[i] = a[i] + b[i % bmod] */
int b200_add_vec(const float* a, const float* b, float* out, int n, int bmod, void* stream);

/* lat -= dt * (u + g (c - u)); uncond may be NULL (no CFG); pred_out optional.  any2video.py:1701-1722 +
 * euler_scheduler.py:67-86.  n % 4 == 0.  star_dots (device float[B200_CFG_DOTS_FLOATS] scratch, or NULL): CFG-Zero* -- u is first rescaled by
 * alpha = <c,u> / (||u||^2 + 1e-8) computed over the whole sample (any2video.py:1706-1714, steps > cfg_zero_step). */
int b200_cfg_euler_step(float* lat, const float* cond, const float* uncond, float guide, float dt, float* pred_out,
                        float* star_dots, long long n, void* stream);
/* same, {guide, dt} read from device memory (float[2]) so that the launch can live in a whole-step CUDA graph */
int b200_cfg_euler_step_dev(float* lat, const float* cond, const float* uncond, const float* guide_dt_dev, float* pred_out,
                            float* star_dots, long long n, void* stream);

/* One FlowUniPCMultistepScheduler.step (shared/utils/fm_solvers_unipc.py:655-740: solver_order 2, bh2, predict_x0,
 * flow_prediction -- WanGP's default sample_solver, any2video.py:518-522) fused with the CFG combine (any2video.py:1701-1722).
 * coef_host8 (HOST pointer) = {sigma_i, ca, cb, cc, cd, pp, pq, pr} from wan2gp_b200/pipeline.py::UniPCSchedule:
 *   v = u + g (c - u);  x0 = x - sigma_i v;  xc = use_corrector ? ca x_last + cb m0 + cc m1 + cd x0 : x;  xn = pp xc + pq x0 + pr m0
 * stores lat <- xn, x_last <- xc, m1 <- x0 (the caller swaps the roles of m0 and m1).  All fp32, n % 4 == 0; star_dots as above. */
int b200_cfg_unipc_step(float* lat, const float* cond, const float* uncond, float guide, float* x_last, const float* m0, float* m1,
                        const float* coef_host8, int use_corrector, float* star_dots, long long n, void* stream);
/* same, {guide, sigma, ca, cb, cc, cd, pp, pq, pr, use_corrector} read from device memory (float[10]): whole-step CUDA graph with the
 * multi-step solvers (UniPC, dpm++) */
int b200_cfg_unipc_step_dev(float* lat, const float* cond, const float* uncond, float* x_last, const float* m0, float* m1,
                            const float* params_dev10, float* star_dots, long long n, void* stream);

/* ---- WanVAE decode (channels-last bf16 activations [T,H,W,C]) ---- */

/* Causal 3-D / 2-D convolution as implicit GEMM (vae.py:43-63 CausalConv3d, :127-133 Conv2d):
 * x bf16 [T,H,W,Cin], w bf16 [Cout, kt*kh*kw, Cin] (tap order t,h,w), bias fp32 [Cout] -> out.
 * Zero padding kh/2, kw/2 in space, (kt-1) frames in FRONT in time.  residual (bf16, same layout as out) is
 * added when non-NULL (ResidualBlock skip, vae.py:273).
 * out_mode 0: bf16 [T,H,W,Cout];  1: time-interleave (vae.py:186-189): Cout = 2C, channel s*C+c of frame t goes to
 * out[(t_off + 2t + s), h, w, c] of a [*,H,W,C] bf16 tensor;  2: planar fp32 [Cout, T, H, W] (decoder head). */
int b200_conv3d_cl(const void* x, const void* w, const float* bias, const void* residual, void* out, int T, int H,
                   int W, int Cin, int Cout, int kt, int kh, int kw, int out_mode, int t_off, void* stream);

/* Fused "conv -> next layer's RMS_norm + SiLU" (Wan VAE ResidualBlock, vae.py:246-273: norm -> SiLU -> CausalConv3d chains).  The
 * epilogue of the producing conv owns all channels of a pixel when Cout is one N tile, so it writes silu(F.normalize(out) * sqrt(C) *
 * gamma) (vae.py:85-103) to norm_out in addition to (out != NULL) or instead of (out == NULL) the raw tensor: the stand-alone norm pass
 * (b200_rms_silu_cl: one read + one write of the activation) is not launched.  b200_conv_norm_fusable() says whether a layer qualifies
 * (row-tiled kernel, Cout of 96 or 192); otherwise the caller uses b200_conv3d_cl + b200_rms_silu_cl. */
int b200_conv_norm_fusable(int W, int Cin, int Cout, int kh, int kw);
int b200_conv3d_cl_norm(const void* x, const void* w, const float* bias, const void* residual, void* out, void* norm_out,
                        const float* gamma, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw, void* stream);
int b200_upconv2x_cl_norm(const void* x, const void* w4, const float* bias, void* out, void* norm_out, const float* gamma, int T, int H,
                          int W, int Cin, int Cout, void* stream);

/* Streaming (time-sliced) causal conv: x_hist = [T + kt - 1, H, W, Cin] (kt-1 history frames of the previous slice, zeros for the first,
 * then this slice's T frames); 'valid' in time, zero-padded in space.  Replaces CausalConv3d.forward with its feature cache
 * (vae.py:55-61) inside the reference's chunked decode loop (vae.py:639-655).  out_mode 0 (+ optional residual, + optional fused
 * norm_out/gamma) or 1 (time_conv interleave). */
int b200_conv3d_cl_stream(const void* x_hist, const void* w, const float* bias, const void* residual, void* out, void* norm_out,
                          const float* gamma, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw, int out_mode, int t_off,
                          void* stream);

/* Decoder head conv Cin -> Cout <= 3, 3x3x3, planar fp32 output (vae.py:505-508 `head`; Hunyuan `conv_out`): the 9 spatial taps are
 * stacked into the GEMM's N (a 3x1x1 conv with 27 -> 32 output channels into the fp32 workspace ws [T,Hg,Wg,32]) and a gather kernel
 * adds the 9 shifted partial sums + bias.  w_stack: bf16 [32][3][Cin], row (dh*3+dw)*Cout + co.  prepadded = 1: x is the replicate-
 * padded [T+2,H+2,W+2,Cin] tensor (Hunyuan), Hg/Wg = H+2/W+2; else zero padding, Hg/Wg = H/W.  ws_bytes >= T*Hg*Wg*128. */
int b200_conv3d_head_cl(const void* x, const void* w_stack, const float* bias, void* ws, long long ws_bytes, float* out, int T, int H,
                        int W, int Cin, int Cout, int prepadded, void* stream);

/* Resample 'upsample2d/3d' spatial part (vae.py:124-133: nearest-exact 2x then Conv2d 3x3 pad 1) as four 2x2 sub-pixel
 * convolutions on the low-resolution input: x bf16 [T,H,W,Cin]; w4 bf16 [4][Cout][4][Cin] = per output parity (py,px) the 3x3
 * taps that hit the same source pixel summed; out bf16 [T,2H,2W,Cout]. */
int b200_upconv2x_cl(const void* x, const void* w4, const float* bias, void* out, int T, int H, int W, int Cin, int Cout,
                     void* stream);

/* y = silu(x / max(||x||_2 over C, 1e-12) * sqrt(C) * gamma) per pixel (vae.py:85-103 RMS_norm + nn.SiLU),
 * bf16 [P, C] -> bf16 [P, C], C % 8 == 0 and C <= 1024; silu=0 skips the activation (AttentionBlock norm, vae.py:293). */
int b200_rms_silu_cl(const void* x, const float* gamma, void* y, long long P, int C, int silu, void* stream);

/* nearest-exact 2x spatial upsample (vae.py:105-111, 124-127), bf16 [T,H,W,C] -> [T,2H,2W,C] */
int b200_upsample2x_cl(const void* x, void* y, int T, int H, int W, int C, void* stream);

/* latent z fp32 [16,T,H,W] -> (z*std+mean) through conv2 1x1x1 (vae.py:631-637) -> bf16 [T,H,W,16] */
int b200_vae_prologue(const float* z, const float* mean, const float* std, const float* w, const float* b, void* out,
                      int T, int H, int W, void* stream);

/* single-head attention, head dim C <= 512 (C % 64 == 0), per frame over N tokens: qkv bf16 [F, N, 3C];
 * out bf16 [F, N, C] (vae.py:276-315 AttentionBlock core, F.scaled_dot_product_attention).
 * causal_frames=1: frame f attends to all tokens of frames 0..f (Hunyuan 1.5 VAE AttnBlock with prepare_causal_attention_mask,
 * hyvideo/vae/hunyuanvideo_15_vae.py:161-214; HunyuanVideo 1.0 VAE mid block, unet_causal_3d_blocks.py:21-30, 728-736);
 * causal_frames=2: every frame attends to all F*N tokens (same mid block with mid_block_causal_attn off); for 1 and 2 the
 * workspace pitch is roundup(F*N,64).
 * workspace: caller-owned, >= N * roundup(N,64) * 6 bytes (fp32 scores + bf16 probabilities of one frame).
 * qkv must have 8 readable (finite) rows after the last frame when N % 8 != 0 (the GEMM extents are rounded up to 8). */
int b200_attention_1head(const void* qkv, void* out, void* workspace, long long workspace_bytes, int F, int N, int C,
                         float scale, int causal_frames, void* stream);

/* frames fp32 planar [3,T,H,W] -> uint8 [3,T,H,W]: round(clamp((clamp(x,-1,1)+1)*127.5,0,255)) (vae.py:18-20) */
int b200_frames_to_u8(const float* x, uint8_t* out, long long n, void* stream);

/* ---- Hunyuan Video 1.5 VAE decode helpers (hyvideo/vae/hunyuanvideo_15_vae.py) ---- */
/* replicate padding of a channels-last bf16 tensor: [T,H,W,C] -> [T+pt, H+2ph, W+2pw, C], pt frames in front (CausalConv3d :137-158) */
int b200_pad_replicate_cl(const void* x, void* y, int T, int H, int W, int C, int pt, int ph, int pw, void* stream);
/* RMS_norm -> SiLU -> replicate padding in one pass for frames [t0, t0+Tc) of x [T,H,W,C]: y = [Tc+pt, H+2ph, W+2pw, C]
 * (ResnetBlock norm -> swish -> CausalConv3d pad, hunyuanvideo_15_vae.py:107-158, 217-250); C % 8 == 0, C <= 1024 */
int b200_rms_silu_pad_cl(const void* x, const float* gamma, void* y, int T, int H, int W, int C, int silu, int t0, int Tc, int pt,
                         int ph, int pw, void* stream);
/* "valid" conv over an explicitly padded input xpad [T+kt-1, H+kh-1, W+kw-1, Cin]; T,H,W = output dims; other arguments as
 * b200_conv3d_cl (out_mode 0, 2, or 3 = fp32 channels-last [T,H,W,Cout] with the optional bf16 residual) */
int b200_conv3d_cl_prepadded(const void* xpad, const void* w, const float* bias, const void* residual, void* out, int T, int H,
                             int W, int Cin, int Cout, int kt, int kh, int kw, int out_mode, void* stream);
/* planar fp32 [C,P] -> channels-last bf16 [P, C*rep] with each channel repeated rep times (z.repeat_interleave, :489-490) */
int b200_planar_to_cl(const float* x, void* y_bf16, int C, long long P, int rep, void* stream);
/* Upsample tail (:309-338): channel -> (time,) space shuffle of the conv output h [T,H,W,F*Co] plus the repeat-interleaved
 * shortcut of x [T,H,W,Ci] -> out bf16 [2T-1 | T, 2H, 2W, Co] */
int b200_hy_upsample_cl(const void* h, const void* x, void* out, int T, int H, int W, int Ci, int Co, int temporal, void* stream);

/* ---- HunyuanVideo 1.0 VAE decode helpers (hyvideo/vae/unet_causal_3d_blocks.py, vae/vae.py) ---- */
/* GroupNorm of a channels-last bf16 tensor [P, C] over all P pixels of the clip (torch.nn.GroupNorm on [B,C,T,H,W],
 * unet_causal_3d_blocks.py:378,399; vae.py:292), pass 1: per-group mean / biased variance folded with the affine parameters
 * into scale_shift[0..C) = gamma[c] / sqrt(var_g + eps) and scale_shift[C..2C) = beta[c] - mean_g * scale[c];
 * workspace >= B200_GROUP_STATS_WS_BYTES(G); C/8 must divide 256; fixed summation order (bit-reproducible) */
#define B200_GROUP_STATS_WS_BYTES(G) (148LL * 8 * (G) * 2 * 4)
int b200_group_stats_cl(const void* x, const float* gamma, const float* beta, float* scale_shift, void* workspace, long long P, int C,
                        int G, float eps, void* stream);
/* pass 2: y = [silu](x * scale[c] + shift[c]) for frames [t0, t0+Tc) of x [T,H,W,C], written replicate-padded as
 * [Tc+pt, H+2ph, W+2pw, C] (pt frames in front, taken from the frames before t0 or frame 0): GroupNorm -> SiLU ->
 * CausalConv3d's F.pad(mode="replicate") (unet_causal_3d_blocks.py:63-66, 455-480) in one pass */
int b200_group_norm_apply_cl(const void* x, const float* scale_shift, void* y, int T, int H, int W, int C, int silu, int t0, int Tc,
                             int pt, int ph, int pw, void* stream);
/* conv over a window of x [Ti,Hi,Wi,Cin] starting at (off_t,off_h,off_w): output pixel (t,h,w), t<T, h<H, w<W, reads taps at
 * x[off_t+t+dt, off_h+h+dh, off_w+w+dw] (zeros beyond the high end of x) and is stored bf16 with element strides (ost_t, ost_h,
 * ost_w); residual (bf16, same addressing as out) may be null.  Uses: one phase of nearest-up-sample + CausalConv3d
 * (UpsampleCausal3D, unet_causal_3d_blocks.py:196-222) on the low-resolution tensor with pre-summed taps; the stride-2 convs of
 * the Wan VAE encoder (Resample down-sampling, vae.py:134-143, 190-212) over space-to-depth / frame-pair views */
int b200_conv3d_cl_view(const void* x, int Ti, int Hi, int Wi, int off_t, int off_h, int off_w, const void* w, const float* bias,
                        const void* residual, void* out, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw, long long ost_t,
                        long long ost_h, long long ost_w, void* stream);

/* ---- Hunyuan 1.5 VAE encode helpers (hyvideo/vae/hunyuanvideo_15_vae.py Downsample :253-296, Encoder :342-430) ---- */
/* Downsample tail: space(-time) -> channel shuffle of the conv output h [T,H,W,Co/F] plus the group-mean shortcut of the shuffled
 * input x [T,H,W,Ci] -> out bf16 [1+(T-1)/2 | T, H/2, W/2, Co]; F = 8 (temporal) or 4; ldh = channel pitch of h (>= Co/F) */
int b200_hy_downsample_cl(const void* h, int ldh, const void* x, void* out, int T, int H, int W, int Ci, int Co, int temporal, void* stream);
/* y[p,c] = mean over g < r of x[p, c*r+g]: bf16 [P,C] -> bf16 [P,C/r] (Encoder.forward shortcut, :424-425) */
int b200_group_mean_cl(const void* x, void* y, long long P, int C, int r, void* stream);

/* seam cross-fade of the tiled VAE decode / encode (vae.py:664-674 blend_v / blend_h, used by spatial_tiled_decode :676-723 and
 * spatial_tiled_encode :841-881): the first min(extent, ...) rows (vertical=1) or columns (vertical=0) of fp32 planar tile
 * b [planes,hb,wb] become a (1 - k/ext) + b (k/ext) with a = the last rows / columns of the neighbour tile a [planes,ha,wa] */
int b200_blend_edge_f32(const float* a, float* b, long long planes, int ha, int wa, int hb, int wb, int extent, int vertical, void* stream);

/* ---- Wan VAE encode helpers (models/wan/modules/vae.py Encoder3d :318-427, Resample :134-143) ---- */
/* planar fp32 [C,P] -> channels-last bf16 [P,Cpad] with zero channels C..Cpad-1 (Cpad % 8 == 0): video [3,T,H,W] -> conv operand */
int b200_planar_to_cl_pad(const float* x, void* y_bf16, int C, long long P, int Cpad, void* stream);
/* space-to-depth 2x2 of a channels-last bf16 tensor: [T,H,W,C] -> [T,H/2,W/2,4C], channel (2p+q)*C+c = x[t,2i+p,2j+q,c] */
int b200_space_to_depth_cl(const void* x, void* y, int T, int H, int W, int C, void* stream);

/* Fused quantise + all-gather of decoded frames (the one collective of the schedule): x fp32 [n] (this
 * rank's frames) -> uint8 written into slot `rank` (byte offset rank*n) of EVERY peer's gather buffer.  peer_bufs: HOST array
 * of n_peers device pointers (peer-mapped, e.g. torch symmetric memory buffer_ptrs); the caller issues the cross-GPU barrier.
 * The reference has no multi-GPU path; this replaces frames_to_u8 + ncclAllGather. */
int b200_frames_to_u8_allgather(const float* x, const uint64_t* peer_bufs, int n_peers, int rank, long long n, void* stream);

/* ---- umT5 text encoder, the step in front of the denoise path (models/wan/modules/t5.py).  Its linear
 * layers go through b200_gemm_bf16; these are the rest ---- */
/* out fp32 [L,dim] = table[ids] (T5Encoder.token_embedding, t5.py:283); table bf16 or fp32 [vocab,dim], ids int64 on the device */
int b200_embed_rows(const long long* ids, const void* table, int table_is_bf16, float* out, int L, int dim, void* stream);
/* T5LayerNorm (t5.py:56-70): out = w * x * rsqrt(mean(x^2) + eps); x fp32 [L,dim]; out bf16 (GEMM operand) or fp32 (final norm) */
int b200_t5_rmsnorm(const float* x, const float* w, void* out, int out_fp32, int L, int dim, float eps, void* stream);
/* out = bf16(a * b), bf16 operands: T5FeedForward's fc1(x) * gelu(gate(x)) (t5.py:145) */
int b200_mul_bf16(const void* a, const void* b, void* out, long long n, void* stream);
/* T5Attention (t5.py:91-128), one sequence of L <= 512 tokens, head dim 64, no score scaling: out[:, h] = softmax(q_h k_h^T +
 * bias_rel[h][j - i + L - 1], keys >= n_valid masked) v_h.  q/k/v bf16 with row stride ld (elements), head h at columns [64h, 64h+64);
 * bias_rel fp32 [heads, 2L-1] = T5RelativeEmbedding (t5.py:219-265) evaluated per offset; out bf16 row stride ldo */
int b200_t5_attention(const void* q, const void* k, const void* v, long long ld, const float* bias_rel, void* out, long long ldo,
                      int L, int heads, int n_valid, void* stream);

/* ---- decoder-only LLM text towers in front of the Hunyuan denoise path (transformers' Qwen2.5-VL / Llama language models as
 * models/hyvideo/text_encoder/text_encoder_1_5.py:86-117, 439-505 and text_encoder/__init__.py run them).  The
 * linear layers go through b200_gemm_bf16, the RMS norm / embedding / gated product through the T5 entries above. ---- */
/* rotate-half RoPE in place (transformers apply_rotary_pos_emb): for every row and each of the first `nheads` 128-wide heads of x bf16
 * [L, ld]: (x1, x2) <- (x1 cos - x2 sin, x2 cos + x1 sin), halves of 64; cos_t / sin_t fp32 [L, 64] */
int b200_rope_half(void* x, long long ld, const float* cos_t, const float* sin_t, int L, int nheads, void* stream);
/* causal grouped-query attention, head dim 128: out[i, h] = softmax_{j <= i}(scale q[i, h] . k[j, h / (q_heads / kv_heads)]) v[j, ...];
 * q bf16 [L, q_heads * 128] row stride ldq, k / v bf16 [L, kv_heads * 128] row stride ldkv (views into one fused q|k|v buffer), out bf16
 * row stride ldo (Qwen2_5_VLAttention / LlamaAttention with a causal mask; right padding never reaches a valid row) */
int b200_causal_gqa_attention(const void* q, const void* k, const void* v, long long ldq, long long ldkv, void* out, long long ldo,
                              int L, int q_heads, int kv_heads, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
