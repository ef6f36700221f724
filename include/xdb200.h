/* xdb200 -- C ABI of the B200-native (sm_100a) kernels behind the xdiffusion sampling hot path.
 *
 * The reference (Dwaynekj/xdiffusion) has no FFI: its hot path is PyTorch eager.  Each entry point
 * below replaces the chain of ATen ops the reference dispatches at the cited file:line (paths
 * relative to the reference root) and is what a binding for this path would bind
 * (INTEGRATION.md shows the ctypes stub).
 *
 * Conventions: raw DEVICE pointers + explicit sizes / leading dimensions (in ELEMENTS); the caller
 * allocates every output and workspace (no cudaMalloc/free inside -> CUDA-graph capturable); no
 * host synchronisation; `stream` is a cudaStream_t; returns 0 on success, otherwise an XD_ERR_*
 * code with a message in xd_last_error().  bf16 = __nv_bfloat16.  dtype codes: 0 = fp32, 1 = bf16.
 * Activations: 0 none, 1 SiLU, 2 GELU(tanh).
 */
#ifndef XDB200_H
#define XDB200_H
#ifdef __cplusplus
extern "C" {
#endif

#define XD_OK 0
#define XD_ERR_ARG 1
#define XD_ERR_CUDA 2
#define XD_ERR_TMAP 3

int xd_abi_version(void);
const char* xd_last_error(void);

/* ---- dense contractions on tcgen05 / TMEM, operands staged by TMA --------------------------- */

/* out[m,n] = epilogue( sum_k A[m,k] Wt[n,k] + sum_k2 A2[m,k2] Wt[n,K+k2] )
 *   epilogue: v += bias[n]; v = act(v); v *= gate[(m / gate_rows)*gate_ld + n];
 *             v += residual[m*res_ld + n]; store as out_dtype.
 * A, A2, Wt bf16 row-major; K, K2 multiples of 64; bias/gate fp32.  force_bn: 0 = auto, else 64/128/256.
 * Replaces torch.nn.Linear / addmm (score_networks/dit.py:26-40,63-68; layers/attention.py:343-347;
 * layers/mlp.py:30-38; layers/embedding.py:89-94,328-332; layers/resnet.py:143-149), Conv1d k=1
 * (layers/attention.py:75,97) and Conv2d 1x1 (layers/resnet.py:168-170), with the adaLN gate +
 * residual (dit.py:46-51) and GELU-tanh (mlp.py:41-42) fused into the epilogue. */
int xd_gemm_bf16_tc(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                    long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                    int gate_rows, long long gate_ld, const void* residual, int res_dtype, long long res_ld,
                    void* out, int out_dtype, long long out_ld, int force_bn, void* stream);

/* LayerNorm (no affine) + adaLN modulate fused into the A operand of the contraction:
 *   out[m,n] = act( sum_k a[m,k] Wt[n,k] + bias[n] ),  a[m,:] = bf16( LN(x[m,:]) * (1 + scale[m / rows_per_mod]) + shift[m / rows_per_mod] )
 * x fp32 [M, K] with K = 128, 256 or 384 (one row = one shared-memory A panel), N a multiple of 192, out bf16,
 * act none or GELU-tanh.  Same numerics as xd_layernorm_modulate followed by xd_gemm_bf16_tc (bit-identical A operand).
 * Replaces `modulate(norm1(x), shift_msa, scale_msa)` -> `attn.qkv` and `modulate(norm2(x), shift_mlp, scale_mlp)` ->
 * `mlp.fc1` (+ GELU) of a DiT block (score_networks/dit.py:37-59). */
int xd_ln_gemm_bf16_tc(const float* X, long long ldx, const float* shift, const float* scale, long long mod_ld,
                       int rows_per_mod, float eps, const void* Wt, long long ldw, int M, int N, int K,
                       const float* bias, int act, void* out, long long out_ld, void* stream);

/* Fused second half of a DiT block (hidden size D = 384), one CTA pair per 256 token rows, in place on the fp32 stream h:
 *   h1 = h + gate1 * (O Wp^T + bp);  a = bf16(LN(h1) * (1 + scale2) + shift2);  u = bf16(gelu_tanh(a W1^T + b1));
 *   h  = h1 + gate2 * (u W2^T + b2);  stats_out[m] = (mean, rstd) of the new row m (optional; feeds the next LayerNorm).
 * O bf16 [M, D] (attention output), Wp [D, D], W1 [hidden, D], W2 [D, hidden] bf16 row-major, hidden % 128 == 0; gate / shift /
 * scale are fp32 rows of D values, row index m / rows_per_mod, mod_ld floats apart.  The [M, hidden] activation never leaves
 * the SM.  Replaces `x + gate_msa * attn.proj(.)` and `x + gate_mlp * mlp(modulate(norm2(x), shift_mlp, scale_mlp))`
 * (score_networks/dit.py:46-59, layers/attention.py:376-380, layers/mlp.py:30-45).
 * h_out may equal h_in (in place: one CTA pair per 256-row tile).  With distinct buffers and few tiles (a small shard of a
 * strong-scaled batch) the hidden units of a tile are split over `split` = 2, 3 or 4 CTA pairs of one thread-block cluster
 * whose partial fc2 tiles are reduced through distributed shared memory in a fixed order; split = 0 chooses (1 = never). */
int xd_dit_proj_mlp_bf16_tc(const void* O, long long ldo, const void* Wp, const float* bp, const void* W1, const float* b1,
                            const void* W2, const float* b2, int hidden, const float* h_in, float* h_out, long long ldh,
                            int M, int D, const float* gate1, const float* shift2, const float* scale2,
                            const float* gate2, long long mod_ld, int rows_per_mod, float eps, float* stats_out,
                            int split, void* stream);

/* Fused first half of a DiT block (D = 384, 16 tokens per image, head dim 64): LayerNorm-modulate + QKV projection + softmax
 * attention in one kernel, one CTA pair per (256 token rows, group of heads):
 *   a = bf16(LN(h) * (1 + scale) + shift);  [q_i | k_i | v_i] = a Wh_i^T + bias_i;  out[:, 64i : 64i + 64] = softmax(q_i k_i^T * sm_scale) v_i
 * Wh bf16 [heads * 192, D] and bias fp32 [heads * 192] are the reference's qkv Linear re-packed per head as [q | k | v]
 * (64 rows each); stats = (mean, rstd) per row of h as emitted by xd_dit_proj_mlp_bf16_tc, or NULL (computed here).
 * Replaces `attn(modulate(norm1(x), shift_msa, scale_msa))` up to (not including) attn.proj
 * (score_networks/dit.py:46-51, layers/attention.py:350-375). */
int xd_dit_ln_qkv_attn_bf16_tc(const float* h, long long ldh, const float* stats, const float* shift, const float* scale,
                               long long mod_ld, int rows_per_mod, float eps, const void* Wh, const float* bias, int heads,
                               int M, int D, float sm_scale, void* out, long long ldo, void* stream);

/* The two fused DiT kernels with indirect modulation rows: image i reads row mod_rows[i] (device int32 [M / rows_per_mod]) of
 * shift / scale / gate instead of row i.  The adaLN modulation of a DiT depends on (timestep, class label) only, so a sampling
 * loop can evaluate `adaLN_modulation(SiLU(t_emb + y_emb))` (score_networks/dit.py:42-51) once for all N x (classes + 1)
 * pairs when it is built; per step the images then share at most classes + 1 distinct rows per block (L2-resident) and the
 * per-step adaLN GEMM with its 113 MB of fp32 output disappears. */
int xd_dit_ln_qkv_attn_bf16_tc_rows(const float* h, long long ldh, const float* stats, const float* shift, const float* scale,
                                    long long mod_ld, int rows_per_mod, float eps, const void* Wh, const float* bias,
                                    int heads, int M, int D, float sm_scale, void* out, long long ldo, const int* mod_rows,
                                    void* stream);
int xd_dit_proj_mlp_bf16_tc_rows(const void* O, long long ldo, const void* Wp, const float* bp, const void* W1, const float* b1,
                                 const void* W2, const float* b2, int hidden, const float* h_in, float* h_out, long long ldh,
                                 int M, int D, const float* gate1, const float* shift2, const float* scale2,
                                 const float* gate2, long long mod_ld, int rows_per_mod, float eps, float* stats_out,
                                 int split, const int* mod_rows, void* stream);

/* Scratch for split-K (fp32 partial tiles of long contractions over few output tiles, e.g. the 8x8 / 4x4 UNet convs).
 * Device pointer, 16-byte aligned, caller-owned; launches that use it must be ordered on one stream.  Optional: without
 * it every contraction runs unsplit. */
int xd_set_workspace(void* ptr, long long bytes);
/* Split-K (and the automatic hidden-unit split of xd_dit_proj_mlp_bf16_tc, split = 0) makes the summation order (hence the
 * low-order bits of a sample) depend on the batch size; 0 disables both and restores bit-exact batch independence.  Default: on (environment XDB200_SPLITK=0 turns it off). */
int xd_set_split_k(int enabled);

/* Implicit-GEMM conv3x3, stride 1, pad 1, NHWC bf16 (pixel stride ldx), packed weights
 * Wp[Cout][9*C + Cs] (tap-major, then channel; then the optional 1x1 skip weights over Xs).
 * Replaces Conv2d 3x3 (layers/resnet.py:129,155-157; score_networks/unet.py:107-114) and Conv3d
 * (1,3,3) with frames folded into nimg (layers/resnet_3d.py:150-157,199-207); the resblock's
 * `skip_connection(x) + h` (layers/resnet.py:161-170,201) is the second K segment / the residual. */
int xd_conv3x3_bf16_tc(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                       long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                       const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                       long long out_ld, int force_bn, void* stream);

/* GroupNorm statistics from the producer.  The `_qstats` variants compute exactly what xd_gemm_bf16_tc / xd_conv3x3_bf16_tc
 * compute and, when the launch takes the unsplit bf16 TMA-epilogue path without activation (*emitted = 1), also write per
 * block of 32 output rows and per quad of 4 output columns the fp32 (sum, sum of squares) of the stored values:
 *   qstats[(m / 32) * qstats_ld + (n / 4) * 2 + {0, 1}]        (M % 32 == 0, N % 4 == 0, fixed summation order).
 * *emitted = 0 (host int; split-K, narrow tiles, fp32 out): nothing was written, run the stand-alone GroupNorm.
 * xd_groupnorm_apply_quads is the GroupNorm (+ scale/shift, + SiLU) of torch.nn.GroupNorm(32, C) over a tensor whose
 * producers all emitted their quads (a concat buffer: each producer its own channel slice): one streaming pass, the
 * `x.float().mean / var` reduction of the reference (layers/resnet.py:126-128,151-153) never re-reads x. */
int xd_gemm_bf16_tc_qstats(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                           long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                           int gate_rows, long long gate_ld, const void* residual, int res_dtype, long long res_ld,
                           void* out, int out_dtype, long long out_ld, int force_bn, float* qstats, long long qstats_ld,
                           int* emitted, void* stream);
int xd_conv3x3_bf16_tc_qstats(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                              long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                              const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                              long long out_ld, int force_bn, float* qstats, long long qstats_ld, int* emitted,
                              void* stream);
int xd_groupnorm_apply_quads(const void* x, long long ld, int nsamples, int P, int C, int groups, const float* qstats,
                             long long qstats_ld, const float* gamma, const float* beta, const float* scale_shift,
                             long long ss_ld, int ss_div, float eps, int silu, void* out, long long ldo, void* stream);

/* out = GroupNorm(32 groups)(conv3x3(X) + bias) [* (1 + scale) + shift] [-> SiLU]: the first convolution of a residual block and the
 * normalisation that consumes it (layers/resnet.py:129 + 151-153,193-197) in the cheapest way the shape allows.  Few output
 * tiles and a long contraction (the 8x8 / 4x4 levels): the contraction is split over K and the reduce pass normalises -- one
 * CTA per sample, the un-normalised activation never reaches memory.  Unsplit: quad statistics from the epilogue + one
 * streaming pass.  Otherwise: conv, then xd_groupnorm_fused / stats + apply.  tmp: bf16 [nimg*H*W, Cout] scratch; scratch:
 * fp32, max(nimg*H*W / 32 * Cout / 2, nsamples * 64 * xd_groupnorm_slabs(nsamples, P, Cout)) elements. */
int xd_conv3x3_groupnorm_bf16_tc(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Wp, int Cout,
                                 const float* bias, int nsamples, const float* gamma, const float* beta,
                                 const float* scale_shift, long long ss_ld, int ss_div, float eps, int silu, void* tmp,
                                 float* scratch, void* out, long long out_ld, void* stream);

/* CUDA-core twins with the identical contract (on-device cross-check; K not a multiple of 64). */
int xd_gemm_bf16_simt(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                      long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                      int gate_rows, long long gate_ld, const void* residual, int res_dtype, long long res_ld,
                      void* out, int out_dtype, long long out_ld, void* stream);
int xd_conv3x3_bf16_simt(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                         long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                         const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                         long long out_ld, void* stream);

/* First / last UNet convolutions (C_in or C_out tiny: bandwidth-bound, not GEMM-shaped).
 * x fp32 NCHW -> bf16 NHWC (score_networks/unet.py:107-114);  bf16 NHWC -> fp32 NCHW (unet.py:248-255).
 * w is the reference's fp32 [Cout][Cin][3][3] tensor unchanged. */
int xd_conv3x3_in_f32_nchw(const float* x, int nimg, int Cin, int H, int W, const float* w, const float* bias,
                           int Cout, void* out, long long ldo, void* stream);
int xd_conv3x3_out_f32_nchw(const void* X, long long ldx, int nimg, int H, int W, int C, const float* w,
                            const float* bias, int Cout, float* out, void* stream);

/* Fused softmax attention, head_dim 64.  Element (b,h,row,d) at ptr + b*bs + h*hs + row*rs + d.
 * q/k/v are bf16 (qkv_dtype 1) or, for the Tq = Tk = 16 relative-position path, fp32 (qkv_dtype 0; strides
 * then count fp32 elements).  logits = scale * q.k (+ q.relk[h, j-i+Tk-1]); relk fp32 [H][2Tk-1][64] or NULL; scramble=1 stores
 * through the reference's raw (B,H,L,D)->(B,H*D,L) reinterpretation with channel stride o_cs.
 * heads_per_group (0 = H): the head index may carry an outer group g (h = g*heads_per_group + head, e.g. g = pixel of a
 * clip, so that ALL clips run in one launch); `head` selects the relk table and the scrambled row, g adds g*o_gs to the store.
 * Replaces bmm+softmax+bmm at layers/attention.py:182-188 (UNet), :371-375 (DiT), :219-224 (PixArt
 * cross), :551-676 (temporal, relative positions). */
int xd_attention_bf16(const void* q, long long q_bs, long long q_hs, long long q_rs, const void* k,
                      long long k_bs, long long k_hs, long long k_rs, const void* v, long long v_bs,
                      long long v_hs, long long v_rs, void* o, long long o_bs, long long o_hs, long long o_rs,
                      int B, int H, int Tq, int Tk, int head_dim, float scale, const float* relk, int scramble,
                      long long o_cs, int qkv_dtype, int heads_per_group, long long o_gs, void* stream);

/* ---- bandwidth-bound kernels ------------------------------------------------------------------ */

/* GroupNorm over NHWC bf16: a "sample" is P pixel rows of C channels; row(s, p) = (s / inner)*P*inner +
 * s % inner + p*inner  (inner = 1: P consecutive pixels; inner = H*W, P = F: the frames of one pixel of
 * a clip, i.e. the temporal block's GroupNorm over (C/32 x F), layers/attention.py:457-462).  stats fp32
 * [nsamples][xd_groupnorm_slabs()][groups][2] partial (sum, sum of squares), summed in fixed order (no
 * floating-point atomics: a sample's result does not depend on the batch it is in).  apply: y = GN(x)*gamma+beta, then
 * y = y*(1+scale)+shift with [scale | shift] = scale_shift[(sample / ss_div)*ss_ld + ...] if non-NULL,
 * then SiLU if silu.  split=1 writes each row as [hi(C) | lo(C)] bf16 with hi + lo = the fp32 result to
 * 2^-16 (operands of the split-precision temporal-attention GEMM).  (torch.nn.GroupNorm(32,C) + SiLU: layers/resnet.py:126-128,151-153,193-197;
 * layers/attention.py:64; score_networks/unet.py:246-247.) */
int xd_groupnorm_slabs(int nsamples, int P, int C);
int xd_groupnorm_stats(const void* x, long long ld, int nsamples, int P, int C, int groups, int inner,
                       float* stats, void* stream);
int xd_groupnorm_apply(const void* x, long long ld, int nsamples, int P, int C, int groups, const float* stats,
                       const float* gamma, const float* beta, const float* scale_shift, long long ss_ld,
                       int ss_div, float eps, int silu, int inner, int split, void* out, long long ldo, void* stream);

/* Same result in ONE launch when a sample fits a thread-block cluster's shared memory (<= 8 CTAs x ~96 KB):
 * the slab stays in smem, group statistics are reduced across the cluster via DSMEM.  Returns -1 (and does
 * nothing) when the sample is too large; inner == 1 layout only. */
int xd_groupnorm_fused(const void* x, long long ld, int nsamples, int P, int C, int groups, const float* gamma,
                       const float* beta, const float* scale_shift, long long ss_ld, int ss_div, float eps,
                       int silu, void* out, long long ldo, void* stream);

/* GroupNorm(32 groups) over the F frames of every (clip, pixel) sample -- rows ordered (clip, frame, pixel), sample rows at
 * stride HW -- with the split bf16 output [hi(C) | lo(C)] of xd_groupnorm_apply(split = 1): the normalisation in front of the
 * temporal attention's split-precision qkv projection (layers/attention.py:551-600: GroupNorm on the "(b h w) c f" view), one
 * pass, one warp per sample (a lane's C / 32 channels are one group).  Returns -1 for C other than 128 / 256 or F > 16. */
int xd_groupnorm_frames_split(const void* x, long long ld, int B, int F, int HW, int C, const float* gamma,
                              const float* beta, float eps, void* out, long long ldo, void* stream);

/* out(bf16)[m,:] = LayerNorm(x[m,:]; no affine) * (1 + scale[r,:]) + shift[r,:], r = (m / rows_per_mod)*mod_ld.
 * (score_networks/dit.py:16-17,46-51,70-72; pixart.py:20-21,82-92) */
int xd_layernorm_modulate(const float* x, long long ld, int M, int D, const float* shift, const float* scale,
                          long long mod_ld, int rows_per_mod, float eps, void* out, long long ldo, void* stream);

/* Sinusoidal timestep embedding.  mode 0: a = t*freq (layers/utils.py:102-117); mode 1: a =
 * (t*1000/max_time)*freq (layers/embedding.py:66-76); mode 2: t <- atan(exp(-clip(t)/2))/(pi/2) first
 * (embedding.py:131-133).  order 0 = [sin|cos], 1 = [cos|sin].  t: int64 or fp32 [B]. */
int xd_timestep_embed(const void* t, int t_is_i64, int B, const float* freq, int half, int mode, float max_time,
                      float clip_lo, float clip_hi, int order, float* out_f32, void* out_bf16, void* stream);
int xd_act_cast(const void* in, int in_dtype, void* out, int out_dtype, int act, long long n, void* stream);
/* c = table[labels] + temb (DiTCombineEmbeddngs, layers/embedding.py:371-406); silu_out = bf16 SiLU(c). */
int xd_class_combine(const float* table, const long long* labels, const float* temb, int B, int D, float* c_out,
                     void* silu_out, void* stream);
/* The same with temb = temb_table[*idx_dev] for every row: the timestep MLP (layers/embedding.py:325-343) depends on the
 * timestep only, so the sampling loop evaluates it once for all N timesteps when it is built and the per-step conditioning
 * chain shrinks to this kernel + the adaLN GEMM.  idx_dev: the loop index (int32, device).  rows_out (optional, int32 [B]):
 * labels[b] * n_steps + *idx_dev, the row of image b in a (label, step) modulation table (xd_dit_*_rows). */
int xd_class_combine_step(const float* table, const long long* labels, const float* temb_table, const int* idx_dev, int B,
                          int D, float* c_out, void* silu_out, int* rows_out, int n_steps, void* stream);
/* out[0 .. W) = table[*idx_dev][0 .. W) (fp32): the current timestep's row of a per-loop table.  PixArt-alpha's adaLN-single
 * modulation `scale_shift_table[None] + t_block(t)` (score_networks/pixart.py:82-92,253-262) depends on the timestep only, so a
 * sampling loop evaluates it once for all N timesteps; per step this copy replaces the timestep MLP, t_block and the table adds,
 * and every image reads the same row (row stride 0 in the modulation arguments of the other entry points). */
int xd_gather_row_f32(const float* table, long long ld, const int* idx_dev, int W, float* out, void* stream);
/* PatchEmbed im2col (layers/embedding.py:455-457,502-504) and unpatchify (score_networks/dit.py:187-204). */
int xd_patchify(const float* x, int B, int C, int H, int W, int p, void* out_bf16, void* stream);
int xd_unpatchify(const float* y, long long ldy, int B, int C, int H, int W, int p, float* out, void* stream);
int xd_add_rows_periodic(const float* a, const float* b, long long rows, int cols, int period, float* out,
                         void* stream);
int xd_add_table(const float* a, const float* tab, int G, int R, int C, float* out, void* stream);
/* Downsample = AvgPool2d(2) (layers/resnet.py:463), Upsample = nearest x2 (resnet.py:495-499), NHWC bf16. */
int xd_avgpool2x2_nhwc(const void* x, long long ldx, int nimg, int H, int W, int C, void* out, long long ldo,
                       void* stream);
int xd_upsample2x_nhwc(const void* x, long long ldx, int nimg, int H, int W, int C, void* out, long long ldo,
                       void* stream);
/* Strided bf16 row copy: row r = (batch r / rows_per_batch, row r % rows_per_batch), batch strides x_bs / o_bs, row strides
 * ldx / ldo (elements).  Appends the self-attention [q | k | v] rows of every image behind that image's encoder rows:
 * `torch.cat([ek, k], dim=-1)` of QKVAttention (layers/attention.py:166-180). */
int xd_copy_rows_bf16(const void* x, long long ldx, long long x_bs, long long rows, long long rows_per_batch, int C,
                      void* out, long long ldo, long long o_bs, void* stream);
/* Stride-2 3x3 convolution (pad 1) = this im2col + xd_gemm_bf16_tc: out bf16 [nimg * H/2 * W/2, 9 * C], column tap * C + c
 * (tap = 3 dy + dx, the K order of the packed conv weights) = X[n, 2y + dy - 1, 2x + dx - 1, c], zero outside.  Replaces
 * `Conv2d(C, C, 3, stride=2, padding=1)` of DBlock (layers/resnet.py:272-280, Imagen efficient UNet). */
int xd_im2col3x3_s2_nhwc(const void* x, long long ldx, int nimg, int H, int W, int C, void* out, void* stream);
/* out[n, p, :] = x[n, p, :] + b[n, :]: bf16 rows of C channels (row strides ldx / ldo), fp32 per-sample bias (row stride ldb).
 * Replaces `h + Linear(SiLU(timestep_embedding))[..., None, None]` of DBlock / UBlock (layers/resnet.py:303-310,395-402). */
int xd_add_channel_bias_nhwc(const void* x, long long ldx, const float* b, long long ldb, int nsamples, long long P, int C,
                             void* out, long long ldo, void* stream);
/* Video-mask blend of the autoregressive / conditional video loops: x[b, c, f] = mask[b, f] ? x[b, c, f] : x0[b, c, f],
 * in place on x (fp32 [B, C, F, HW]), mask = B x F bytes (non-zero = generate, zero = keep the conditioning frame).
 * Replaces `torch.where(video_mask[:, None, :, None, None], x_t, x0)` before and after every reverse-process step
 * (diffusion/ddpm.py:963-982). */
int xd_blend_frames(float* x, const float* x0, const void* mask, int B, int C, int F, int HW, void* stream);
/* eps = u + w (c - u)  (samplers/ancestral.py:229-231, samplers/ddim.py:69-71) */
int xd_cfg_combine(const float* cond, const float* uncond, float w, float* out, long long n, void* stream);

/* ---- fused sampler step --------------------------------------------------------------------- */

/* Input of a cascade's super-resolution stage (layers/super_resolution.py:47-121): out[b] = [x[b] (nx floats) | a * low[b] +
 * c * z[b] (nl floats)], i.e. `torch.cat([x, q_sample(low_res_x_0, s)], dim=1)` with a = sqrt_alphas_cumprod[s], c =
 * sqrt_one_minus_alphas_cumprod[s].  The reference re-draws the conditioning noise z at EVERY network evaluation: z = row
 * *idx_dev (else idx_host) of an injected table (rows z_step_stride floats apart), or NULL -> in-kernel Philox normals keyed
 * like xd_sampler_step's (distinct stream).  `low` = resized, normalised low-resolution images (timestep-invariant). */
int xd_sr_input(const float* x, const float* low, const float* z, long long z_step_stride, float* out, int B, int nx, int nl,
                float a, float c, const int* idx_dev, int idx_host, unsigned long long seed,
                const unsigned long long* seed_dev, long long elem_offset, void* stream);
/* EDM (Karras et al. 2022) stochastic sampler, fp64 state (samplers/edm.py:85-137), with the EDMPrecond arithmetic
 * (score_networks/edm.py:663-693) folded in.  F = raw network output (fp32) at the state the stage names:
 *   D = c_skip * float(x) + c_out * F
 *   stage 0 (Euler):  d_out = (x_hat - D) / t_div;            x_out = x_hat + h * d_out           (t_div = t_hat)
 *   stage 1 (Heun):   d' = (x_mid - D') / t_div;              x_out = x_hat + h * (0.5 d_in + 0.5 d')   (t_div = t_next)
 *   stage 2:          den_out = D                              (EDMPrecond.forward alone)
 * xin_out (optional) = c_in_next * float(x_out): the next network input.  Every operation is the reference's un-fused IEEE
 * operation in the reference's order: bit-identical to it given the same F. */
int xd_edm_step(int stage, const double* x_hat, const double* x_mid, const double* d_in, const float* F, double* d_out,
                double* x_out, float* den_out, float* xin_out, double t_div, double h, float c_skip, float c_out,
                float c_in_next, long long n, void* stream);
/* x_hat = x + c_noise * z (fp64; z NULL: x_hat = x, nothing written) and xin = c_in * float(x_hat) (fp32, optional):
 * the temporary noise increase of samplers/edm.py:108-120 and the network-input scaling of EDMPrecond.forward. */
int xd_edm_prepare(const double* x, const double* z, double c_noise, double* x_hat, float c_in, float* xin, long long n,
                   void* stream);

/* mode 0 ancestral (samplers/ancestral.py:21-72,189-267), 1 DDIM (samplers/ddim.py:43-123), 2 Euler
 * (samplers/rectified_flow.py:46-84).  coefs fp32 [N][8], row = loop index, read from *idx_dev when
 * non-NULL else idx_host.  z NULL -> in-kernel Philox normals keyed by *seed_dev when non-NULL (graph replay), else seed.  threshold=1 -> dynamic thresholding
 * (utils.py:379-396) with floor(rank)=thr_k, frac=thr_w, cap=thr_c.  In place (out == x) allowed.
 * elem_offset (multiple of 4): index of x[0] in the UNSHARDED batch (first global row * n_per_sample); it is added to the
 * Philox counter so that a batch sharded over ranks draws exactly the noise of the one-GPU run with the same seed. */
int xd_sampler_step(int mode, int form, int pred_v, const float* x, const float* o, const float* z,
                    long long z_step_stride, float* out, const float* coefs, const int* idx_dev, int idx_host,
                    long long n_total, int n_per_sample, int threshold, int thr_k, float thr_w, float thr_c,
                    unsigned long long seed, const unsigned long long* seed_dev, long long elem_offset, void* stream);
/* *idx_dev <- set_to (>= 0), *idx_dev - 1 (set_to == -1) or unchanged (-2); then out_*[b] = tab_*[*idx_dev], b < B
 * (the per-step `t = torch.tensor([idx]*B)` / logsnr lookups of diffusion/ddpm.py:928-955). */
int xd_schedule_advance(int* idx_dev, int set_to, const long long* tab_i64, const float* tab_f32a,
                        const float* tab_f32b, long long* out_i64, float* out_f32a, float* out_f32b, int B,
                        void* stream);
/* (clamp(x,-1,1)+1)/2  (utils.py:62-64) */
int xd_unnormalize(const float* x, float* out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XDB200_H */
