/* mfk.h — C ABI of libmfk (MaPLe-Federated Kernels for NVIDIA B200, sm_100a).
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * tahaspc82442/federated_multi_modal: the MaPLe CustomCLIP forward/backward step and the
 * per-round FedAvg. The reference has NO native code or FFI of its own (pure PyTorch); each
 * entry point below replaces the torch call sites cited next to it (paths relative to the
 * reference tree). INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions:
 *  - plain pointers and sizes, no torch types; all pointers are DEVICE pointers unless noted;
 *  - every call enqueues work on `stream` (a cudaStream_t passed as void*) of the caller's current
 *    device and returns immediately; no internal threads, no allocation (caller passes outputs and
 *    workspaces), no exceptions: return 0 = ok, <0 = argument/shape/alignment error, >0 = cudaError_t;
 *  - matrices are row-major token-major [rows, D]; `ld*` are leading dimensions in ELEMENTS;
 *    bf16 matrices need 16-byte aligned base pointers and ld % 8 == 0;
 *  - "bf16" is __nv_bfloat16, "f16" is __half; activations/weights of the towers are bf16, the
 *    residual stream, LayerNorm statistics, gradients of parameters and the loss head are fp32.
 */
#ifndef MFK_H_
#define MFK_H_

#ifdef __cplusplus
extern "C" {
#endif

int mfk_version(void);
/* profiling aid: device buffer (2 x 64 int64) that receives clock64() stamps of CTA 0 of the fused attention
 * backward (control thread in row 0, first compute warp in row 1); NULL (default) disables it.        */
int mfk_debug_set_attn_trace(void* dev_buf);
/* profiling aid: device buffer (grid x 3 x 64 int64) that receives clock64() stamps of every CTA of the GEMM
 * (row 0 TMA producer, row 1 MMA issuer, row 2 first epilogue warp; tools/gemm_trace.py); NULL disables it. */
int mfk_debug_set_gemm_trace(void* dev_buf);
const char* mfk_error_string(int code);

/* ------------------------------------------------------------------ tensor-core GEMM (tcgen05/TMEM/TMA)
 * out[M,N] = epi( A[M,K] * B[N,K]^T ), A and B bf16 with K contiguous, fp32 accumulation.
 *   epi: (+bias[N] fp32) -> act -> (+residual[M,N] fp32) -> out_f32 and/or out_bf16
 *   act 0: none | 1: QuickGELU x*sigmoid(1.702x), pre-activation optionally stored (out_pre_bf16)
 *       2: multiply by QuickGELU'(aux[M,N] bf16)   (backward of act 1)
 * N % 32 == 0. tile_n: 0 = auto, 128 or 256 = forced full-tile width, 2 = CTA-pair kernel (cta_group::2, M = 256
 * per pair, each CTA holds half of every B tile).
 * splitk_ws (optional, NULL = off): device workspace of splitk_ws_bytes (4096 + 148 * 128 KiB covers every shape),
 * ZERO-FILLED once by the caller and then owned by GEMMs of ONE stream at a time. With it, the tiles of the last
 * partial wave are cut along K over the idle SMs (fp32 partials summed in fixed slice order: results stay
 * bit-reproducible). Kernels that use it spin on device-side counters, so two streams must not share one.
 * Replaces: nn.MultiheadAttention in_proj/out_proj (clip/model.py:274,303-305,350), mlp.c_fc +
 * QuickGELU + c_proj (clip/model.py:276-280,162-164,351), conv1 as GEMM (clip/model.py:484,514),
 * `x @ self.proj` (clip/model.py:569-570), `@ self.text_projection` (trainers/maple.py:76), and
 * autograd's dgrad/wgrad of the same (trainers/maple.py:590).                                        */
int mfk_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                  const float* bias, int act, const void* aux, long long ldaux, const float* residual,
                  long long ldres, float* out_f32, long long ld32, void* out_bf16, long long ld16,
                  void* out_pre_bf16, long long ldpre, int tile_n, void* splitk_ws, long long splitk_ws_bytes,
                  void* stream);

/* out[M,N] fp32 = At^T * Bt, At[K,M] and Bt[K,N] bf16 row-major (MN-major UMMA operands, no transposes).
 * The wgrad form of autograd for the trainable resblocks.11 (trainers/maple.py:472-474,590):
 * dW[out,in] = dY[rows,out]^T X[rows,in].                                                             */
int mfk_gemm_bf16_at_b(const void* At, long long lda, const void* Bt, long long ldb, int M, int N, int K,
                       float* out_f32, long long ld32, void* stream);

/* ------------------------------------------------------------------ attention (head dim 64, T <= 256)
 * qkv[N*T, 3*heads*64] bf16 (q|k|v, head h = columns h*64..h*64+63 of each part) -> out[N*T, heads*64].
 * softmax(q k^T / 8 [+ causal mask]) v per (sequence, head). lse[N,heads,T] (log2 domain) is saved for
 * backward (may be NULL for inference). Replaces nn.MultiheadAttention's core inside
 * ResidualAttentionBlock_MaPLe.attention (clip/model.py:303-305) with the additive causal mask of
 * CLIP.build_attention_mask (clip/model.py:679-685) when causal != 0.                                 */
int mfk_attn_fwd(const void* qkv, void* out, float* lse, int N, int T, int heads, int causal, void* stream);
/* Same contract as mfk_attn_fwd on the tcgen05/TMEM/TMA path (S = QK^T and O = PV as UMMA tiles, softmax one
 * TMEM lane per query row); used for long sequences (the 199-token vision tower).                      */
int mfk_attn_fwd_tc(const void* qkv, void* out, float* lse, int N, int T, int heads, int causal, void* stream);
/* dqkv[N*T, 3*heads*64] bf16 from d_out; delta_ws: N*heads*T floats of scratch. */
int mfk_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                 void* dqkv, int N, int T, int heads, int causal, void* stream);
int mfk_attn_bwd_tc(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                    void* dqkv, int N, int T, int heads, int causal, void* stream);
/* Non-causal backward with Q, K, V, dO of a (sequence, head) resident in smem and S / dP computed once per
 * 128x128 block (five UMMA GEMMs per block, accumulators in TMEM) — the vision-tower path.           */
int mfk_attn_bwd_fused(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                       void* dqkv, int N, int T, int heads, void* stream);

/* ------------------------------------------------------------------ LayerNorm (clip/model.py:153-159)
 * fp32 statistics, eps as given (1e-5), D in {128, 512, 768}. rowidx (int32[M], may be NULL) gathers
 * source rows first — used for ln_post on CLS rows (clip/model.py:567) and ln_final on EOT rows
 * (trainers/maple.py:72-76; LN is row-wise so gather-then-LN == LN-then-gather).
 * Outputs (each optional): y_bf16, y_f32, x_save (gathered fp32 input), mean[M], rstd[M].           */
int mfk_layernorm_fwd(const float* x, const int* rowidx, const float* gamma, const float* beta, void* y_bf16,
                      float* y_f32, float* x_save, float* mean, float* rstd, int M, int D, float eps,
                      void* stream);
/* g_out = (g_in ? g_in : 0) + dLN(dy); optional bf16 copy; optional dgamma/dbeta (fixed-order two-stage
 * reduction). partial_ws: 2*D*mfk_ln_bwd_ctas(M) floats. g_in may alias g_out. `accumulate`: bit 0 adds into
 * dgamma/dbeta; bit 1 DEFERS the second stage — the per-CTA partials stay in partial_ws ([P][2][D], P =
 * mfk_ln_bwd_ctas(M)) and the caller reduces many LayerNorms at once with mfk_partial_reduce_grouped.          */
/* Variants with the deep-prompt splice of ResidualAttentionBlock_MaPLe (clip/model.py:320-349) fused in:
 * forward: rows t in [row0, row0+n_ctx) of every T-row sequence are first overwritten (also in x) with
 * fp16-rounded prompt[t-row0]; backward: the gradient of those rows goes to gprompt [M/T, n_ctx, D] instead of
 * g_out (which gets zeros there) — sum it over the batch with mfk_prompt_splice_bwd on that tensor.            */
int mfk_layernorm_fwd_splice(float* x, const int* rowidx, const float* gamma, const float* beta, void* y_bf16,
                             float* y_f32, float* x_save, float* mean, float* rstd, int M, int D, float eps,
                             const float* prompt, int T, int row0, int n_ctx, void* stream);
int mfk_layernorm_bwd_splice(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                             const float* gamma, const float* g_in, float* g_out, void* g_out_bf16, float* dgamma,
                             float* dbeta, float* partial_ws, int accumulate, int M, int D, float* gprompt, int T,
                             int row0, int n_ctx, void* stream);
int mfk_ln_bwd_ctas(int M);
int mfk_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                      const float* gamma, const float* g_in, float* g_out, void* g_out_bf16, float* dgamma,
                      float* dbeta, float* partial_ws, int accumulate, int M, int D, void* stream);
typedef struct mfk_partial_reduce_problem {
  const float* partial; int P, N; float* out0; float* out1; int accumulate, pad;   /* partial: [P][2][N] */
} mfk_partial_reduce_problem;
int mfk_partial_reduce_grouped(const void* problems_dev, int n_problems, int max_N, void* stream);
/* out[N] (+)= column sums of x[M,N] (bias gradients). partial_ws: 32*N floats. */
int mfk_colsum(const void* x, int is_bf16, long long ld, int M, int N, float* out, float* partial_ws,
               int accumulate, void* stream);

/* ------------------------------------------------------------------ token assembly / prompt splice
 * conv1 16x16 stride 16 as im2col: img fp32 [B,3,S,S] -> bf16 [B*(S/16)^2, 768] (clip/model.py:514-518). */
int mfk_patch_im2col(const float* img, void* out_bf16, int B, int S, void* stream);
/* [cls+pos ; patch tokens+pos ; fp16-rounded shared_ctx] -> ln_pre (clip/model.py:522-544).
 * x0_save (pre-LN, optional) and mean/rstd are kept for backward.                                      */
int mfk_vis_assemble_lnpre(const float* tok, const float* cls, const float* pos, const float* shared_ctx,
                           const float* gamma, const float* beta, float* x0_save, float* x, float* mean,
                           float* rstd, int B, int T, int n_ctx, int D, float eps, void* stream);
/* cat(prefix, ctx, suffix) + positional_embedding, first Te positions (trainers/maple.py:152-166,181-187,54). */
int mfk_text_assemble(const float* prefix, const float* ctx, const float* suffix, const float* pos, float* x,
                      int C, int Te, int n_ctx, int Tfull, int D, void* stream);
/* x[b, row0+j, :] = fp16_round(prompt[j, :]) for all b (clip/model.py:320-349; `.half()` at 327,344). */
int mfk_prompt_splice_fwd(float* x, const float* prompt, int N, int T, int row0, int n_ctx, int D, void* stream);
/* dprompt[j,:] = sum_b (round_fp16 ? fp16_round(g[b,row0+j,:]) : g[...]) in batch order; if zero_rows the
 * rows are then cleared in g and g_bf16 (the spliced-away outputs of the previous layer get no gradient). */
int mfk_prompt_splice_bwd(float* g, void* g_bf16, float* dprompt, int N, int T, int row0, int n_ctx, int D,
                          int round_fp16, int zero_rows, void* stream);
/* The same reduction for `layers` tensors at once (strides in elements): g + l*g_stride -> dprompt + l*dp_stride. */
int mfk_prompt_splice_bwd_batched(float* g, long long g_stride, float* dprompt, long long dp_stride, int layers,
                                  int N, int T, int row0, int n_ctx, int D, int round_fp16, void* stream);
int mfk_scatter_rows(const float* dx, const int* rowidx, float* g, void* g_bf16, int R, int D, void* stream);
/* Dense form for the last block's backward (one consumed row per sequence, rowidx[n] in [n*T, (n+1)*T)): writes ALL
 * N*T rows of g — dx[n,:] at row rowidx[n], zeros elsewhere — in one pass, so g needs no zero fill beforehand. */
int mfk_scatter_rows_dense(const float* dx, const int* rowidx, float* g, int N, int T, int D, void* stream);
/* scatter == 0: dst[r,:] = src[rowidx[r],:]; scatter != 0: dst[rowidx[r],:] = src[r,:] (dst pre-zeroed). Rows are
 * row_bytes long (multiple of 16). Only the CLS row (clip/model.py:567) / EOT row (trainers/maple.py:76) of the
 * last block's output is consumed, so that block's out-proj + MLP run on the gathered rows only.       */
int mfk_gather_rows(const void* src, const int* rowidx, void* dst, int R, long long row_bytes, int scatter,
                    void* stream);
/* out[N, ldo] = in[M, ldi]^T as bf16 (in fp32 or bf16); optional straight bf16 copy. Used to keep
 * K-major copies of weights (dgrad) and of activations/gradients (wgrad of resblocks.11).             */
int mfk_transpose_bf16(const void* in, int in_is_f32, long long ldi, void* out, long long ldo, void* copy,
                       long long ldc, int M, int N, void* stream);
int mfk_cast_f32_bf16(const float* in, void* out, long long n, void* stream);
/* out[rows, 3D] bf16 = [hi | lo | hi] with x = hi + lo: K-concatenated A operand of a split-precision GEMM whose
 * B operand is packed [hi | hi | lo]; one mfk_gemm_bf16 of depth 3D then carries ~16 mantissa bits. Used for the
 * feature heads (`x @ proj`, clip/model.py:569-570; `@ text_projection`, trainers/maple.py:76).        */
int mfk_split_bf16x3(const float* x, void* out_bf16, int rows, int D, void* stream);
/* Single-query attention for the LAST block of a tower: only the CLS / EOT row of each sequence is consumed after
 * it (clip/model.py:567; trainers/maple.py:72-76), so the core runs for that query row alone. rows[n] = global row
 * (n*T + position) of the consumed row; out_rows / d_out_rows bf16 [N, D]; lse_rows fp32 [N, heads].
 * Backward fills ALL rows of dqkv for the N sequences: dQ is zero except the consumed row.                       */
int mfk_attn_rows_fwd(const void* qkv, const int* rows, void* out_rows, float* lse_rows, int N, int T, int heads,
                      int causal, void* stream);
int mfk_attn_rows_bwd(const void* qkv, const int* rows, const void* d_out_rows, const float* lse_rows, void* dqkv,
                      int N, int T, int heads, int causal, void* stream);
/* fp32 mode (parity contract: logits within 1e-3 of the reference's fp32 path; inference only). Every GEMM runs
 * on the same tcgen05 kernel with hi|lo|hi x hi|hi|lo split operands (K -> 3K, fp32 out); these are the pieces in
 * between: exact-sigmoid QuickGELU + split (clip/model.py:162-164), fp32 im2col (clip/model.py:514) and an fp32
 * SIMT attention core (nn.MultiheadAttention, clip/model.py:303-305; qkv fp32 [N*T, 3D] -> out fp32 [N*T, D]). */
int mfk_quickgelu_split_bf16x3(const float* u, void* out_bf16, int rows, int D, void* stream);
int mfk_patch_im2col_f32(const float* img, float* out, int B, int S, void* stream);
int mfk_attn_fwd_f32(const float* qkv, float* out, int N, int T, int heads, int causal, void* stream);
/* fp32 TRAINING mode (cfg PREC = "fp32": the reference calls clip_model.float() and trains the fp32 model,
 * trainers/maple.py:438-439, 590): the backward pieces between the split-operand GEMMs.
 *  - mfk_attn_bwd_f32: dqkv fp32 [N*T, 3D] from d_out fp32 [N*T, D] and the forward's fp32 qkv (softmax recomputed;
 *    stat_ws: 2*N*heads*T floats of scratch for log-sum-exp and delta); autograd of nn.MultiheadAttention's core
 *    (clip/model.py:303-305), T <= 256;
 *  - mfk_dquickgelu_mul_f32: du = dact * QuickGELU'(u) (clip/model.py:162-164);
 *  - mfk_split_bf16x3_rhs: out[rows, 3D] = [hi | hi | lo], the B-side packing for activations (wgrad dY^T X of
 *    resblocks.11 as ONE mfk_gemm_bf16_at_b over 3*rows: the [rows, 3D] buffers are read as [3 rows, D]).   */
int mfk_attn_bwd_f32(const float* qkv, const float* d_out, float* dqkv, float* stat_ws, int N, int T, int heads,
                     int causal, void* stream);
int mfk_dquickgelu_mul_f32(const float* dact, const float* u, float* du, long long n, void* stream);
int mfk_split_bf16x3_rhs(const float* x, void* out_bf16, int rows, int D, void* stream);

/* ------------------------------------------------------------------ prompt-learner projections (fp32)
 * y[m,N] = x[m,K] W[N,K]^T + b (trainers/maple.py:194-215; m = n_ctx) and its backward.               */
int mfk_linear_small_fwd(const float* x, const float* W, const float* b, float* y, int m, int N, int K,
                         void* stream);
int mfk_linear_small_bwd(const float* x, const float* W, const float* dy, float* dW, float* db,
                         const float* dx_add, float* dx, int m, int N, int K, void* stream);
/* Grouped forms: ONE launch (two for the backward: dW/db, dx) over a DEVICE table of problems — all J-1 compound
 * projections + proj_lang_to_vis of a step (trainers/maple.py:194-215). Unused outputs are NULL. Same arithmetic
 * and reduction order as the single-problem calls.                                                     */
typedef struct mfk_small_linear_problem {
  const float* x; const float* W; const float* b; float* y;                  /* forward:  y = x W^T + b          */
  const float* dy; float* dW; float* db; const float* dx_add; float* dx;     /* backward: dW, db, dx = dy W + add */
  int m, N, K, pad;
} mfk_small_linear_problem;
int mfk_linear_small_fwd_grouped(const void* problems_dev, int n_problems, int max_N, void* stream);
int mfk_linear_small_bwd_grouped(const void* problems_dev, int n_problems, int max_m, int max_N, int max_K,
                                 void* stream);
/* Grouped refresh of the bf16 copies of trainable fp32 weights after an optimiser step (one launch):
 * copy[M,N] = bf16(in), out_t[N,M] = bf16(in)^T (the K-major operand of the dgrad GEMM).               */
typedef struct mfk_repack_problem {
  const float* in; void* out_t; void* copy; int M, N;
  int tile0, tiles_n;   /* first global 64x64 tile of this problem (ascending), ceil(N/64) */
} mfk_repack_problem;
int mfk_repack_grouped(const void* problems_dev, int n_problems, int total_tiles, void* stream);

/* ------------------------------------------------------------------ training-time augmentation
 * Dassl build_transform for the reference's yaml (configs/trainers/MaPLe/vit_b16_c2_ep5_batch4_2ctx.yaml:8-13):
 * random_resized_crop (bicubic, antialiased like PIL) -> random_flip -> ToTensor -> normalize, for a whole batch.
 * src uint8 [B,3,H,W]; boxes int32 [B,4] = (top, left, height, width) and flip uint8 [B] are drawn on the host;
 * out fp32 [B,3,S,S]. round_u8 = 1 reproduces the uint8 rounding of the PIL / uint8-tensor pipeline.   */
int mfk_rrc_flip_normalize(const void* src_u8, int B, int H, int W, const int* boxes, const void* flip_u8,
                           const float* mean, const float* stdv, float* out, int S, int round_u8, void* stream);

/* ------------------------------------------------------------------ logits + loss head (trainers/maple.py:325-372)
 * label == NULL: inference, only `logits` [B,C] is written. Otherwise also loss[1], d_img[B,E], d_txt[C,E].
 * ws: mfk_head_workspace_floats(B,C,E) floats.                                                         */
long long mfk_head_workspace_floats(int B, int C, int E);
int mfk_head_forward_backward(const float* img_feat, const float* txt_feat, const float* logit_scale,
                              const long long* label, float* logits, float* loss, float* d_img, float* d_txt,
                              float* ws, int B, int C, int E, void* stream);

/* ------------------------------------------------------------------ FedAvg (trainers/maple_fed.py:309-325)
 * client_ptrs_dev: device array of K device pointers, each to n elements (fp32, or fp16 if in_is_fp16).
 * weights_dev: NULL = uniform (reference behaviour, divisor = K), else K floats (sample counts, divisor =
 * their sum). flags_dev (optional, K ints, caller zeroes): bit0 = client had NaN, bit1 = client had Inf.  */
int mfk_fedavg_reduce(const void* const* client_ptrs_dev, const float* weights_dev, float divisor, int K,
                      long long n, int in_is_fp16, float* out_f32, void* out_f16, int* flags_dev, void* stream);
/* Sharded form for W ranks of one NVLink box: reduces elements [lo, hi) (lo % 4 == 0) of the K fp32 client rows in the
 * same fixed order (bit-identical values) and stores them into EVERY rank's output buffers — out_f32_ptrs_dev /
 * out_f16_ptrs_dev are device arrays of W (peer) pointers to n-element buffers. Replaces safe_average_weights +
 * broadcast_weights traffic (trainers/maple_fed.py:309-339) with (W-1)/W * 10 bytes per element over NVLink.       */
int mfk_fedavg_reduce_scatter(const void* const* client_ptrs_dev, const float* weights_dev, float divisor, int K,
                              long long n, long long lo, long long hi, float* const* out_f32_ptrs_dev,
                              void* const* out_f16_ptrs_dev, int W, void* stream);
/* check_weights_valid: ORs bit0 (NaN) / bit1 (Inf) into *flag_dev. dtype 0 f32, 1 f16, 2 bf16.       */
int mfk_check_finite(const void* p, long long n, int dtype, int* flag_dev, void* stream);

/* ------------------------------------------------------------------ clip_grad_norm_ + SGD (trainers/maple.py:592-598)
 * norm_out[0] = ||g||_2 (fixed-order reduction; partial_ws: 296 floats).
 * hyper_dev = {lr, momentum, dampening, weight_decay, max_norm, nesterov, first_step} as 7 floats.
 * loss_dev / flag_dev (both optional): the update is skipped on the device when *loss_dev is NaN/Inf or
 * *flag_dev != 0 — the reference raises before optim.step() in those cases (trainers/maple.py:556-557, 375-376),
 * leaving parameters and momentum untouched. A NaN gradient norm propagates like torch.clamp does.       */
int mfk_grad_norm(const float* g, long long n, float* partial_ws, float* norm_out, void* stream);
int mfk_sgd_step(float* p, float* g, float* mom, long long n, const float* hyper_dev,
                 const float* total_norm_dev, const float* loss_dev, const int* flag_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFK_H_ */
