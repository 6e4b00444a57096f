/* leco_b200 — C ABI of the B200-native LECO training-step hot path.
 *
 * The reference (p1atdev/LECO) is pure Python; its hot path reaches native code only
 * through torch / diffusers / xformers calls.  Each entry point below replaces one such
 * call site with a hand-written sm_100a kernel.  Citations are to /root/reference.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; leco_last_error() gives the text
 *     (no C++ exception crosses the ABI; the Python wrapper raises RuntimeError).
 *   - all buffers are caller-owned device memory; the library allocates nothing that
 *     outlives a call except cached TMA descriptors.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - activations are bf16, row-major "tokens x channels" ([N*H*W, C], i.e. NHWC).
 */
#ifndef LECO_B200_H_
#define LECO_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* leco_last_error(void);
int leco_abi_version(void);
/* 1: reductions that normally meet in fp32 atomics (LoRA weight-gradient row splits, split-K) run in a fixed order, so
 * a training step is bit-repeatable (slower).  Also switched on by LECO_DETERMINISTIC=1 in the environment. */
int leco_set_deterministic(int on);
/* number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t leco_launch_count(void);
/* fills sm count / cc major / cc minor of the current device */
int leco_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---------------------------------------------------------------------------------
 * leco_gemm_bf16 — tcgen05/TMEM/TMA GEMM   D = epi(alpha * (A.B^T [+ A2.B2^T]))
 * Replaces: nn.Linear / nn.Conv2d forward of every UNet layer AND the LoRA residual
 * `org_forward(x) + lora_up(lora_down(x)) * multiplier * scale` (lora.py:102-106): the
 * LoRA branch is the second K-segment (A2 = scale*down(x) [M,K2], B2 = lora_up weight).
 * mode 0: A is [batch1][batch0][M][K] (element strides given), B is [..][N][K].
 * mode 1: A is an NHWC image [cn][ch][cw][cc]; implicit 3x3 / stride 1 / pad 1 conv,
 *         B is [N][9*cc] with k = (kh*3+kw)*cc + c  (nn.Conv2d weight permuted OHWI).
 * epilogue: + bias[n] + rowbias[m / rows_per_group][n] + residual[m][n];
 *           epilogue==1: GEGLU (diffusers attention.py GEGLU: hidden * gelu(gate)); B keeps the
 *           nn.Linear row order [hidden rows ; gate rows]; D has N/2 columns.
 * ------------------------------------------------------------------------------- */
typedef struct leco_gemm_args {
  const void* a;
  const void* b;
  void* d;
  int32_t mode;
  int32_t M, N, K;
  int64_t lda, ldb, ldd;
  int32_t batch0, batch1;
  int64_t a_bs0, a_bs1, b_bs0, b_bs1, d_bs0, d_bs1;
  int32_t cn, ch, cw, cc;
  const void* a2;
  const void* b2;
  int32_t K2;
  int64_t lda2, ldb2;
  const void* bias;
  const void* rowbias;
  int32_t rows_per_group;
  int64_t ld_rowbias;
  const void* residual;
  int64_t ldr;
  int32_t epilogue;
  float alpha;
  int32_t out_fp32;
  int32_t block_n; /* 0 = heuristic; else 64 / 128 / 160 / 256 */
  int32_t b_rows;  /* rows of B that exist per batch entry (0 = N); rows in [b_rows,N) read as 0 */
  int32_t cta_pair; /* 0 = 1-CTA kernel, 1 = 2-CTA (cta_group::2) kernel: 256 x block_n tiles per CTA pair,
                       2 = let the library choose per shape */
  void* splitk_ws;  /* optional ZEROED fp32 workspace (left zeroed): enables split-K for small-M, long-K problems */
  int64_t splitk_ws_bytes;
  /* in-kernel LoRA (alternative to a2/b2): fl_ad = stacked lora_down [fl_kl][K] (row stride fl_ld_ad), fl_bup =
   * stacked lora_up [N][fl_kl]; D += fl_scale * (A.fl_ad^T).fl_bup^T computed inside the same kernel;
   * fl_t_out (optional) receives fl_scale * A.fl_ad^T as bf16 [M][fl_ld_t] */
  const void* fl_ad;
  const void* fl_bup;
  int32_t fl_kl, fl_rank;
  int64_t fl_ld_ad, fl_ld_bup;
  float fl_scale;
  void* fl_t_out;
  int64_t fl_ld_t;
  int32_t debug_mode; /* perf triage: 1 = no MMAs, 2 = no TMA loads (results are garbage) */
} leco_gemm_args;
int leco_gemm_bf16(const leco_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------
 * Boundary convolutions (diffusers UNet2DConditionModel.conv_in / conv_out, called from
 * train_util.py:156-160).  conv_in: NCHW [n,4,h,w] (fp32 or bf16) -> NHWC bf16, w OIHW.
 * conv_out: NHWC bf16 -> NCHW fp32, w [cout][9][c] (OHWI).  conv_out_bwd: d(input).
 * ------------------------------------------------------------------------------- */
int leco_conv_in(const void* x, int x_is_fp32, const void* w, const void* bias, void* y, int n, int h, int w_,
                 int cout, void* stream);
int leco_conv_out(const void* x, const void* w, const void* bias, float* y, int n, int h, int w_, int c, int cout,
                  void* stream);
/* conv_out as a tensor-core conv: leco_gemm_bf16 (mode 1, weights padded to N = 8, fp32 out) followed by this
 * [n*hw][ld] fp32 (+ bias) -> NCHW fp32 repack */
int leco_cols_to_nchw(const float* y8, int ld, const void* bias, float* out, int n, int hw, int cout, void* stream);
int leco_conv_out_bwd(const float* dy, const void* w, void* dx, int n, int h, int w_, int c, int cout, void* stream);

/* diffusers Timesteps(flip_sin_to_cos=True, freq_shift=0): out[n][dim] = [cos | sin] (bf16) */
int leco_timestep_embedding(const float* t, void* out, int n, int dim, void* stream);

/* elementwise / data-movement glue (bf16, sizes multiples of 8) */
int leco_silu(const void* x, void* y, int64_t numel, void* stream);
int leco_add_inplace(void* y, const void* x, int64_t numel, void* stream);
int leco_geglu_fwd(const void* pre, void* out, int64_t M, int H, void* stream);
int leco_geglu_bwd(const void* pre, const void* dout, void* dpre, int64_t M, int H, void* stream);
int leco_copy_cols(const void* src, int64_t lds, int scol0, void* dst, int64_t ldd, int dcol0, int64_t M, int ncols,
                   void* stream);
int leco_upsample2x(const void* x, void* y, int n, int h, int w, int c, void* stream);
int leco_upsample2x_bwd(const void* dy, void* dx, int n, int h, int w, int c, void* stream);
int leco_im2col_s2(const void* x, void* col, int n, int h, int w, int c, void* stream);
int leco_im2col_s1(const void* x, void* col, int n, int h, int w, int c, void* stream);
int leco_rowgroup_sum(const void* x, void* out, int n, int hw, int c, void* stream);
int leco_col2im_s2(const void* dcol, void* dx, int n, int h, int w, int c, void* stream);
int leco_transpose(const void* in, void* out, int rows, int cols, int rows_pad, int64_t in_ld, int64_t in_bs0,
                   int64_t in_bs1, int64_t out_ld, int64_t out_bs0, int64_t out_bs1, int batch0, int batch1,
                   void* stream);
int leco_softmax_rows(const float* s, void* p, int64_t rows, int n_valid, int n_pad, int64_t ld_s, int64_t ld_p,
                      void* stream);
int leco_softmax_bwd_rows(const void* p, const float* dp, void* ds, int64_t rows, int n_valid, int n_pad,
                          int64_t ld_p, int64_t ld_dp, float scale, void* stream);

/* ---- text-encoder prologue (replaces transformers' CLIPTextModel inside
 * /root/reference/train_util.py:76-77 text_encode and :92-108 text_encode_xl).  The projections and LayerNorms of the
 * encoder run on leco_gemm_bf16 / leco_layer_norm; these three are the pieces the UNet path has no use for:
 *   leco_embed_tokens        out[r] = token_embedding[ids[r]] + position_embedding[r % seq]      (CLIPTextEmbeddings)
 *   leco_softmax_rows_causal leco_softmax_rows with row r limited to columns <= r % sq           (causal mask)
 *   leco_activation          kind 0 SiLU, 1 quick_gelu x*sigmoid(1.702x), 2 erf GELU             (CLIPMLP) */
int leco_embed_tokens(const int* ids, const void* tok, const void* pos, void* out, int64_t rows, int seq, int dim,
                      int vocab, void* stream);
int leco_softmax_rows_causal(const float* s, void* p, int64_t rows, int n_valid, int n_pad, int64_t ld_s,
                             int64_t ld_p, int sq, void* stream);
int leco_activation(const void* x, void* y, int64_t numel, int kind, void* stream);

/* ---------------------------------------------------------------------------------
 * leco_flash_attn_fwd — fused softmax(scale Q K^T) V, head dim <= 64, S/P never leave the SM.
 * Replaces xformers.memory_efficient_attention (enabled at train_lora.py:68).  q/k/v are [rows, ld]
 * bf16 buffers (head h in columns [h*d, h*d+d)); v_t (optional) is V^T [batch][heads][d][skv_pad].
 * ------------------------------------------------------------------------------- */
int leco_flash_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                        const void* v_t, int64_t skv_pad, void* out, int64_t ldo, int batch, int heads, int sq,
                        int skv, int d, float scale, void* stream);

/* leco_flash_attn_fwd_lse: the same forward, additionally writing lse[batch][heads][sq] (fp32, log2 domain) for the backward.
 * leco_flash_attn_bwd: fused backward (dQ, dK, dV; S / P / dP / dS never reach HBM) of the reference's one grad pass
 *   through attention (train_lora.py:279).  dvec = rowsum(dO o O) from leco_attn_bwd_prep; dq_acc = ZEROED fp32
 *   [batch][heads][sq][64] workspace (per-key-tile partial dQ meet there in fp32 reductions), cast by leco_attn_dq_cast. */
int leco_flash_attn_fwd_lse(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                            int64_t ldo, float* lse, int batch, int heads, int sq, int skv, int d, float scale,
                            void* stream);
int leco_attn_bwd_prep(const void* o, int64_t ldo, const void* dout, int64_t lddo, float* dvec, int batch, int heads, int sq,
                       int d, void* stream);
int leco_flash_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout,
                        int64_t lddo, const float* lse, const float* dvec, float* dq_acc, void* dk, int64_t lddk, void* dv,
                        int64_t lddv, int batch, int heads, int sq, int skv, int d, float scale, void* stream);
int leco_attn_dq_cast(const float* dq_acc, void* dq, int64_t lddq, int batch, int heads, int sq, int d, void* stream);

/* ---------------------------------------------------------------------------------
 * GroupNorm(+SiLU) / LayerNorm (diffusers ResnetBlock2D.norm1/2, Transformer2DModel.norm,
 * BasicTransformerBlock.norm1-3), forward and backward-data.  stats = float2 (mean, rstd).
 * ------------------------------------------------------------------------------- */
int64_t leco_group_norm_workspace_bytes(int n, int G);
int leco_group_norm(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw, int C,
                    int G, float eps, int silu, void* workspace, void* stream);
/* single-launch forward (statistics + normalise behind a per-sample grid barrier).  `barriers` is a PERSISTENT caller-owned
 * buffer of leco_group_norm_barrier_bytes(n) bytes that was zero when first used and is never written by anyone else. */
int64_t leco_group_norm_barrier_bytes(int n);
/* leco_group_norm_v2: two launches; the statistics kernel's last block per sample folds the partial sums (fixed order) and
 * writes mean / rstd, the normalise kernel only reads them.  `counters`: persistent buffer as above (4 bytes per sample). */
int leco_group_norm_v2(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw, int C, int G,
                       float eps, int silu, void* workspace, void* counters, void* stream);
/* leco_group_norm_v3: ONE launch with a thread-block cluster per sample (rows held in shared memory, partial sums
 * exchanged through distributed shared memory) when a sample fits one cluster, leco_group_norm_v2 otherwise; same
 * arguments and results (GroupNorm + SiLU of ResnetBlock2D / Transformer2DModel.norm, diffusers 0.20.0). */
int leco_group_norm_v3(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw, int C, int G,
                       float eps, int silu, void* workspace, void* counters, void* stream);
int leco_group_norm_fused(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw, int C,
                          int G, float eps, int silu, void* workspace, void* barriers, void* stream);
int leco_group_norm_bwd(const void* x, const void* dz, void* dx, const void* stats, const void* gamma,
                        const void* beta, int n, int hw, int C, int G, int silu, void* workspace, void* stream);
int leco_layer_norm(const void* x, void* y, void* stats, const void* gamma, const void* beta, int64_t M, int C,
                    float eps, void* stream);
int leco_layer_norm_bwd(const void* x, const void* dy, void* dx, const void* stats, const void* gamma, int64_t M,
                        int C, void* stream);

/* ---------------------------------------------------------------------------------
 * Training-side kernels.
 * leco_tn_reduce: out[N1,N2] (fp32, +=; or out[N2,N1] with transpose_out) = scale * A[M,N1]^T B[M,N2] — the LoRA weight gradients that
 *   autograd derives from lora.py:102-106 (dB = dY^T down(x), dA = (dY B)^T x).
 * leco_adamw_flat: torch.optim.AdamW semantics (train_lora.py:89,280) over one flat buffer;
 *   hyper_dev = fp32[8] {lr, beta1, beta2, eps, weight_decay, step, grad_scale, -}.
 * leco_guided_step: CFG combine (train_util.py:163-166) + DDIM eta=0 update as x' = cx x + ce guided;
 *   coef_dev = fp32[3] {guidance, cx, ce}.
 * leco_loss: prompt_util.py:107-135 erase/enhance MSE and d(loss)/d(target).
 * ------------------------------------------------------------------------------- */
int leco_tn_reduce(const void* a, int64_t lda, const void* b, int64_t ldb, float* out, int64_t ldo, int64_t M, int N1,
                   int N2, float scale, int transpose_out, void* stream);
int leco_adamw_flat(void* params_bf16, float* grads, void* exp_avg, void* exp_avg_sq, int state_is_fp32,
                    const void* mask_u8, const float* hyper_dev, int64_t n, int zero_grad, void* stream);
/* leco_optim_flat: the optimizers of train_util.get_optimizer (train_util.py:333-370) that have a closed elementwise
 *   form, fused over the flat buffer: hyper16_dev = fp32[16] {lr, beta1, beta2, eps, weight_decay, step, grad_scale,
 *   mode (0 adamw | 1 adam | 2 lion), lr/(1-beta1^step), sqrt(1-beta2^step), 1-lr*wd, 1-beta1, 1-beta2, -, -, -}
 *   (entries 8..12 are host double-precision values rounded to fp32; 0 = derive on the device).  With bf16 state the
 *   kernel rounds to bf16 after every elementwise op exactly where torch's foreach optimizer does on bf16 parameters
 *   (the reference keeps parameters, gradients and state in the training dtype, train_lora.py:78-89).
 * leco_transpose_tiles: one launch transposing a list of [rows,cols] blocks between two flat bf16 buffers; tiles =
 *   device array of {int64 src_off, dst_off; int32 rows, cols, r0, c0} (32x32 tiles). */
int leco_optim_flat(void* params_bf16, float* grads, void* exp_avg, void* exp_avg_sq, int state_is_fp32,
                    const void* mask_u8, const float* hyper16_dev, int64_t n, int zero_grad, void* stream);
/* leco_optim_flat_master: the same optimizers for `train.precision: float32` (config_util.py:62-83): fp32 master
 *   parameters and fp32 moments updated as torch does on fp32 tensors (no intermediate rounding); `shadow_bf16` receives
 *   the bf16 copy of every updated parameter = the operand buffer the GEMM kernels read.  hyper16_dev as above. */
int leco_optim_flat_master(float* master, void* shadow_bf16, float* grads, float* exp_avg, float* exp_avg_sq,
                           const void* mask_u8, const float* hyper16_dev, int64_t n, int zero_grad, void* stream);
int leco_transpose_tiles(const void* src, void* dst, const void* tiles, int n_tiles, void* stream);
int leco_guided_step(const float* eps_pair, const float* x, float* x_out, float* guided_out, const float* coef_dev,
                     int64_t half_numel, void* stream);
/* leco_sched_step: the general scheduler update (DDPM / LMS / Euler-ancestral of model_util.py:247-274) fused with the CFG
 *   combine: guided = e_u + g(e_c - e_u); d = dx x + dg guided -> hist[slot] (LMS derivative ring, 4 x half_numel, may
 *   be NULL); x_out = cx x + ce guided + cn noise + sum_j l_j hist[(slot-j)&3].  coef_dev = fp32[12]
 *   {g, cx, ce, cn, in_scale, dx, dg, l0, l1, l2, l3, slot}.  noise may be NULL (cn ignored).
 * leco_scale_by_dev: y = x * coef_dev[idx] — scheduler.scale_model_input (train_util.py:153) with the factor on the device. */
int leco_sched_step(const float* eps_pair, const float* x, const float* noise, float* hist, float* x_out,
                    const float* coef_dev, int64_t half_numel, void* stream);
int leco_scale_by_dev(const float* x, float* y, const float* coef_dev, int idx, int64_t n, void* stream);
int leco_loss(const float* target, const float* positive, const float* neutral, const float* uncond,
              float sign_times_guidance, float* loss_out, float* dtarget, int64_t numel, void* stream);
/* out = a*x + b*y (the DDIM update for drop-in scheduler.step callers, train_util.py:190) */
int leco_axpby(const void* x, const void* y, void* out, float a, float b, int64_t n, int is_fp32, void* stream);
int leco_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
int leco_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LECO_B200_H_ */
