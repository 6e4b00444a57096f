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
 *           epilogue==1: GEGLU (attention.py GEGLU: hidden * gelu(gate)), B rows
 *           interleaved in blocks of 64 (hidden block j, gate block j), D has N/2 cols.
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
} leco_gemm_args;
int leco_gemm_bf16(const leco_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LECO_B200_H_ */
