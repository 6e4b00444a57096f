// Training-side kernels of the LECO step:
//   * tn_reduce    : out[N1,N2] += scale * A[M,N1]^T . B[M,N2]  (the skinny LoRA weight gradients
//                    dB = dY^T (s x A^T), dA = (s dY B)^T x; N2 = padded rank <= 64) — HBM-bound:
//                    A is streamed exactly once.
//   * adamw_flat   : fused decoupled-weight-decay Adam over ONE flat LoRA parameter buffer
//                    (replaces torch.optim.AdamW over 384 tensors, train_lora.py:89,280).
//   * guided_step  : classifier-free-guidance combine (train_util.py:163-166) fused with the
//                    DDIM update (scheduling_ddim.py step, eta=0) as one affine map.
//   * leco_loss    : erase/enhance MSE objective (prompt_util.py:107-135) + its gradient
//                    w.r.t. the target prediction, on device (the reference does this on the CPU).
#include "../../include/leco_b200.h"
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace leco {
void count_launch();
bool deterministic();

constexpr int TN_TILE_N1 = 128;
constexpr int TN_ROWS = 32;

__global__ void __launch_bounds__(256)
tn_reduce_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ B,
                 long long ldb, float* __restrict__ out, long long ldo, long long M, int N1, int N2, float scale,
                 int rows_per_split, int transpose_out) {
  pdl_entry();
  __shared__ __align__(16) __nv_bfloat16 sA[TN_ROWS][TN_TILE_N1];
  __shared__ __align__(16) __nv_bfloat16 sB[TN_ROWS][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n0 = blockIdx.x * TN_TILE_N1;
  const long long m_begin = (long long)blockIdx.y * rows_per_split;
  const long long m_end = min(M, m_begin + rows_per_split);
  const int JB = (N2 + 7) / 8;  // b-columns per thread (ty*JB + jj)
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (long long m0 = m_begin; m0 < m_end; m0 += TN_ROWS) {
    // A tile: 32 rows x 128 cols = 512 16-byte vectors, 2 per thread
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int vi = threadIdx.x + t * 256;
      const int r = vi >> 4, cv = vi & 15;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (m0 + r < m_end && n0 + cv * 8 < N1)
        q = *reinterpret_cast<const uint4*>(A + (m0 + r) * lda + n0 + cv * 8);
      *reinterpret_cast<uint4*>(&sA[r][cv * 8]) = q;
    }
    {  // B tile: 32 rows x 64 cols = 256 vectors, 1 per thread
      const int r = threadIdx.x >> 3, cv = threadIdx.x & 7;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (m0 + r < m_end && cv * 8 < N2) q = *reinterpret_cast<const uint4*>(B + (m0 + r) * ldb + cv * 8);
      *reinterpret_cast<uint4*>(&sB[r][cv * 8]) = q;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < TN_ROWS; ++r) {
      const uint2 av = *reinterpret_cast<const uint2*>(&sA[r][tx * 4]);
      const float a0 = bf16_lo(av.x), a1 = bf16_hi(av.x), a2 = bf16_lo(av.y), a3 = bf16_hi(av.y);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (jj < JB) {
          const float b = __bfloat162float(sB[r][ty * JB + jj]);
          acc[0][jj] = fmaf(a0, b, acc[0][jj]);
          acc[1][jj] = fmaf(a1, b, acc[1][jj]);
          acc[2][jj] = fmaf(a2, b, acc[2][jj]);
          acc[3][jj] = fmaf(a3, b, acc[3][jj]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tx * 4 + i;
    if (n >= N1) continue;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = ty * JB + jj;
      if (jj < JB && j < N2)
        atomicAdd(transpose_out ? out + (long long)j * ldo + n : out + (long long)n * ldo + j, acc[i][jj] * scale);
    }
  }
}

// The same reduction on the tensor cores (default; LECO_TN_MMA=0 selects the FMA kernel above).  The contraction runs
// over the ROWS of both operands, so both tiles sit in shared memory "the wrong way round" for mma.sync and are read
// with ldmatrix.trans; rows are padded (272 B / 144 B) so the eight 16-byte row segments of one 8x8 matrix fall in
// different banks.  One block = 128 columns of A x all N2 <= 64 columns of B x one row split; a warp owns 16 columns
// of A.  Tiles of 32 rows stream through a 3-stage cp.async ring.  (mma.sync, not tcgen05: the output is at most
// 128 x 64 per block and the op is bound by reading A once, not by math.)
constexpr int TM_ROWS = 32, TM_STAGES = 3, TM_LDA = TN_TILE_N1 + 8, TM_LDB = 64 + 8;

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
tn_reduce_mma_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ B,
                     long long ldb, float* __restrict__ out, long long ldo, long long M, int N1, int N2, float scale,
                     int rows_per_split, int transpose_out) {
  pdl_entry();
  __shared__ __align__(16) __nv_bfloat16 sA[TM_STAGES][TM_ROWS][TM_LDA];
  __shared__ __align__(16) __nv_bfloat16 sB[TM_STAGES][TM_ROWS][TM_LDB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TN_TILE_N1;
  const long long m_begin = (long long)blockIdx.y * rows_per_split;
  const long long m_end = min(M, m_begin + rows_per_split);
  const int n_stages = (int)((m_end - m_begin + TM_ROWS - 1) / TM_ROWS);
  const int pairs = (N2 + 15) >> 4;          // pairs of 8-wide B column tiles (1..4)

  auto load_stage = [&](int st) {
    const int buf = st % TM_STAGES;
    const long long m0 = m_begin + (long long)st * TM_ROWS;
#pragma unroll
    for (int t = 0; t < 2; ++t) {            // A: 32 rows x 16 vectors
      const int vi = threadIdx.x + t * 256;
      const int r = vi >> 4, cv = vi & 15;
      const bool ok = m0 + r < m_end && n0 + cv * 8 < N1;
      cp_async_16_zfill(&sA[buf][r][cv * 8], ok ? A + (m0 + r) * lda + n0 + cv * 8 : A, ok);
    }
    {                                         // B: 32 rows x 8 vectors
      const int r = threadIdx.x >> 3, cv = threadIdx.x & 7;
      const bool ok = m0 + r < m_end && cv * 8 < N2;
      cp_async_16_zfill(&sB[buf][r][cv * 8], ok ? B + (m0 + r) * ldb + cv * 8 : B, ok);
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int st = 0; st < TM_STAGES - 1; ++st) {
    if (st < n_stages) load_stage(st);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // ldmatrix row addresses: lane l supplies row (l & 7) of matrix (l >> 3); matrices = (k 0-7 | k 8-15) x (col 0-7 | 8-15)
  const int lm_row = (lane & 7) + ((lane >> 4) << 3);      // k offset inside the 16-row step
  const int lm_col = ((lane >> 3) & 1) << 3;               // column offset 0 / 8
  for (int st = 0; st < n_stages; ++st) {
    asm volatile("cp.async.wait_group %0;" ::"n"(TM_STAGES - 2) : "memory");
    __syncthreads();
    if (st + TM_STAGES - 1 < n_stages) load_stage(st + TM_STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int buf = st % TM_STAGES;
#pragma unroll
    for (int ks = 0; ks < TM_ROWS / 16; ++ks) {
      uint32_t a0, a1, a2, a3;
      // A'[m][k] = tile[k][m]: matrices in fragment order (m 0-7,k 0-7) (m 8-15,k 0-7) (m 0-7,k 8-15) (m 8-15,k 8-15)
      ldmatrix_x4_trans(smem_u32(&sA[buf][ks * 16 + (lane & 7) + ((lane >> 4) << 3)][warp * 16 + (((lane >> 3) & 1) << 3)]),
                        a0, a1, a2, a3);
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        if (pr < pairs) {
          uint32_t b0, b1, b2, b3;
          // B'[k][n] = tile[k][n]: (k 0-7,n 0-7) (k 8-15,n 0-7) (k 0-7,n 8-15) (k 8-15,n 8-15)
          ldmatrix_x4_trans(smem_u32(&sB[buf][ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3)][pr * 16 + ((lane >> 4) << 3)]),
                            b0, b1, b2, b3);
          mma_bf16_16816(acc[pr * 2], a0, a1, a2, a3, b0, b1);
          mma_bf16_16816(acc[pr * 2 + 1], a0, a1, a2, a3, b2, b3);
        }
      }
    }
  }
  (void)lm_row;
  (void)lm_col;
  // C fragment: rows (lane >> 2) and +8 of the warp's 16 A-columns, B-columns (lane & 3) * 2 + {0, 1}
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < pairs * 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + warp * 16 + (lane >> 2) + ((e >> 1) << 3);
        const int j = nt * 8 + (lane & 3) * 2 + (e & 1);
        if (n < N1 && j < N2)
          atomicAdd(transpose_out ? out + (long long)j * ldo + n : out + (long long)n * ldo + j, acc[nt][e] * scale);
      }
    }
  }
}

// Flat optimizer step (train_lora.py:89,280 / train_util.py:333-370: adamw, adam, lion) over ONE buffer.
// hyper (device, fp32[8] or fp32[16]): lr, beta1, beta2, eps, weight_decay, step, grad_scale, mode
//   (0 AdamW, 1 Adam with L2 decay, 2 Lion); fp32[16] adds host-computed double-precision scalars
//   [8] lr/(1-beta1^step), [9] sqrt(1-beta2^step), [10] 1-lr*wd, [11] 1-beta1, [12] 1-beta2 (0 = derive here).
// TState = bf16 reproduces the reference bit-for-bit where it can: parameters, gradients and optimizer state all
// live in the training dtype (network.to(dtype) precedes the optimizer, train_lora.py:78-89) and torch's foreach
// implementation rounds to that dtype after EVERY elementwise op (mul, lerp, addcmul, sqrt, div, add, addcdiv);
// `rb()` marks those rounding points.  TState = float keeps fp32 moments and a single rounding of the parameter.
__device__ __forceinline__ float rb(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <typename TState>
__global__ void optim_flat_kernel(__nv_bfloat16* __restrict__ p, float* __restrict__ g, TState* __restrict__ m,
                                  TState* __restrict__ v, const uint8_t* __restrict__ mask,
                                  const float* __restrict__ hyper, int hyper_len, long long n, int zero_grad) {
  pdl_entry();
  constexpr bool kRef = !std::is_same<TState, float>::value;   // reference rounding semantics
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[5],
              gs = hyper[6];
  const int mode = (int)hyper[7];
  float step_size = 0.f, bc2_sqrt = 0.f, decay = 0.f, omb1 = 0.f, omb2 = 0.f;
  if (hyper_len >= 16) {
    step_size = hyper[8];
    bc2_sqrt = hyper[9];
    decay = hyper[10];
    omb1 = hyper[11];
    omb2 = hyper[12];
  }
  if (step_size == 0.f) step_size = lr / (1.0f - powf(b1, step));
  if (bc2_sqrt == 0.f) bc2_sqrt = sqrtf(1.0f - powf(b2, step));
  if (decay == 0.f) decay = 1.0f - lr * wd;
  if (omb1 == 0.f) omb1 = 1.0f - b1;
  if (omb2 == 0.f) omb2 = 1.0f - b2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) {
      if (zero_grad) g[i] = 0.f;
      continue;
    }
    float grad = rb(g[i] * gs);   // the reference's .grad is a bf16 tensor
    float w = __bfloat162float(p[i]);
    const float m0 = (float)m[i];
    if (mode == 2) {
      // lion_pytorch.update_fn: p *= 1-lr*wd; u = sign(m*b1 + (1-b1) g); p -= lr*u; m = m*b2 + (1-b2) g
      if (wd != 0.f) w = kRef ? rb(w * decay) : w * decay;
      float u = kRef ? rb(rb(m0 * b1) + omb1 * grad) : fmaf(omb1, grad, m0 * b1);
      const float sg = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
      w = w - lr * sg;
      const float mn = kRef ? rb(rb(m0 * b2) + omb2 * grad) : fmaf(omb2, grad, m0 * b2);
      m[i] = (TState)mn;
      p[i] = __float2bfloat16(w);
    } else {
      if (mode == 1) {            // torch.optim.Adam: L2 decay enters the gradient
        if (wd != 0.f) grad = kRef ? rb(grad + wd * w) : fmaf(wd, w, grad);
      } else if (wd != 0.f) {     // AdamW: decoupled decay, _foreach_mul_(params, 1 - lr*wd)
        w = kRef ? rb(w * decay) : w * decay;
      }
      float mi, vi, denom;
      if (kRef) {
        mi = rb(m0 + omb1 * (grad - m0));                 // _foreach_lerp_(exp_avgs, grads, 1-beta1)
        vi = rb(rb((float)v[i] * b2) + omb2 * grad * grad);  // _foreach_mul_(beta2); _foreach_addcmul_(g, g, 1-beta2)
        denom = rb(rb(rb(sqrtf(vi)) / bc2_sqrt) + eps);    // _foreach_sqrt; _foreach_div_; _foreach_add_
      } else {
        mi = m0 * b1 + omb1 * grad;
        vi = (float)v[i] * b2 + omb2 * grad * grad;
        denom = sqrtf(vi) / bc2_sqrt + eps;
      }
      m[i] = (TState)mi;
      v[i] = (TState)vi;
      w = w - step_size * (mi / denom);                    // _foreach_addcdiv_(params, exp_avgs, denom, -step_size)
      p[i] = __float2bfloat16(w);
    }
    if (zero_grad) g[i] = 0.f;
  }
}

// fp32 training precision (`train.precision: float32`, config_util.py:62-83; the notebook's setting): the adapters are
// fp32 MASTER parameters with fp32 moments — torch.optim.AdamW / Adam / lion_pytorch on fp32 tensors, no intermediate
// rounding — and every step also writes the bf16 operand copy the tensor-core kernels read (`shadow`).
__global__ void optim_flat_master_kernel(float* __restrict__ p, __nv_bfloat16* __restrict__ shadow,
                                         float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                         const uint8_t* __restrict__ mask, const float* __restrict__ hyper,
                                         long long n, int zero_grad) {
  pdl_entry();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[5],
              gs = hyper[6];
  const int mode = (int)hyper[7];
  float step_size = hyper[8], bc2_sqrt = hyper[9], decay = hyper[10], omb1 = hyper[11], omb2 = hyper[12];
  if (step_size == 0.f) step_size = lr / (1.0f - powf(b1, step));
  if (bc2_sqrt == 0.f) bc2_sqrt = sqrtf(1.0f - powf(b2, step));
  if (decay == 0.f) decay = 1.0f - lr * wd;
  if (omb1 == 0.f) omb1 = 1.0f - b1;
  if (omb2 == 0.f) omb2 = 1.0f - b2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) {
      if (zero_grad) g[i] = 0.f;
      continue;
    }
    float grad = g[i] * gs;
    float w = p[i];
    const float m0 = m[i];
    if (mode == 2) {              // lion_pytorch.update_fn
      if (wd != 0.f) w *= decay;
      const float u = fmaf(omb1, grad, m0 * b1);
      w -= lr * ((u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f));
      m[i] = fmaf(omb2, grad, m0 * b2);
    } else {
      if (mode == 1) {            // torch.optim.Adam: L2 decay enters the gradient
        if (wd != 0.f) grad = fmaf(wd, w, grad);
      } else if (wd != 0.f) {     // AdamW: decoupled decay
        w *= decay;
      }
      const float mi = m0 + omb1 * (grad - m0);          // exp_avg.lerp_(grad, 1 - beta1)
      const float vi = fmaf(omb2 * grad, grad, v[i] * b2);
      m[i] = mi;
      v[i] = vi;
      w -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
    p[i] = w;
    shadow[i] = __float2bfloat16(w);
    if (zero_grad) g[i] = 0.f;
  }
}

// One launch that transposes every [rows, cols] block listed in `tiles` (32x32 tiles of the adapter operands
// ad [Kl,K] / bup [N,Kl] of all fused sites) from the flat parameter buffer into the flat transposed buffer:
// the backward GEMMs need ad^T / bup^T, which only change at the optimizer step.
struct TransposeTile {
  long long src_off, dst_off;   // element offsets of the matrix in the two flat buffers
  int rows, cols, r0, c0;       // matrix shape, tile origin
};
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const __nv_bfloat16* __restrict__ src,
                                                              __nv_bfloat16* __restrict__ dst,
                                                              const TransposeTile* __restrict__ tiles) {
  pdl_entry();
  __shared__ __nv_bfloat16 t[32][33];
  const TransposeTile tt = tiles[blockIdx.x];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = tt.r0 + i, c = tt.c0 + tx;
    if (r < tt.rows && c < tt.cols) t[i][tx] = src[tt.src_off + (long long)r * tt.cols + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = tt.c0 + i, r = tt.r0 + tx;   // dst is [cols, rows]
    if (r < tt.rows && c < tt.cols) dst[tt.dst_off + (long long)c * tt.rows + r] = t[tx][i];
  }
}

// guided = e_u + g (e_c - e_u);  x' = cx * x + ce * guided   (coef = {g, cx, ce})
// eps: fp32 NCHW [2B, chw] (first B uncond, last B cond); x fp32 [B, chw]
__global__ void guided_step_kernel(const float* __restrict__ eps, const float* __restrict__ x,
                                   float* __restrict__ x_out, float* __restrict__ guided_out,
                                   const float* __restrict__ coef, long long half) {
  pdl_entry();
  const float g = coef[0], cx = coef[1], ce = coef[2];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < half;
       i += (long long)gridDim.x * blockDim.x) {
    const float eu = eps[i], ec = eps[half + i];
    const float gd = eu + g * (ec - eu);
    if (guided_out) guided_out[i] = gd;
    if (x_out) x_out[i] = cx * x[i] + ce * gd;
  }
}

// General scheduler step (DDPM / LMS / Euler-ancestral, model_util.py:247-274; DDIM keeps the lean kernel above):
//   guided = e_u + g (e_c - e_u)
//   d      = dx x + dg guided                    -> hist[slot]      (LMS derivative history, 4 deep)
//   x'     = cx x + ce guided + cn noise + sum_j l_j hist[(slot - j) & 3]
// coef (device, fp32[12]) = {g, cx, ce, cn, in_scale, dx, dg, l0, l1, l2, l3, slot}; noise / hist may be NULL.
__global__ void sched_step_kernel(const float* __restrict__ eps, const float* __restrict__ x,
                                  const float* __restrict__ noise, float* __restrict__ hist, float* __restrict__ x_out,
                                  const float* __restrict__ coef, long long half) {
  pdl_entry();
  const float g = coef[0], cx = coef[1], ce = coef[2], cn = coef[3], dx = coef[5], dg = coef[6];
  const float l0 = coef[7], l1 = coef[8], l2 = coef[9], l3 = coef[10];
  const int slot = (int)coef[11];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < half;
       i += (long long)gridDim.x * blockDim.x) {
    const float eu = eps[i], ec = eps[half + i];
    const float gd = eu + g * (ec - eu);
    const float xv = x[i];
    float r = cx * xv + ce * gd;
    if (noise) r = fmaf(cn, noise[i], r);
    if (hist) {
      const float d = dx * xv + dg * gd;
      hist[(long long)slot * half + i] = d;
      r = fmaf(l0, d, r);
      if (l1 != 0.f) r = fmaf(l1, hist[(long long)((slot - 1) & 3) * half + i], r);
      if (l2 != 0.f) r = fmaf(l2, hist[(long long)((slot - 2) & 3) * half + i], r);
      if (l3 != 0.f) r = fmaf(l3, hist[(long long)((slot - 3) & 3) * half + i], r);
    }
    x_out[i] = r;
  }
}

// y = x * coef[idx]   (UNet input scaling 1/sqrt(sigma^2+1) of the sigma schedulers, coefficient on the device)
__global__ void scale_by_dev_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ coef,
                                    int idx, long long n) {
  pdl_entry();
  const float s = coef[idx];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] * s;
}

// loss = mean((t - (nn + sgn*gs*(pp - uu)))^2);  dt = 2 (t - goal) / numel.   single block, deterministic
__global__ void leco_loss_kernel(const float* __restrict__ t, const float* __restrict__ pp,
                                 const float* __restrict__ nn, const float* __restrict__ uu, float sgn_gs,
                                 float* __restrict__ loss, float* __restrict__ dt, long long numel) {
  pdl_entry();
  __shared__ float red[32];
  float s = 0.f;
  const float inv = 1.0f / (float)numel;
  for (long long i = threadIdx.x; i < numel; i += blockDim.x) {
    const float goal = nn[i] + sgn_gs * (pp[i] - uu[i]);
    const float d = t[i] - goal;
    s = fmaf(d, d, s);
    if (dt) dt[i] = 2.0f * d * inv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *loss = s * inv;
  }
}

// fp32 -> bf16 / bf16 -> fp32 flat casts, and scaled fp32 fill (graph-friendly helpers)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

template <typename T>
__global__ void axpby_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, float a, float b,
                             long long n) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = static_cast<T>(a * static_cast<float>(x[i]) + b * static_cast<float>(y[i]));
}

}  // namespace leco

using namespace leco;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

extern "C" int leco_tn_reduce(const void* a, int64_t lda, const void* b, int64_t ldb, float* out, int64_t ldo,
                              int64_t M, int N1, int N2, float scale, int transpose_out, void* stream) {
  LECO_REQUIRE(a && b && out && M > 0, "leco_tn_reduce: null / empty");
  LECO_REQUIRE(N1 % 8 == 0 && N2 % 8 == 0 && N2 <= 64 && lda % 8 == 0 && ldb % 8 == 0,
               "leco_tn_reduce: N1%%8, N2%%8, N2<=64, strides%%8 required (N1=%d N2=%d)", N1, N2);
  // enough blocks to fill the machine: (N1/128) column blocks x M splits of >= 64 rows
  const int col_blocks = (N1 + TN_TILE_N1 - 1) / TN_TILE_N1;
  int splits = (int)((M + 63) / 64);
  const int want = (148 * 4 + col_blocks - 1) / col_blocks;
  if (splits > want) splits = want;
  if (splits < 1) splits = 1;
  // leco_set_deterministic(1): one row split, so every output element receives exactly one (ordered) add per call
  if (deterministic()) splits = 1;
  int rows_per = (int)((M + splits - 1) / splits);
  rows_per = (rows_per + TN_ROWS - 1) / TN_ROWS * TN_ROWS;
  splits = (int)((M + rows_per - 1) / rows_per);
  dim3 grid((N1 + TN_TILE_N1 - 1) / TN_TILE_N1, splits);
  count_launch();
  static const bool use_mma = [] { const char* e = getenv("LECO_TN_MMA"); return !(e && e[0] == '0'); }();
  if (use_mma)
    LECO_LAUNCH(tn_reduce_mma_kernel, grid, 256, 0, STREAM(stream), BF(a), lda, BF(b), ldb, out, ldo, M, N1, N2, scale, rows_per, transpose_out);
  else
    LECO_LAUNCH(tn_reduce_kernel, grid, 256, 0, STREAM(stream), BF(a), lda, BF(b), ldb, out, ldo, M, N1, N2, scale, rows_per, transpose_out);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_optim(void* params_bf16, float* grads, void* exp_avg, void* exp_avg_sq, int state_is_fp32,
                        const void* mask_u8, const float* hyper_dev, int hyper_len, int64_t n, int zero_grad,
                        void* stream) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  if (state_is_fp32)
    LECO_LAUNCH(optim_flat_kernel<float>, (int)blocks, 256, 0, STREAM(stream), BFW(params_bf16), grads,
                reinterpret_cast<float*>(exp_avg), reinterpret_cast<float*>(exp_avg_sq),
                reinterpret_cast<const uint8_t*>(mask_u8), hyper_dev, hyper_len, (long long)n, zero_grad);
  else
    LECO_LAUNCH(optim_flat_kernel<__nv_bfloat16>, (int)blocks, 256, 0, STREAM(stream), BFW(params_bf16), grads,
                BFW(exp_avg), BFW(exp_avg_sq), reinterpret_cast<const uint8_t*>(mask_u8), hyper_dev, hyper_len,
                (long long)n, zero_grad);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_adamw_flat(void* params_bf16, float* grads, void* exp_avg, void* exp_avg_sq, int state_is_fp32,
                               const void* mask_u8, const float* hyper_dev, int64_t n, int zero_grad, void* stream) {
  LECO_REQUIRE(params_bf16 && grads && exp_avg && exp_avg_sq && hyper_dev && n > 0, "leco_adamw_flat: null / empty");
  return launch_optim(params_bf16, grads, exp_avg, exp_avg_sq, state_is_fp32, mask_u8, hyper_dev, 8, n, zero_grad, stream);
}

extern "C" int leco_optim_flat(void* params_bf16, float* grads, void* exp_avg, void* exp_avg_sq, int state_is_fp32,
                               const void* mask_u8, const float* hyper16_dev, int64_t n, int zero_grad, void* stream) {
  LECO_REQUIRE(params_bf16 && grads && exp_avg && exp_avg_sq && hyper16_dev && n > 0, "leco_optim_flat: null / empty");
  return launch_optim(params_bf16, grads, exp_avg, exp_avg_sq, state_is_fp32, mask_u8, hyper16_dev, 16, n, zero_grad, stream);
}

extern "C" int leco_optim_flat_master(float* master, void* shadow_bf16, float* grads, float* exp_avg, float* exp_avg_sq,
                                      const void* mask_u8, const float* hyper16_dev, int64_t n, int zero_grad,
                                      void* stream) {
  LECO_REQUIRE(master && shadow_bf16 && grads && exp_avg && exp_avg_sq && hyper16_dev && n > 0,
               "leco_optim_flat_master: null / empty");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  LECO_LAUNCH(optim_flat_master_kernel, (int)blocks, 256, 0, STREAM(stream), master, BFW(shadow_bf16), grads, exp_avg,
              exp_avg_sq, reinterpret_cast<const uint8_t*>(mask_u8), hyper16_dev, (long long)n, zero_grad);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_transpose_tiles(const void* src, void* dst, const void* tiles, int n_tiles, void* stream) {
  LECO_REQUIRE(src && dst && tiles && n_tiles > 0, "leco_transpose_tiles: null / empty");
  static_assert(sizeof(TransposeTile) == 32, "tile descriptor is 32 bytes (int64 x2 + int32 x4)");
  count_launch();
  LECO_LAUNCH(transpose_tiles_kernel, n_tiles, 256, 0, STREAM(stream), BF(src), BFW(dst),
              reinterpret_cast<const TransposeTile*>(tiles));
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_guided_step(const float* eps_pair, const float* x, float* x_out, float* guided_out,
                                const float* coef_dev, int64_t half_numel, void* stream) {
  LECO_REQUIRE(eps_pair && coef_dev && (x_out || guided_out) && (!x_out || x), "leco_guided_step: null pointer");
  long long blocks = (half_numel + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  count_launch();
  LECO_LAUNCH(guided_step_kernel, (int)blocks, 256, 0, STREAM(stream), eps_pair, x, x_out, guided_out, coef_dev, half_numel);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_sched_step(const float* eps_pair, const float* x, const float* noise, float* hist, float* x_out,
                               const float* coef_dev, int64_t half_numel, void* stream) {
  LECO_REQUIRE(eps_pair && x && x_out && coef_dev && half_numel > 0, "leco_sched_step: null pointer");
  long long blocks = (half_numel + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  count_launch();
  LECO_LAUNCH(sched_step_kernel, (int)blocks, 256, 0, STREAM(stream), eps_pair, x, noise, hist, x_out, coef_dev,
              (long long)half_numel);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_scale_by_dev(const float* x, float* y, const float* coef_dev, int idx, int64_t n, void* stream) {
  LECO_REQUIRE(x && y && coef_dev && n > 0 && idx >= 0, "leco_scale_by_dev: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  count_launch();
  LECO_LAUNCH(scale_by_dev_kernel, (int)blocks, 256, 0, STREAM(stream), x, y, coef_dev, idx, (long long)n);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_loss(const float* target, const float* positive, const float* neutral, const float* uncond,
                         float sign_times_guidance, float* loss_out, float* dtarget, int64_t numel, void* stream) {
  LECO_REQUIRE(target && positive && neutral && uncond && loss_out && numel > 0, "leco_loss: null / empty");
  count_launch();
  LECO_LAUNCH(leco_loss_kernel, 1, 1024, 0, STREAM(stream), target, positive, neutral, uncond, sign_times_guidance, loss_out,
                                                  dtarget, numel);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  LECO_REQUIRE(x && y, "leco_cast: null");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  LECO_LAUNCH(cast_f32_bf16_kernel, (int)blocks, 256, 0, STREAM(stream), x, BFW(y), n);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
  LECO_REQUIRE(x && y, "leco_cast: null");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  LECO_LAUNCH(cast_bf16_f32_kernel, (int)blocks, 256, 0, STREAM(stream), BF(x), y, n);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_axpby(const void* x, const void* y, void* out, float a, float b, int64_t n, int is_fp32,
                          void* stream) {
  LECO_REQUIRE(x && y && out, "leco_axpby: null");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  if (is_fp32)
    LECO_LAUNCH(axpby_kernel<float>, (int)blocks, 256, 0, STREAM(stream), reinterpret_cast<const float*>(x),
                                                                reinterpret_cast<const float*>(y),
                                                                reinterpret_cast<float*>(out), a, b, n);
  else
    LECO_LAUNCH(axpby_kernel<__nv_bfloat16>, (int)blocks, 256, 0, STREAM(stream), BF(x), BF(y), BFW(out), a, b, n);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
