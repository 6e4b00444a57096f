// GroupNorm(+SiLU) and LayerNorm, forward and backward-data, on NHWC / token-major bf16
// activations.  HBM-bound warp-shuffle / shared-atomic reductions (no tensor-core work):
//   GroupNorm: pass 1 = per-(sample,split) partial group sums, pass 2 = normalise (+SiLU);
//   LayerNorm: one warp per row, row held in registers, exact two-pass variance.
// Statistics are fp32; gamma/beta are frozen (no parameter gradients are ever needed:
// LECO trains only the LoRA matrices, train_lora.py:69).
#include <stdlib.h>

#include "../../include/leco_b200.h"
#include "common.cuh"

namespace leco {
void count_launch();

using v8 = uint4;
__device__ __forceinline__ void up8(const v8& q, float (&f)[8]) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x);
  f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z);
  f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ v8 pk8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ float dsilu_f(float y) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-y));
  return s * (1.0f + y * (1.0f - s));
}

constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_MAX_SPLITS = 128;

// ---- pass 1 (forward): partial[n][split][g] = (sum x, sum x^2)
// blockDim.x = vpp * R (vpp = C/8 vector columns, R row lanes); each thread owns one vector column.
// Ordered (deterministic) block reduction of per-thread 8-channel partial sums into per-group sums: every thread parks
// its 2 x 8 values in shared memory ([2][R][C] floats, dynamic), then one thread per group adds them in a fixed order.
__device__ __forceinline__ void gn_block_group_sums(float* red, const float (&s1)[8], const float (&s2)[8], int v, int rl,
                                                    int R, int C, int G, float2* dst) {
  if (rl < R) {
    float* d1 = red + (size_t)rl * C + v * 8;
    float* d2 = red + (size_t)(R + rl) * C + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d1[j] = s1[j];
      d2[j] = s2[j];
    }
  }
  __syncthreads();
  // one (full) warp per group: lane = channel of the group (consecutive banks), rows in order, fixed xor tree.  (One
  // THREAD per group walked R * cpg values serially: ~2 us of the statistics kernels' fixed latency.)
  const int cpg = C / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (warp < nwarps) {
    for (int g = warp; g < G; g += nwarps) {
      float a = 0.f, b = 0.f;
      for (int co = lane; co < cpg; co += 32) {
        const float* p1 = red + g * cpg + co;
        const float* p2 = p1 + (size_t)R * C;
        for (int q = 0; q < R; ++q) {
          a += p1[(size_t)q * C];
          b += p2[(size_t)q * C];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) dst[g] = make_float2(a, b);
    }
  }
}
// Ordered fold of the per-split partials of sample n: warp w handles groups w, w+nwarps, ...; lanes stride the splits,
// then a fixed xor-shuffle tree.  Result (sum, sum2) per group in `out` (shared).
__device__ __forceinline__ void gn_fold_partials(const float2* partial, int n, int splits, int G, float (*out)[2]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // FULL warps only
  for (int g = (warp < nwarps ? warp : G); g < G; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int q = lane; q < splits; q += 32) {
      const float2 pp = partial[((size_t)n * splits + q) * G + g];
      a += pp.x;
      b += pp.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      out[g][0] = a;
      out[g][1] = b;
    }
  }
}

__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ partial, int hw, int C,
                                int G, int vpp, int splits) {
  pdl_entry();
  extern __shared__ float gn_red[];
  const int n = blockIdx.x, sp = blockIdx.y;
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  const int rows_per = (hw + splits - 1) / splits;
  const int r0 = sp * rows_per, r1 = min(hw, r0 + rows_per);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (rl < R) {
    const __nv_bfloat16* base = x + ((size_t)n * hw) * C + v * 8;
    int r = r0 + rl;
    for (; r + 3 * R < r1; r += 4 * R) {  // four independent 16-byte loads in flight
      const v8 qa = *reinterpret_cast<const v8*>(base + (size_t)r * C);
      const v8 qb = *reinterpret_cast<const v8*>(base + (size_t)(r + R) * C);
      const v8 qc = *reinterpret_cast<const v8*>(base + (size_t)(r + 2 * R) * C);
      const v8 qd = *reinterpret_cast<const v8*>(base + (size_t)(r + 3 * R) * C);
      float f[8], g[8];
      up8(qa, f);
      up8(qb, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += f[j] + g[j];
        s2[j] = fmaf(f[j], f[j], fmaf(g[j], g[j], s2[j]));
      }
      up8(qc, f);
      up8(qd, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += f[j] + g[j];
        s2[j] = fmaf(f[j], f[j], fmaf(g[j], g[j], s2[j]));
      }
    }
    for (; r < r1; r += R) {
      float f[8];
      up8(*reinterpret_cast<const v8*>(base + (size_t)r * C), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += f[j];
        s2[j] = fmaf(f[j], f[j], s2[j]);
      }
    }
  }
  gn_block_group_sums(gn_red, s1, s2, v, rl, R, C, G, partial + ((size_t)n * splits + sp) * G);
}

// ---- pass 2 (forward): y = (x-mean)*rstd*gamma+beta [silu]; also writes stats[n][g] = (mean, rstd)
// Same thread layout as pass 1 (blockDim.x = vpp * R, one 8-channel vector column per thread), so the per-channel
// scale = rstd*gamma and shift = beta - mean*scale are loop invariants: the row loop is load, 8 FMA, SiLU, store.
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                const float2* __restrict__ partial, float2* __restrict__ stats,
                                const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                int hw, int C, int G, int splits, float eps, int silu, int vpp) {
  pdl_entry();
  __shared__ float2 ms[GN_MAX_GROUPS];
  __shared__ float acc[GN_MAX_GROUPS][2];
  const int n = blockIdx.x;
  const int cpg = C / G;
  gn_fold_partials(partial, n, splits, G, acc);   // ordered: deterministic
  __syncthreads();
  if (threadIdx.x < G) {
    const float cnt = (float)hw * (float)cpg;
    const float mean = acc[threadIdx.x][0] / cnt;
    const float var = fmaxf(acc[threadIdx.x][1] / cnt - mean * mean, 0.f);
    const float2 r = make_float2(mean, rsqrtf(var + eps));
    ms[threadIdx.x] = r;
    if (blockIdx.y == 0) stats[(size_t)n * G + threadIdx.x] = r;
  }
  __syncthreads();
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  if (rl >= R) return;
  float sc[8], sh[8];
  {
    float gm[8], bt[8];
    up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
    up8(__ldg(reinterpret_cast<const v8*>(beta + v * 8)), bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 m = ms[(v * 8 + j) / cpg];
      sc[j] = m.y * gm[j];
      sh[j] = fmaf(-m.x, sc[j], bt[j]);
    }
  }
  const int rows_per = (hw + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(hw, r0 + rows_per);
  const __nv_bfloat16* xb = x + ((size_t)n * hw) * C + v * 8;
  __nv_bfloat16* yb = y + ((size_t)n * hw) * C + v * 8;
  int r = r0 + rl;
  for (; r + R < r1; r += 2 * R) {  // two independent 16-byte loads in flight
    float f[8], g[8];
    const v8 qa = *reinterpret_cast<const v8*>(xb + (size_t)r * C);
    const v8 qb = *reinterpret_cast<const v8*>(xb + (size_t)(r + R) * C);
    up8(qa, f);
    up8(qb, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      g[j] = fmaf(g[j], sc[j], sh[j]);
      if (silu) {
        f[j] = silu_fast(f[j]);
        g[j] = silu_fast(g[j]);
      }
    }
    *reinterpret_cast<v8*>(yb + (size_t)r * C) = pk8(f);
    *reinterpret_cast<v8*>(yb + (size_t)(r + R) * C) = pk8(g);
  }
  for (; r < r1; r += R) {
    float f[8];
    up8(*reinterpret_cast<const v8*>(xb + (size_t)r * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (silu) f[j] = silu_fast(f[j]);
    }
    *reinterpret_cast<v8*>(yb + (size_t)r * C) = pk8(f);
  }
}

// ---- fused forward: statistics + normalise(+SiLU) in ONE launch.
// grid (n, bps): block (n, sp) owns rows [sp*rows_per, ...) of sample n.  Phase 1 = its partial group sums (ordered
// shared-memory reduction: deterministic) -> partial[n][sp][g]; a per-sample arrive/wait barrier in global memory
// (self-resetting generation counter, so it survives CUDA-graph replays); phase 2 = every block folds the sample's
// partials in a fixed order and normalises its own rows, which are still L2/L1 resident.  All blocks of a launch must be
// co-resident (n*bps <= a fraction of 148 x 8 blocks of 256 threads, enforced by the host); PDL is safe: dependents only
// launch once every block here has started.
struct GnBarrier {
  unsigned int count, gen;
};
__global__ void __launch_bounds__(384)
gn_fused_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, float2* __restrict__ partial,
                float2* __restrict__ stats, GnBarrier* __restrict__ bars, const __nv_bfloat16* __restrict__ gamma,
                const __nv_bfloat16* __restrict__ beta, int hw, int C, int G, float eps, int silu, int vpp) {
  pdl_entry();
  extern __shared__ float gn_red[];                 // [2][R][C]: per-thread partial sums, then reused
  __shared__ float2 ms[GN_MAX_GROUPS];
  const int n = blockIdx.x, sp = blockIdx.y, bps = gridDim.y;
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  const int rows_per = (hw + bps - 1) / bps;
  const int r0 = sp * rows_per, r1 = min(hw, r0 + rows_per);
  const int cpg = C / G;
  const __nv_bfloat16* xb = x + ((size_t)n * hw) * C + v * 8;
  // ---------------- phase 1: partial sums of this block's rows
  {
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (rl < R) {
      int r = r0 + rl;
      for (; r + 3 * R < r1; r += 4 * R) {  // four independent 16-byte loads in flight
        const v8 qa = *reinterpret_cast<const v8*>(xb + (size_t)r * C);
        const v8 qb = *reinterpret_cast<const v8*>(xb + (size_t)(r + R) * C);
        const v8 qc = *reinterpret_cast<const v8*>(xb + (size_t)(r + 2 * R) * C);
        const v8 qd = *reinterpret_cast<const v8*>(xb + (size_t)(r + 3 * R) * C);
        float f[8], g[8];
        up8(qa, f);
        up8(qb, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += f[j] + g[j];
          s2[j] = fmaf(f[j], f[j], fmaf(g[j], g[j], s2[j]));
        }
        up8(qc, f);
        up8(qd, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += f[j] + g[j];
          s2[j] = fmaf(f[j], f[j], fmaf(g[j], g[j], s2[j]));
        }
      }
      for (; r < r1; r += R) {
        float f[8];
        up8(*reinterpret_cast<const v8*>(xb + (size_t)r * C), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += f[j];
          s2[j] = fmaf(f[j], f[j], s2[j]);
        }
      }
      float* d1 = gn_red + (size_t)rl * C + v * 8;
      float* d2 = gn_red + (size_t)(R + rl) * C + v * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d1[j] = s1[j];
        d2[j] = s2[j];
      }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {   // fixed summation order: deterministic
      float a = 0.f, b = 0.f;
      for (int q = 0; q < R; ++q)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
          a += gn_red[(size_t)q * C + c];
          b += gn_red[(size_t)(R + q) * C + c];
        }
      partial[((size_t)n * bps + sp) * G + g] = make_float2(a, b);
    }
  }
  // ---------------- per-sample barrier
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* genp = &bars[n].gen;
    const unsigned int gen = *genp;
    __threadfence();
    if (atomicAdd(&bars[n].count, 1u) == (unsigned)bps - 1u) {
      bars[n].count = 0;
      __threadfence();
      atomicAdd(&bars[n].gen, 1u);
    } else {
      const long long t0 = clock64();
      while (*genp == gen) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) {
          printf("leco_b200: group-norm grid barrier watchdog (sample %d block %d of %d)\n", n, sp, bps);
          __trap();
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
  // ---------------- phase 2: statistics (every block, fixed order) + normalise own rows
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // FULL warps only
    for (int g = (warp < nwarps ? warp : G); g < G; g += nwarps) {
      float a = 0.f, b = 0.f;
      for (int q = lane; q < bps; q += 32) {
        const float2 pp = __ldcg(&partial[((size_t)n * bps + q) * G + g]);
        a += pp.x;
        b += pp.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) {
        const float cnt = (float)hw * (float)cpg;
        const float mean = a / cnt;
        const float var = fmaxf(b / cnt - mean * mean, 0.f);
        const float2 rr = make_float2(mean, rsqrtf(var + eps));
        ms[g] = rr;
        if (sp == 0) stats[(size_t)n * G + g] = rr;
      }
    }
  }
  __syncthreads();
  if (rl >= R) return;
  float sc[8], sh[8];
  {
    float gm[8], bt[8];
    up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
    up8(__ldg(reinterpret_cast<const v8*>(beta + v * 8)), bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 m = ms[(v * 8 + j) / cpg];
      sc[j] = m.y * gm[j];
      sh[j] = fmaf(-m.x, sc[j], bt[j]);
    }
  }
  __nv_bfloat16* yb = y + ((size_t)n * hw) * C + v * 8;
  int r = r0 + rl;
  for (; r + R < r1; r += 2 * R) {  // two independent 16-byte loads in flight
    float f[8], g[8];
    const v8 qa = *reinterpret_cast<const v8*>(xb + (size_t)r * C);
    const v8 qb = *reinterpret_cast<const v8*>(xb + (size_t)(r + R) * C);
    up8(qa, f);
    up8(qb, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      g[j] = fmaf(g[j], sc[j], sh[j]);
      if (silu) {
        f[j] = silu_fast(f[j]);
        g[j] = silu_fast(g[j]);
      }
    }
    *reinterpret_cast<v8*>(yb + (size_t)r * C) = pk8(f);
    *reinterpret_cast<v8*>(yb + (size_t)(r + R) * C) = pk8(g);
  }
  for (; r < r1; r += R) {
    float f[8];
    up8(*reinterpret_cast<const v8*>(xb + (size_t)r * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (silu) f[j] = silu_fast(f[j]);
    }
    *reinterpret_cast<v8*>(yb + (size_t)r * C) = pk8(f);
  }
}

// ---- forward v2: statistics kernel that FINISHES the statistics (the last block of a sample to arrive folds the
// per-split partials in a fixed order and writes mean / rstd), so the normalise kernel starts with 2 x G floats instead
// of re-folding splits x G partials in every one of its ~700 blocks (ncu: that prologue was most of its 18 us).
// counters: persistent zero-initialised unsigned int per sample (self-resetting).
__global__ void gn_stats_v2_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ partial,
                                   float2* __restrict__ stats, unsigned int* __restrict__ counters, int hw, int C,
                                   int G, int vpp, int splits, float eps) {
  pdl_entry();
  extern __shared__ float gn_red[];
  __shared__ float facc[GN_MAX_GROUPS][2];
  __shared__ unsigned int s_ticket;
  const int n = blockIdx.x, sp = blockIdx.y;
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  const int rows_per = (hw + splits - 1) / splits;
  const int r0 = sp * rows_per, r1 = min(hw, r0 + rows_per);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (rl < R) {
    const __nv_bfloat16* base = x + ((size_t)n * hw) * C + v * 8;
    int r = r0 + rl;
    for (; r + 7 * R < r1; r += 8 * R) {  // eight independent 16-byte loads in flight per thread
      v8 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = *reinterpret_cast<const v8*>(base + (size_t)(r + u * R) * C);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float f[8];
        up8(q[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += f[j];
          s2[j] = fmaf(f[j], f[j], s2[j]);
        }
      }
    }
    for (; r + R < r1; r += 2 * R) {
      const v8 qa = *reinterpret_cast<const v8*>(base + (size_t)r * C);
      const v8 qb = *reinterpret_cast<const v8*>(base + (size_t)(r + R) * C);
      float f[8], g[8];
      up8(qa, f);
      up8(qb, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += f[j] + g[j];
        s2[j] = fmaf(f[j], f[j], fmaf(g[j], g[j], s2[j]));
      }
    }
    for (; r < r1; r += R) {
      float f[8];
      up8(*reinterpret_cast<const v8*>(base + (size_t)r * C), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += f[j];
        s2[j] = fmaf(f[j], f[j], s2[j]);
      }
    }
  }
  gn_block_group_sums(gn_red, s1, s2, v, rl, R, C, G, partial + ((size_t)n * splits + sp) * G);
  // ---- the last block of this sample finishes the statistics
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&counters[n], 1u);
  __syncthreads();
  if (s_ticket != (unsigned)splits - 1u) return;
  __threadfence();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // FULL warps only
    for (int g = (warp < nwarps ? warp : G); g < G; g += nwarps) {
      float a = 0.f, b = 0.f;
      for (int q = lane; q < splits; q += 32) {
        const float2 pp = __ldcg(&partial[((size_t)n * splits + q) * G + g]);
        a += pp.x;
        b += pp.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) {
        facc[g][0] = a;
        facc[g][1] = b;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const float cnt = (float)hw * (float)(C / G);
    const float mean = facc[threadIdx.x][0] / cnt;
    const float var = fmaxf(facc[threadIdx.x][1] / cnt - mean * mean, 0.f);
    stats[(size_t)n * G + threadIdx.x] = make_float2(mean, rsqrtf(var + eps));
  }
  if (threadIdx.x == 0) counters[n] = 0;   // ready for the next launch (stream-ordered)
}

__global__ void gn_apply_v2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   const float2* __restrict__ stats, const __nv_bfloat16* __restrict__ gamma,
                                   const __nv_bfloat16* __restrict__ beta, int hw, int C, int G, int silu, int vpp) {
  pdl_entry();
  const int n = blockIdx.x;
  const int cpg = C / G;
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  if (rl >= R) return;
  float sc[8], sh[8];
  {
    float gm[8], bt[8];
    up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
    up8(__ldg(reinterpret_cast<const v8*>(beta + v * 8)), bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 m = __ldg(&stats[(size_t)n * G + (v * 8 + j) / cpg]);
      sc[j] = m.y * gm[j];
      sh[j] = fmaf(-m.x, sc[j], bt[j]);
    }
  }
  const int rows_per = (hw + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(hw, r0 + rows_per);
  const __nv_bfloat16* xb = x + ((size_t)n * hw) * C + v * 8;
  __nv_bfloat16* yb = y + ((size_t)n * hw) * C + v * 8;
  int r = r0 + rl;
  for (; r + 3 * R < r1; r += 4 * R) {  // four independent 16-byte loads in flight
    v8 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const v8*>(xb + (size_t)(r + u * R) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      up8(q[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = fmaf(f[j], sc[j], sh[j]);
        if (silu) f[j] = silu_fast(f[j]);
      }
      *reinterpret_cast<v8*>(yb + (size_t)(r + u * R) * C) = pk8(f);
    }
  }
  for (; r < r1; r += R) {
    float f[8];
    up8(*reinterpret_cast<const v8*>(xb + (size_t)r * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (silu) f[j] = silu_fast(f[j]);
    }
    *reinterpret_cast<v8*>(yb + (size_t)r * C) = pk8(f);
  }
}

// ---- forward v3: ONE launch, one thread-block CLUSTER per sample (in-graph timeline, profiles/r2x_timeline_sd21.md:
// the statistics kernel costs 9-12 us at every size, 8 CTAs or 256, because its tail is a chain of global round trips
// -- partial store, __threadfence, ticket atomic, __ldcg of the partials, statistics store -- and the normalise kernel
// then waits for it and re-reads the tensor from L2: 16 us per GroupNorm, 61 GroupNorms per UNet forward).
// Here CTA `rank` of the cluster keeps its rows [rank*rows_per, ...) of the sample in SHARED memory while summing them,
// the per-CTA (sum, sum^2) per group are exchanged through distributed shared memory behind the hardware cluster barrier
// (no global atomics, no fences), every CTA folds the CL partials in rank order (deterministic) and normalises its rows
// straight out of shared memory: x is read once, y written once, one launch.
constexpr int GNC_MAX_CL = 16;
constexpr int GNC_THREADS = 512;
__device__ __forceinline__ uint32_t gnc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t gnc_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gnc_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// SiLU with ONE special-function op: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx: relative error 2^-11, below the
// bf16 rounding of the result).  The exp + reciprocal form costs two; with a sample's rows concentrated on one cluster
// the MUFU pipe (16 / clk / SM) is what the normalise phase waits for.
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// second barrier of the kernel: pure "I have finished reading my peers' shared memory", no data is published
__device__ __forceinline__ void gnc_cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
// streaming 16-byte load (read once, never written by this kernel)
__device__ __forceinline__ v8 gnc_ldg_stream(const void* p) {
  v8 q;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
               : "l"(p));
  return q;
}
__device__ __forceinline__ void gnc_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// float2 at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ float2 gnc_ld_peer_f2(const float2* local, uint32_t rank) {
  uint32_t ra;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ra) : "memory");
  return v;
}

// grid (CL, parts, n), cluster (CL, 1, 1): cluster (part, n) owns channels [part*Cs, (part+1)*Cs) = Gs whole groups of
// sample n.  Splitting the channels over `parts` clusters puts 2-4x as many SMs on a batch of 4 samples (the SiLU is
// MUFU work and the copies are per-SM L2 bandwidth: both scale with the number of SMs that take part).
__global__ void __launch_bounds__(GNC_THREADS)
gn_cluster_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, float2* __restrict__ stats,
                  const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta, int hw, int C, int Cs,
                  int G, int Gs, float eps, int silu, int vpp, int rows_per) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t gnc_smem[];     // [2][R][Cs] floats (reduction scratch) | [rows_per][Cs] bf16
  __shared__ float2 part[GN_MAX_GROUPS];                  // this CTA's (sum, sum^2) per group; read by the whole cluster
  __shared__ float2 gath[GNC_MAX_CL * GN_MAX_GROUPS];     // [CL][Gs]
  __shared__ float2 ms[GN_MAX_GROUPS];
  const int n = blockIdx.z, c0 = blockIdx.y * Cs;
  const uint32_t rank = gnc_cluster_rank(), CL = gnc_cluster_size();
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;    // FULL warps only
  const int r0 = (int)rank * rows_per, r1 = min(hw, r0 + rows_per);
  const int cpg = Cs / Gs;
  float* red = reinterpret_cast<float*>(gnc_smem);
  v8* slab = reinterpret_cast<v8*>(gnc_smem + (size_t)2 * R * Cs * sizeof(float));
  // ---------------- phase 1: this CTA's rows -> shared memory, per-thread channel sums on the way
  {
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (rl < R) {
      const __nv_bfloat16* base = x + ((size_t)n * hw) * C + c0 + v * 8;
      int r = r0 + rl;
      for (; r + 7 * R < r1; r += 8 * R) {  // eight independent 16-byte loads in flight per thread
        // (volatile asm + barrier: with plain loads the compiler sank every load down to its shared-memory store, one
        // load in flight per thread -- first GPU run of this kernel: 19 us for 10 MB, no better than two launches)
        v8 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = gnc_ldg_stream(base + (size_t)(r + u * R) * C);
        asm volatile("" ::: "memory");
#pragma unroll
        for (int u = 0; u < 8; ++u) slab[(size_t)(r + u * R - r0) * vpp + v] = q[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float f[8];
          up8(q[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1[j] += f[j];
            s2[j] = fmaf(f[j], f[j], s2[j]);
          }
        }
      }
      for (; r + R < r1; r += 2 * R) {
        const v8 qa = gnc_ldg_stream(base + (size_t)r * C);
        const v8 qb = gnc_ldg_stream(base + (size_t)(r + R) * C);
        asm volatile("" ::: "memory");
        slab[(size_t)(r - r0) * vpp + v] = qa;
        slab[(size_t)(r + R - r0) * vpp + v] = qb;
        float f[8], g[8];
        up8(qa, f);
        up8(qb, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += f[j] + g[j];
          s2[j] = fmaf(f[j], f[j], fmaf(g[j], g[j], s2[j]));
        }
      }
      for (; r < r1; r += R) {
        const v8 q = gnc_ldg_stream(base + (size_t)r * C);
        slab[(size_t)(r - r0) * vpp + v] = q;
        float f[8];
        up8(q, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += f[j];
          s2[j] = fmaf(f[j], f[j], s2[j]);
        }
      }
      float* d1 = red + (size_t)rl * Cs + v * 8;
      float* d2 = red + (size_t)(R + rl) * Cs + v * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d1[j] = s1[j];
        d2[j] = s2[j];
      }
    }
  }
  __syncthreads();
  // ---------------- phase 2: ordered fold of the R x cpg parked values of each group (lanes stride, fixed xor tree)
  if (warp < nwarps) {
    for (int g = warp; g < Gs; g += nwarps) {
      float a = 0.f, b = 0.f;
      for (int co = lane; co < cpg; co += 32) {       // lane = channel of the group (consecutive banks), rows in order
        const float* p1 = red + g * cpg + co;
        const float* p2 = p1 + (size_t)R * Cs;
        for (int q = 0; q < R; ++q) {
          a += p1[(size_t)q * Cs];
          b += p2[(size_t)q * Cs];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) part[g] = make_float2(a, b);
    }
  }
  __syncthreads();
  gnc_cluster_arrive();          // release: part[] of every CTA is visible cluster-wide after the wait
  gnc_cluster_wait();
  for (int t = threadIdx.x; t < (int)CL * Gs; t += blockDim.x) {
    const int w = t / Gs, g = t - w * Gs;
    gath[t] = gnc_ld_peer_f2(&part[g], (uint32_t)w);
  }
  __syncthreads();
  gnc_cluster_arrive_relaxed();  // "done reading my peers": nobody leaves before every reader has (wait at the end)
  if (threadIdx.x < Gs) {
    float a = 0.f, b = 0.f;
    for (uint32_t w = 0; w < CL; ++w) {           // rank order: every CTA of the cluster computes the same bits
      a += gath[w * Gs + threadIdx.x].x;
      b += gath[w * Gs + threadIdx.x].y;
    }
    const float cnt = (float)hw * (float)cpg;
    const float mean = a / cnt;
    const float var = fmaxf(b / cnt - mean * mean, 0.f);
    const float2 m = make_float2(mean, rsqrtf(var + eps));
    ms[threadIdx.x] = m;
    if (rank == 0) stats[(size_t)n * G + blockIdx.y * Gs + threadIdx.x] = m;
  }
  __syncthreads();
  // ---------------- phase 3: normalise (+SiLU) out of shared memory
  if (rl < R) {
    float sc[8], sh[8];
    {
      float gm[8], bt[8];
      up8(__ldg(reinterpret_cast<const v8*>(gamma + c0 + v * 8)), gm);
      up8(__ldg(reinterpret_cast<const v8*>(beta + c0 + v * 8)), bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 m = ms[(v * 8 + j) / cpg];
        sc[j] = m.y * gm[j];
        sh[j] = fmaf(-m.x, sc[j], bt[j]);
      }
    }
    __nv_bfloat16* yb = y + ((size_t)n * hw) * C + c0 + v * 8;
    int r = r0 + rl;
    for (; r + 3 * R < r1; r += 4 * R) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        up8(slab[(size_t)(r + u * R - r0) * vpp + v], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaf(f[j], sc[j], sh[j]);
          if (silu) f[j] = silu_tanh(f[j]);
        }
        *reinterpret_cast<v8*>(yb + (size_t)(r + u * R) * C) = pk8(f);
      }
    }
    for (; r < r1; r += R) {
      float f[8];
      up8(slab[(size_t)(r - r0) * vpp + v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = fmaf(f[j], sc[j], sh[j]);
        if (silu) f[j] = silu_tanh(f[j]);
      }
      *reinterpret_cast<v8*>(yb + (size_t)r * C) = pk8(f);
    }
  }
  gnc_cluster_wait();            // pairs with the second arrive: my part[] may be freed now
}

// ---- backward pass 1: partial[n][split][g] = (sum g, sum g*xhat) with g = dy*gamma (dy through SiLU')
__global__ void gn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                                    const float2* __restrict__ stats, const __nv_bfloat16* __restrict__ gamma,
                                    const __nv_bfloat16* __restrict__ beta, float2* __restrict__ partial, int hw,
                                    int C, int G, int vpp, int splits, int silu) {
  pdl_entry();
  extern __shared__ float gn_red[];
  __shared__ float2 ms[GN_MAX_GROUPS];
  const int n = blockIdx.x, sp = blockIdx.y;
  for (int i = threadIdx.x; i < G; i += blockDim.x) ms[i] = stats[(size_t)n * G + i];
  __syncthreads();
  const int R = blockDim.x / vpp;
  const int v = threadIdx.x % vpp, rl = threadIdx.x / vpp;
  const int rows_per = (hw + splits - 1) / splits;
  const int r0 = sp * rows_per, r1 = min(hw, r0 + rows_per);
  const int cpg = C / G;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (rl < R) {
    float gm[8], bt[8];
    up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
    up8(__ldg(reinterpret_cast<const v8*>(beta + v * 8)), bt);
    const size_t off = ((size_t)n * hw) * C + v * 8;
    float2 mj[8];                                   // this thread's 8 channels: (mean, rstd) of their groups
#pragma unroll
    for (int j = 0; j < 8; ++j) mj[j] = ms[(v * 8 + j) / cpg];
    auto row = [&](const v8& qx, const v8& qd) {
      float f[8], d[8];
      up8(qx, f);
      up8(qd, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (f[j] - mj[j].x) * mj[j].y;
        float dy = d[j];
        if (silu) dy *= dsilu_f(xh * gm[j] + bt[j]);
        const float g = dy * gm[j];
        s1[j] += g;
        s2[j] = fmaf(g, xh, s2[j]);
      }
    };
    int r = r0 + rl;
    for (; r + 3 * R < r1; r += 4 * R) {            // eight independent 16-byte loads in flight per thread
      v8 qx[4], qd[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                 // (volatile asm: the compiler sinks plain loads to their first use)
        qx[u] = gnc_ldg_stream(x + off + (size_t)(r + u * R) * C);
        qd[u] = gnc_ldg_stream(dz + off + (size_t)(r + u * R) * C);
      }
      asm volatile("" ::: "memory");
#pragma unroll
      for (int u = 0; u < 4; ++u) row(qx[u], qd[u]);
    }
    for (; r < r1; r += R)
      row(*reinterpret_cast<const v8*>(x + off + (size_t)r * C), *reinterpret_cast<const v8*>(dz + off + (size_t)r * C));
  }
  gn_block_group_sums(gn_red, s1, s2, v, rl, R, C, G, partial + ((size_t)n * splits + sp) * G);
}

// ---- backward pass 2: dx = rstd * (g - mean(g) - xhat * mean(g*xhat))
__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                                    __nv_bfloat16* __restrict__ dx, const float2* __restrict__ stats,
                                    const float2* __restrict__ partial, const __nv_bfloat16* __restrict__ gamma,
                                    const __nv_bfloat16* __restrict__ beta, int hw, int C, int G, int splits,
                                    int silu) {
  pdl_entry();
  __shared__ float2 ms[GN_MAX_GROUPS];
  __shared__ float2 gs[GN_MAX_GROUPS];
  const int n = blockIdx.x;
  const int cpg = C / G;
  __shared__ float acc[GN_MAX_GROUPS][2];
  gn_fold_partials(partial, n, splits, G, acc);   // ordered: deterministic
  __syncthreads();
  if (threadIdx.x < G) {
    const float cnt = (float)hw * (float)cpg;
    gs[threadIdx.x] = make_float2(acc[threadIdx.x][0] / cnt, acc[threadIdx.x][1] / cnt);
    ms[threadIdx.x] = stats[(size_t)n * G + threadIdx.x];
  }
  __syncthreads();
  const int vpp = C / 8;
  const long long total = (long long)hw * vpp;
  const size_t base = (size_t)n * hw * C;
  auto one = [&](long long i, const v8& qx, const v8& qd) {
    const int v = (int)((unsigned long long)i % (unsigned)vpp);
    float f[8], d[8], gm[8], bt[8];
    up8(qx, f);
    up8(qd, d);
    up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
    up8(__ldg(reinterpret_cast<const v8*>(beta + v * 8)), bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (v * 8 + j) / cpg;
      const float2 m = ms[g];
      const float xh = (f[j] - m.x) * m.y;
      float dy = d[j];
      if (silu) dy *= dsilu_f(xh * gm[j] + bt[j]);
      const float gg = dy * gm[j];
      f[j] = m.y * (gg - gs[g].x - xh * gs[g].y);
    }
    *reinterpret_cast<v8*>(dx + base + i * 8) = pk8(f);
  };
  const long long stride = (long long)gridDim.y * blockDim.x;
  long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x;
  for (; i + stride < total; i += 2 * stride) {     // two vectors (four 16-byte loads) in flight per thread
    const v8 xa = gnc_ldg_stream(x + base + i * 8), da = gnc_ldg_stream(dz + base + i * 8);
    const v8 xb = gnc_ldg_stream(x + base + (i + stride) * 8), db = gnc_ldg_stream(dz + base + (i + stride) * 8);
    asm volatile("" ::: "memory");
    one(i, xa, da);
    one(i + stride, xb, db);
  }
  if (i < total) one(i, gnc_ldg_stream(x + base + i * 8), gnc_ldg_stream(dz + base + i * 8));
}

// ---------------------------------------------------------------- LayerNorm
constexpr int LN_MAX_VEC = 8;  // per lane -> C <= 2048

// LayerNorm forward.  A row is handled by a GROUP of G lanes (G = 8 / 16 / 32, chosen so that a lane holds at most VPL
// 8-channel vectors), i.e. 32/G rows per warp: all of a lane's loads are issued before the first use (VPL x 16 B in
// flight per lane), the register footprint is VPL*8 floats (high occupancy), and no lane idles on narrow rows (C = 320
// is 40 vectors: a full-warp-per-row mapping leaves 24 lanes empty in its second round).  Exact two-pass variance.
template <int VPL>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, float2* __restrict__ stats,
              const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta, long long M, int C,
              float eps, int G) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int sub = lane % G, grp = lane / G, rpw = 32 / G;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int vpr = C / 8;
  const float inv_c = 1.0f / (float)C;
  for (long long r0 = warp * rpw; r0 < M; r0 += nwarps * rpw) {
    const long long r = r0 + grp;
    const bool row_ok = r < M;
    v8 q[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = sub + G * i;
      q[i] = (row_ok && v < vpr) ? *reinterpret_cast<const v8*>(x + r * C + v * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      up8(q[i], f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (sub + G * i < vpr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[i][j] - mean;
          qq = fmaf(d, d, qq);
        }
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    const float rstd = rsqrtf(qq * inv_c + eps);
    if (sub == 0 && stats && row_ok) stats[r] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = sub + G * i;
      if (row_ok && v < vpr) {
        float gm[8], bt[8], o[8];
        up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
        up8(__ldg(reinterpret_cast<const v8*>(beta + v * 8)), bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * gm[j] + bt[j];
        *reinterpret_cast<v8*>(y + r * C + v * 8) = pk8(o);
      }
    }
  }
}

__global__ void ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                              __nv_bfloat16* __restrict__ dx, const float2* __restrict__ stats,
                              const __nv_bfloat16* __restrict__ gamma, long long M, int C) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int vpr = C / 8;
  for (long long r = warp; r < M; r += nwarps) {
    const float2 m = stats[r];
    float xh[LN_MAX_VEC][8], g[LN_MAX_VEC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int v = lane + 32 * i;
      if (v < vpr) {
        float f[8], d[8], gm[8];
        up8(*reinterpret_cast<const v8*>(x + r * C + v * 8), f);
        up8(*reinterpret_cast<const v8*>(dy + r * C + v * 8), d);
        up8(__ldg(reinterpret_cast<const v8*>(gamma + v * 8)), gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (f[j] - m.x) * m.y;
          g[i][j] = d[j] * gm[j];
          s1 += g[i][j];
          s2 = fmaf(g[i][j], xh[i][j], s2);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / (float)C, m2 = s2 / (float)C;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
      const int v = lane + 32 * i;
      if (v < vpr) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = m.y * (g[i][j] - m1 - xh[i][j] * m2);
        *reinterpret_cast<v8*>(dx + r * C + v * 8) = pk8(o);
      }
    }
  }
}

static int gn_launch_cfg(int hw, int C, int* vpp, int* threads, int* splits) {
  *vpp = C / 8;
  if (*vpp > 1024) return -1;
  int R = 256 / *vpp;
  if (R < 1) R = 1;
  *threads = *vpp * R;
  // >= `rows` pixel rows per block; many small blocks keep enough loads in flight.  LECO_GN_ROWS / LECO_GN_SPLITS
  // override the defaults (tuning hooks of tests/gpu_checks/kernel_cases.py::case_norm_perf)
  static const int env_rows = [] { const char* e = getenv("LECO_GN_ROWS"); return e ? atoi(e) : 32; }();
  static const int env_max = [] { const char* e = getenv("LECO_GN_SPLITS"); return e ? atoi(e) : 64; }();
  int s = hw / (env_rows > 0 ? env_rows : 32);
  if (s < 1) s = 1;
  if (s > env_max) s = env_max;
  if (s > GN_MAX_SPLITS) s = GN_MAX_SPLITS;
  *splits = s;
  return 0;
}

}  // namespace leco

using namespace leco;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

// workspace: float2[n * GN_MAX_SPLITS * G]  (leco_group_norm_workspace_bytes)
extern "C" int64_t leco_group_norm_workspace_bytes(int n, int G) { return (int64_t)n * GN_MAX_SPLITS * G * 8; }

extern "C" int leco_group_norm(const void* x, void* y, void* stats /*float2[n*G]*/, const void* gamma,
                               const void* beta, int n, int hw, int C, int G, float eps, int silu,
                               void* workspace, void* stream) {
  LECO_REQUIRE(x && y && stats && gamma && beta && workspace, "leco_group_norm: null pointer");
  LECO_REQUIRE(C % 8 == 0 && G > 0 && G <= GN_MAX_GROUPS && C % G == 0, "leco_group_norm: C=%d G=%d unsupported", C, G);
  int vpp, threads, splits;
  LECO_REQUIRE(gn_launch_cfg(hw, C, &vpp, &threads, &splits) == 0, "leco_group_norm: C=%d too wide", C);
  count_launch();
  LECO_LAUNCH(gn_stats_kernel, dim3(n, splits), threads, (size_t)2 * (threads / vpp) * C * sizeof(float), STREAM(stream), BF(x),
              reinterpret_cast<float2*>(workspace), hw, C, G, vpp, splits);
  LECO_CHECK_CUDA(cudaGetLastError());
  // pass 2: ~8 blocks per SM over the whole batch, but at least 4 pixel rows per thread so the prologue amortises
  const int R = threads / vpp;
  static const int env_bpsm = [] { const char* e = getenv("LECO_GN_APPLY_BPSM"); return e ? atoi(e) : 8; }();
  static const int env_rpt = [] { const char* e = getenv("LECO_GN_APPLY_RPT"); return e ? atoi(e) : 4; }();
  int gy = (env_bpsm * 148) / (n < 1 ? 1 : n);
  if (gy > hw / (env_rpt * R)) gy = hw / (env_rpt * R);
  if (gy < 1) gy = 1;
  count_launch();
  LECO_LAUNCH(gn_apply_kernel, dim3(n, gy), threads, 0, STREAM(stream), BF(x), BFW(y), reinterpret_cast<const float2*>(workspace),
                                                          reinterpret_cast<float2*>(stats), BF(gamma), BF(beta), hw,
                                                          C, G, splits, eps, silu, vpp);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Fused single-launch GroupNorm forward.  workspace: float2[n * GN_MAX_SPLITS * G] (any contents); barriers:
// PERSISTENT device buffer of >= n * 8 bytes that was ZERO when first used (the per-sample barrier state lives there
// across launches / graph replays: leco_group_norm_barrier_bytes).
extern "C" int64_t leco_group_norm_barrier_bytes(int n) { return (int64_t)n * (int64_t)sizeof(GnBarrier); }

extern "C" int leco_group_norm_fused(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n,
                                     int hw, int C, int G, float eps, int silu, void* workspace, void* barriers,
                                     void* stream) {
  LECO_REQUIRE(x && y && stats && gamma && beta && workspace && barriers, "leco_group_norm_fused: null pointer");
  LECO_REQUIRE(C % 8 == 0 && G > 0 && G <= GN_MAX_GROUPS && C % G == 0, "leco_group_norm_fused: C=%d G=%d unsupported", C, G);
  int vpp, threads, splits;
  LECO_REQUIRE(gn_launch_cfg(hw, C, &vpp, &threads, &splits) == 0 && threads <= 384, "leco_group_norm_fused: C=%d too wide", C);
  const int R = threads / vpp;
  const size_t smem = (size_t)2 * R * C * sizeof(float);
  LECO_REQUIRE(smem <= 48 * 1024, "leco_group_norm_fused: C=%d needs %zu B of shared memory", C, smem);
  // every block of the launch must be resident at once (grid barrier): ask the runtime how many fit
  int per_sm = 0;
  LECO_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, threads, smem));
  int bps = splits;
  const int cap = (sm_count() * per_sm) / (n < 1 ? 1 : n);
  if (bps > cap) bps = cap;
  LECO_REQUIRE(bps >= 1, "leco_group_norm_fused: batch %d too large for the single-launch kernel", n);
  count_launch();
  LECO_LAUNCH(gn_fused_kernel, dim3(n, bps), threads, smem, STREAM(stream), BF(x), BFW(y),
              reinterpret_cast<float2*>(workspace), reinterpret_cast<float2*>(stats),
              reinterpret_cast<GnBarrier*>(barriers), BF(gamma), BF(beta), hw, C, G, eps, silu, vpp);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Two-launch forward with the statistics finished by the statistics kernel (see gn_stats_v2_kernel).
// counters: PERSISTENT caller-owned buffer of >= n * 4 bytes, zero when first used.
extern "C" int leco_group_norm_v2(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw,
                                  int C, int G, float eps, int silu, void* workspace, void* counters, void* stream) {
  LECO_REQUIRE(x && y && stats && gamma && beta && workspace && counters, "leco_group_norm_v2: null pointer");
  LECO_REQUIRE(C % 8 == 0 && G > 0 && G <= GN_MAX_GROUPS && C % G == 0, "leco_group_norm_v2: C=%d G=%d unsupported", C, G);
  int vpp, threads, splits;
  LECO_REQUIRE(gn_launch_cfg(hw, C, &vpp, &threads, &splits) == 0, "leco_group_norm_v2: C=%d too wide", C);
  const int R = threads / vpp;
  count_launch();
  LECO_LAUNCH(gn_stats_v2_kernel, dim3(n, splits), threads, (size_t)2 * R * C * sizeof(float), STREAM(stream), BF(x),
              reinterpret_cast<float2*>(workspace), reinterpret_cast<float2*>(stats),
              reinterpret_cast<unsigned int*>(counters), hw, C, G, vpp, splits, eps);
  LECO_CHECK_CUDA(cudaGetLastError());
  static const int env_bpsm = [] { const char* e = getenv("LECO_GN_APPLY_BPSM"); return e ? atoi(e) : 8; }();
  static const int env_rpt = [] { const char* e = getenv("LECO_GN_APPLY_RPT"); return e ? atoi(e) : 4; }();
  int gy = (env_bpsm * 148) / (n < 1 ? 1 : n);
  if (gy > hw / (env_rpt * R)) gy = hw / (env_rpt * R);
  if (gy < 1) gy = 1;
  count_launch();
  LECO_LAUNCH(gn_apply_v2_kernel, dim3(n, gy), threads, 0, STREAM(stream), BF(x), BFW(y),
              reinterpret_cast<const float2*>(stats), BF(gamma), BF(beta), hw, C, G, silu, vpp);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// One-launch cluster forward where a sample's rows fit the shared memory of one cluster (every GroupNorm of the SD1.x /
// SD2.x / SDXL UNets at <= 64 samples' worth of 32x32 latents, and the 320-channel ones at 64x64), the two-launch v2 pair
// otherwise.  LECO_GN_CLUSTER=0 disables the cluster path, =8 uses clusters of 8 (portable size).
namespace leco {
// LECO_GN_CLUSTER: 0 = never use the cluster kernel, 8 / 16 = only that cluster size, unset = choose per launch
static int gnc_cluster_size_cfg() {
  static const int v = [] {
    const char* e = getenv("LECO_GN_CLUSTER");
    const int c = e ? atoi(e) : -1;
    return (c == 0 || c == 8 || c == 16) ? c : -1;
  }();
  return v;
}
constexpr size_t GNC_MAX_DYN_SMEM = 216 * 1024;      // + ~9.5 KB static: under the 227 KB per-CTA limit
// How many clusters of `cl` CTAs with the largest footprint can be resident at once (ncu: 7 of 16 on a B200, some GPCs
// expose fewer than 16 SMs).  A launch with more clusters than that runs in two waves and loses to the two-launch path.
static int gnc_max_active(int cl) {
  static int cache[2] = {-1, -1};
  int& slot = cache[cl == 16 ? 1 : 0];
  if (slot >= 0) return slot;
  slot = 0;
  static bool attrs = false, attrs_ok = false;
  if (!attrs) {
    attrs = true;
    attrs_ok = cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GNC_MAX_DYN_SMEM) == cudaSuccess &&
               // same shared-memory carve-out as the GEMMs around it, whatever this launch needs: no SM reconfiguration
               cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
               cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  }
  if (attrs_ok) {
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(cl, 1, 1);
    q.blockDim = dim3(GNC_THREADS, 1, 1);
    q.dynamicSmemBytes = GNC_MAX_DYN_SMEM;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = cl;
    a[0].val.clusterDim.y = 1;
    a[0].val.clusterDim.z = 1;
    q.attrs = a;
    q.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, gn_cluster_kernel, &q) == cudaSuccess && nclusters >= 1) slot = nclusters;
  }
  (void)cudaGetLastError();
  return slot;
}
// 0 = launched; 1 = shape not eligible (caller falls back); <0 = error
static int launch_gn_cluster(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw,
                             int C, int G, float eps, int silu, cudaStream_t stream) {
  const int forced = gnc_cluster_size_cfg();
  if (forced == 0 || n > 65535) return 1;
  // channel parts (clusters per sample): measured slower than one cluster per sample at every UNet shape (8 clusters
  // do not fit the 7 resident ones, and the row segments get short), so only on request (LECO_GN_PARTS = 2 / 4)
  static const int max_parts = [] { const char* e = getenv("LECO_GN_PARTS"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
  int CL = 0, parts = 0, vpp = 0, R = 0, rows_per = 0;
  size_t smem = 0;
  for (int cl = 16; cl >= 8 && !parts; cl >>= 1) {
    if (forced > 0 && cl != forced) continue;
    const int resident = gnc_max_active(cl);
    const int rp = (hw + cl - 1) / cl;
    for (int pc = 1; pc <= 4 && pc <= max_parts; pc *= 2) {
      if (G % pc || (C / pc) % 8 || (C / pc) % (G / pc)) break;
      const int cs = C / pc, vp = cs / 8;
      if (vp > GNC_THREADS) continue;
      int r = GNC_THREADS / vp;
      if (r < 1) r = 1;
      if (vp * r < 32) break;
      const size_t sm = (size_t)2 * r * cs * sizeof(float) + (size_t)rp * cs * 2;
      if (sm > GNC_MAX_DYN_SMEM) continue;
      if ((long long)n * pc > resident) break;          // would need a second wave of clusters
      CL = cl; parts = pc; vpp = vp; R = r; smem = sm; rows_per = rp;
      break;
    }
  }
  if (!parts) return 1;
  const int threads = vpp * R;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, parts, n);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  count_launch();
  LECO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gn_cluster_kernel, BF(x), BFW(y), reinterpret_cast<float2*>(stats), BF(gamma),
                                     BF(beta), hw, C, C / parts, G, G / parts, eps, silu, vpp, rows_per));
  return 0;
}
}  // namespace leco

extern "C" int leco_group_norm_v3(const void* x, void* y, void* stats, const void* gamma, const void* beta, int n, int hw,
                                  int C, int G, float eps, int silu, void* workspace, void* counters, void* stream) {
  LECO_REQUIRE(x && y && stats && gamma && beta && workspace && counters, "leco_group_norm_v3: null pointer");
  LECO_REQUIRE(C % 8 == 0 && G > 0 && G <= GN_MAX_GROUPS && C % G == 0, "leco_group_norm_v3: C=%d G=%d unsupported", C, G);
  const int rc = launch_gn_cluster(x, y, stats, gamma, beta, n, hw, C, G, eps, silu, STREAM(stream));
  if (rc <= 0) return rc;
  return leco_group_norm_v2(x, y, stats, gamma, beta, n, hw, C, G, eps, silu, workspace, counters, stream);
}

extern "C" int leco_group_norm_bwd(const void* x, const void* dz, void* dx, const void* stats, const void* gamma,
                                   const void* beta, int n, int hw, int C, int G, int silu, void* workspace,
                                   void* stream) {
  LECO_REQUIRE(x && dz && dx && stats && gamma && beta && workspace, "leco_group_norm_bwd: null pointer");
  LECO_REQUIRE(C % 8 == 0 && G > 0 && G <= GN_MAX_GROUPS && C % G == 0, "leco_group_norm_bwd: C=%d G=%d", C, G);
  int vpp, threads, splits;
  LECO_REQUIRE(gn_launch_cfg(hw, C, &vpp, &threads, &splits) == 0, "leco_group_norm_bwd: C=%d too wide", C);
  count_launch();
  LECO_LAUNCH(gn_bwd_stats_kernel, dim3(n, splits), threads, (size_t)2 * (threads / vpp) * C * sizeof(float), STREAM(stream), 
      BF(x), BF(dz), reinterpret_cast<const float2*>(stats), BF(gamma), BF(beta),
      reinterpret_cast<float2*>(workspace), hw, C, G, vpp, splits, silu);
  LECO_CHECK_CUDA(cudaGetLastError());
  long long work = (long long)hw * vpp;
  int gy = (int)((work + 256 * 4 - 1) / (256 * 4));
  if (gy > 296 / (n < 1 ? 1 : n) + 1) gy = 296 / (n < 1 ? 1 : n) + 1;
  if (gy < 1) gy = 1;
  count_launch();
  LECO_LAUNCH(gn_bwd_apply_kernel, dim3(n, gy), 256, 0, STREAM(stream), 
      BF(x), BF(dz), BFW(dx), reinterpret_cast<const float2*>(stats), reinterpret_cast<const float2*>(workspace),
      BF(gamma), BF(beta), hw, C, G, splits, silu);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_layer_norm(const void* x, void* y, void* stats /*float2[M] or NULL*/, const void* gamma,
                               const void* beta, int64_t M, int C, float eps, void* stream) {
  LECO_REQUIRE(x && y && gamma && beta, "leco_layer_norm: null pointer");
  LECO_REQUIRE(C % 8 == 0 && C <= 8 * 32 * LN_MAX_VEC, "leco_layer_norm: C=%d unsupported", C);
  const int vpr = C / 8;
  // lanes per row: the smallest of 8 / 16 / 32 that keeps a lane at <= 5 vectors (else 32 lanes with up to 8)
  int G = 8;
  while (G < 32 && (vpr + G - 1) / G > 5) G *= 2;
  const int vpl = (vpr + G - 1) / G;
  const int rpw = 32 / G;
  long long warps = (M + rpw - 1) / rpw;
  long long blocks = (warps + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  count_launch();
  if (vpl <= 5)
    LECO_LAUNCH((ln_fwd_kernel<5>), (int)blocks, 256, 0, STREAM(stream), BF(x), BFW(y), reinterpret_cast<float2*>(stats),
                BF(gamma), BF(beta), (long long)M, C, eps, G);
  else
    LECO_LAUNCH((ln_fwd_kernel<LN_MAX_VEC>), (int)blocks, 256, 0, STREAM(stream), BF(x), BFW(y),
                reinterpret_cast<float2*>(stats), BF(gamma), BF(beta), (long long)M, C, eps, G);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_layer_norm_bwd(const void* x, const void* dy, void* dx, const void* stats, const void* gamma,
                                   int64_t M, int C, void* stream) {
  LECO_REQUIRE(x && dy && dx && stats && gamma, "leco_layer_norm_bwd: null pointer");
  LECO_REQUIRE(C % 8 == 0 && C <= 8 * 32 * LN_MAX_VEC, "leco_layer_norm_bwd: C=%d unsupported", C);
  long long blocks = (M + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  count_launch();
  LECO_LAUNCH(ln_bwd_kernel, (int)blocks, 256, 0, STREAM(stream), BF(x), BF(dy), BFW(dx), reinterpret_cast<const float2*>(stats),
                                                       BF(gamma), M, C);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
