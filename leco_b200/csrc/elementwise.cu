// HBM-bound glue kernels of the UNet step: boundary convolutions (conv_in K=36, conv_out N=4),
// timestep sinusoid, SiLU, GEGLU (fwd/bwd), channel concat/split, nearest 2x upsample (fwd/bwd),
// stride-2 im2col/col2im, batched transpose, row softmax (fwd/bwd) and axpy.  All are
// coalesced 16-byte-vector kernels; none of them is worth a tensor core.
#include "../../include/leco_b200.h"
#include <stdlib.h>

#include "common.cuh"

namespace leco {
void count_launch();

static inline int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct alignas(16) bf16x8 {
  uint32_t u[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& q, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(q.u[i]);
    f[2 * i + 1] = bf16_hi(q.u[i]);
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 q;
#pragma unroll
  for (int i = 0; i < 4; ++i) q.u[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
  return q;
}

// ------------------------------------------------------------------ conv_in (3x3, Cin=4)
// x: NCHW [n,4,h,w] (fp32 or bf16) -> y: NHWC [n*h*w, cout] bf16.  w: OIHW [cout,4,3,3] bf16.
template <typename TIn>
__global__ void conv_in_kernel(const TIn* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                               const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                               int n, int h, int wd, int cout) {
  pdl_entry();
  extern __shared__ float s_w[];  // [36][cout] (k-major so a thread's 8 channels are contiguous) + bias[cout]
  float* s_b = s_w + 36 * cout;
  for (int i = threadIdx.x; i < 36 * cout; i += blockDim.x) {
    const int o = i / 36, k = i - o * 36;
    s_w[k * cout + o] = __bfloat162float(w[i]);
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? __bfloat162float(bias[i]) : 0.f;
  __syncthreads();
  const int groups = cout / 8;
  const long long total = 1LL * n * h * wd * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const int px = (int)(pix % wd);
    const int py = (int)((pix / wd) % h);
    const int img = (int)(pix / ((long long)wd * h));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = s_b[g * 8 + j];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py + ky - 1, xx = px + kx - 1;
          if (yy >= 0 && yy < h && xx >= 0 && xx < wd) {
            const float v = static_cast<float>(x[((1LL * img * 4 + c) * h + yy) * wd + xx]);
            const float4* wr = reinterpret_cast<const float4*>(s_w + (c * 9 + ky * 3 + kx) * cout + g * 8);
            const float4 w0 = wr[0], w1 = wr[1];
            acc[0] = fmaf(v, w0.x, acc[0]);
            acc[1] = fmaf(v, w0.y, acc[1]);
            acc[2] = fmaf(v, w0.z, acc[2]);
            acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]);
            acc[5] = fmaf(v, w1.y, acc[5]);
            acc[6] = fmaf(v, w1.z, acc[6]);
            acc[7] = fmaf(v, w1.w, acc[7]);
          }
        }
    *reinterpret_cast<bf16x8*>(y + pix * cout + g * 8) = pack8(acc);
  }
}

// Four consecutive pixels of a row per thread (w % 4 == 0): every 8-channel weight slice read from shared memory feeds
// 4 x 8 FMAs instead of 8, and the 3 x 6 input patch of a channel is loaded once for the four outputs.  The one-pixel
// kernel above was bound by its shared-memory weight reads (72 LDS.128 per 288 FMAs): 74 us per UNet forward at
// 4 x 64 x 64 -> 320, 1 % of a denoise step (in-graph timeline).  Same accumulation order per output: same bits.
template <typename TIn>
__global__ void __launch_bounds__(256, 2)
conv_in4_kernel(const TIn* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
                __nv_bfloat16* __restrict__ y, int n, int h, int wd, int cout) {
  pdl_entry();
  extern __shared__ float s_w[];  // [36][cout] + bias[cout]
  float* s_b = s_w + 36 * cout;
  for (int i = threadIdx.x; i < 36 * cout; i += blockDim.x) {   // consecutive threads -> consecutive banks
    const int k = i / cout, o = i - k * cout;
    s_w[i] = __bfloat162float(w[o * 36 + k]);
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? __bfloat162float(bias[i]) : 0.f;
  __syncthreads();
  const int groups = cout / 8, qpr = wd / 4;
  const long long total = 1LL * n * h * qpr * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long q = i / groups;
    const int px0 = (int)(q % qpr) * 4;
    const int py = (int)((q / qpr) % h);
    const int img = (int)(q / ((long long)qpr * h));
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = s_b[g * 8 + j];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = py + ky - 1;
        float xv[6];
        const TIn* row = x + ((1LL * img * 4 + c) * h + yy) * wd;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const int xx = px0 - 1 + t;
          xv[t] = (yy >= 0 && yy < h && xx >= 0 && xx < wd) ? static_cast<float>(row[xx]) : 0.f;   // zero padding
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + (c * 9 + ky * 3 + kx) * cout + g * 8);
          const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = xv[p + kx];
            acc[p][0] = fmaf(v, w0.x, acc[p][0]);
            acc[p][1] = fmaf(v, w0.y, acc[p][1]);
            acc[p][2] = fmaf(v, w0.z, acc[p][2]);
            acc[p][3] = fmaf(v, w0.w, acc[p][3]);
            acc[p][4] = fmaf(v, w1.x, acc[p][4]);
            acc[p][5] = fmaf(v, w1.y, acc[p][5]);
            acc[p][6] = fmaf(v, w1.z, acc[p][6]);
            acc[p][7] = fmaf(v, w1.w, acc[p][7]);
          }
        }
      }
    const long long pix0 = ((long long)img * h + py) * wd + px0;
#pragma unroll
    for (int p = 0; p < 4; ++p) *reinterpret_cast<bf16x8*>(y + (pix0 + p) * cout + g * 8) = pack8(acc[p]);
  }
}

// ------------------------------------------------------------------ conv_out (3x3, Cout<=8)
// x: NHWC [n*h*w, c] bf16 -> y: NCHW [n,cout,h,w] fp32.  w: [cout][9][c] bf16 (OHWI).
__global__ void conv_out_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                const __nv_bfloat16* __restrict__ bias, float* __restrict__ y, int n,
                                int h, int wd, int c, int cout) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long npix = 1LL * n * h * wd;
  const int vpp = c / 8;
  for (long long pix = warp; pix < npix; pix += nwarps) {
    const int px = (int)(pix % wd);
    const int py = (int)((pix / wd) % h);
    const int img = (int)(pix / ((long long)wd * h));
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      if (yy < 0 || yy >= h || xx < 0 || xx >= wd) continue;
      const __nv_bfloat16* xr = x + ((1LL * img * h + yy) * wd + xx) * c;
      for (int v = lane; v < vpp; v += 32) {
        float xv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(xr + v * 8), xv);
        for (int o = 0; o < cout; ++o) {
          float wv[8];
          unpack8(*reinterpret_cast<const bf16x8*>(w + ((size_t)o * 9 + tap) * c + v * 8), wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[o] = fmaf(xv[j], wv[j], acc[o]);
        }
      }
    }
    for (int o = 0; o < cout; ++o) {
      float a = acc[o];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      if (lane == 0)
        y[((1LL * img * cout + o) * h + py) * wd + px] = a + (bias ? __bfloat162float(bias[o]) : 0.f);
    }
  }
}

// [M = n*hw][ld] fp32 GEMM result (first cout columns) + bias -> NCHW fp32 [n][cout][hw]: the tail of conv_out when it
// runs as a tcgen05 implicit-GEMM conv with N padded to 8
__global__ void cols_to_nchw_kernel(const float* __restrict__ y8, int ld, const __nv_bfloat16* __restrict__ bias,
                                    float* __restrict__ out, int n, int hw, int cout) {
  pdl_entry();
  const long long total = 1LL * n * hw;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(pix / hw);
    const int p = (int)(pix - (long long)img * hw);
    const float4 a = *reinterpret_cast<const float4*>(y8 + pix * ld);
    const float4 b = *reinterpret_cast<const float4*>(y8 + pix * ld + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < cout) out[((long long)img * cout + o) * hw + p] = v[o] + (bias ? __bfloat162float(bias[o]) : 0.f);
  }
}

// conv_out backward-data: dy NCHW fp32 [n,cout,h,w] -> dx NHWC bf16 [n*h*w, c]
// dx[p, ci] = sum_{o,tap} dy[o, p - off(tap)] * w[o][tap][ci]
__global__ void conv_out_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ w,
                                    __nv_bfloat16* __restrict__ dx, int n, int h, int wd, int c, int cout) {
  pdl_entry();
  const int vpp = c / 8;
  const long long total = 1LL * n * h * wd * vpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    const long long pix = i / vpp;
    const int px = (int)(pix % wd);
    const int py = (int)((pix / wd) % h);
    const int img = (int)(pix / ((long long)wd * h));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      // output pixel q that used input p with this tap: q = p - (tap offset)
      const int yy = py - (tap / 3 - 1), xx = px - (tap % 3 - 1);
      if (yy < 0 || yy >= h || xx < 0 || xx >= wd) continue;
      for (int o = 0; o < cout; ++o) {
        const float g = dy[((1LL * img * cout + o) * h + yy) * wd + xx];
        float wv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(w + ((size_t)o * 9 + tap) * c + v * 8), wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g, wv[j], acc[j]);
      }
    }
    *reinterpret_cast<bf16x8*>(dx + pix * c + v * 8) = pack8(acc);
  }
}

// ------------------------------------------------------------------ timestep sinusoid
// diffusers get_timestep_embedding(flip_sin_to_cos=True, freq_shift=0): [cos | sin]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out,
                                          int n, int dim) {
  pdl_entry();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int j = i % half, r = i / half;
  const float freq = expf(-9.210340371976184f * (float)j / (float)half);
  const float a = t[r] * freq;
  out[(size_t)r * dim + j] = __float2bfloat16(cosf(a));
  out[(size_t)r * dim + half + j] = __float2bfloat16(sinf(a));
}

// ------------------------------------------------------------------ elementwise
__global__ void silu_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ y, long long nvec) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(x[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
    y[i] = pack8(f);
  }
}
// text-encoder MLP activations (transformers CLIPMLP, modeling_clip.py: ACT2FN[config.hidden_act]):
// kind 1 = quick_gelu x*sigmoid(1.702x) (CLIP ViT-L), kind 2 = erf GELU (OpenCLIP ViT-H / bigG), kind 0 = SiLU
__global__ void activation_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ y, long long nvec, int kind) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(x[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      f[j] = kind == 1 ? f[j] / (1.0f + __expf(-1.702f * f[j])) : (kind == 2 ? gelu_erf_f(f[j]) : silu_f(f[j]));
    y[i] = pack8(f);
  }
}
// CLIPTextEmbeddings (modeling_clip.py): out[r] = token_embedding[ids[r]] + position_embedding[r % seq], fp32 add,
// one bf16 rounding.  8 columns per thread.
__global__ void embed_tokens_kernel(const int* __restrict__ ids, const bf16x8* __restrict__ tok,
                                    const bf16x8* __restrict__ pos, bf16x8* __restrict__ out, long long rows, int seq,
                                    int dvec, int vocab) {
  pdl_entry();
  const long long total = rows * dvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dvec;
    const int c = static_cast<int>(i - r * dvec);
    int id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    float a[8], b[8];
    unpack8(tok[(long long)id * dvec + c], a);
    unpack8(pos[(r % seq) * dvec + c], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    out[i] = pack8(a);
  }
}
// y += x   (gradient accumulation)
__global__ void add_inplace_kernel(bf16x8* __restrict__ y, const bf16x8* __restrict__ x, long long nvec) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(y[i], a);
    unpack8(x[i], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    y[i] = pack8(a);
  }
}
// GEGLU: pre [M, 2H] = [hidden | gate] -> out [M,H] = hidden * gelu(gate)
__global__ void geglu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ out,
                                 long long M, int H) {
  pdl_entry();
  const int vpr = H / 8;
  const long long total = M * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int v = (int)(i % vpr);
    float h[8], g[8];
    unpack8(*reinterpret_cast<const bf16x8*>(pre + r * 2 * H + v * 8), h);
    unpack8(*reinterpret_cast<const bf16x8*>(pre + r * 2 * H + H + v * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] *= gelu_erf_f(g[j]);
    *reinterpret_cast<bf16x8*>(out + r * H + v * 8) = pack8(h);
  }
}
__global__ void geglu_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dout,
                                 __nv_bfloat16* __restrict__ dpre, long long M, int H) {
  pdl_entry();
  const int vpr = H / 8;
  const long long total = M * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int v = (int)(i % vpr);
    float h[8], g[8], d[8], dh[8], dg[8];
    unpack8(*reinterpret_cast<const bf16x8*>(pre + r * 2 * H + v * 8), h);
    unpack8(*reinterpret_cast<const bf16x8*>(pre + r * 2 * H + H + v * 8), g);
    unpack8(*reinterpret_cast<const bf16x8*>(dout + r * H + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cdf = 0.5f * (1.0f + erff(g[j] * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * g[j] * g[j]);
      dh[j] = d[j] * g[j] * cdf;
      dg[j] = d[j] * h[j] * (cdf + g[j] * pdf);
    }
    *reinterpret_cast<bf16x8*>(dpre + r * 2 * H + v * 8) = pack8(dh);
    *reinterpret_cast<bf16x8*>(dpre + r * 2 * H + H + v * 8) = pack8(dg);
  }
}

// generic strided row copy: dst[r, dcol0 + j] = src[r, scol0 + j], j < ncols (all multiples of 8)
__global__ void copy_cols_kernel(const __nv_bfloat16* __restrict__ src, long long lds, int scol0,
                                 __nv_bfloat16* __restrict__ dst, long long ldd, int dcol0, long long M,
                                 int ncols) {
  pdl_entry();
  const int vpr = ncols / 8;
  const long long total = M * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int v = (int)(i % vpr);
    *reinterpret_cast<bf16x8*>(dst + r * ldd + dcol0 + v * 8) =
        *reinterpret_cast<const bf16x8*>(src + r * lds + scol0 + v * 8);
  }
}

// nearest 2x upsample NHWC: y[n, 2h, 2w, c]
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n,
                                  int h, int w, int c) {
  pdl_entry();
  const int vpp = c / 8;
  const long long total = 1LL * n * 2 * h * 2 * w * vpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    long long p = i / vpp;
    const int ox = (int)(p % (2 * w));
    p /= 2 * w;
    const int oy = (int)(p % (2 * h));
    const int img = (int)(p / (2 * h));
    const long long src = ((1LL * img * h + oy / 2) * w + ox / 2) * c + v * 8;
    *reinterpret_cast<bf16x8*>(y + (i / vpp) * c + v * 8) = *reinterpret_cast<const bf16x8*>(x + src);
  }
}
// backward: dx[n,h,w,c] = sum of the 2x2 block of dy[n,2h,2w,c]
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                      int n, int h, int w, int c) {
  pdl_entry();
  const int vpp = c / 8;
  const long long total = 1LL * n * h * w * vpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    long long p = i / vpp;
    const int ix = (int)(p % w);
    p /= w;
    const int iy = (int)(p % h);
    const int img = (int)(p / h);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(
                    dy + ((1LL * img * 2 * h + 2 * iy + dyy) * 2 * w + 2 * ix + dxx) * c + v * 8),
                f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    *reinterpret_cast<bf16x8*>(dx + (i / vpp) * c + v * 8) = pack8(acc);
  }
}

// stride-2 3x3 pad-1 im2col: x NHWC [n,h,w,c] -> col [n*(h/2)*(w/2), 9*c] (k = tap*c + ch)
__global__ void im2col_s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, int n,
                                 int h, int w, int c, int stride) {
  pdl_entry();
  const int vpp = c / 8, oh = h / stride, ow = w / stride;
  const long long total = 1LL * n * oh * ow * 9 * vpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    long long p = i / vpp;
    const int tap = (int)(p % 9);
    p /= 9;
    const int ox = (int)(p % ow);
    const int oy = (int)((p / ow) % oh);
    const int img = (int)(p / ((long long)ow * oh));
    const int yy = stride * oy + tap / 3 - 1, xx = stride * ox + tap % 3 - 1;
    bf16x8 q = {{0, 0, 0, 0}};
    if (yy >= 0 && yy < h && xx >= 0 && xx < w)
      q = *reinterpret_cast<const bf16x8*>(x + ((1LL * img * h + yy) * w + xx) * c + v * 8);
    *reinterpret_cast<bf16x8*>(col + p * 9 * c + (size_t)tap * c + v * 8) = q;
  }
}
// col2im (gather form): dx[n,h,w,c] = sum over (out pixel, tap) that read this input pixel
__global__ void col2im_s2_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int n,
                                 int h, int w, int c) {
  pdl_entry();
  const int vpp = c / 8, oh = h / 2, ow = w / 2;
  const long long total = 1LL * n * h * w * vpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    long long p = i / vpp;
    const int ix = (int)(p % w);
    const int iy = (int)((p / w) % h);
    const int img = (int)(p / ((long long)w * h));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int tap = 0; tap < 9; ++tap) {
      const int ny = iy - (tap / 3 - 1), nx = ix - (tap % 3 - 1);  // = 2*oy, 2*ox
      if (ny < 0 || nx < 0 || (ny & 1) || (nx & 1)) continue;
      const int oy = ny >> 1, ox = nx >> 1;
      if (oy >= oh || ox >= ow) continue;
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dcol + ((1LL * img * oh + oy) * ow + ox) * 9 * c +
                                               (size_t)tap * c + v * 8),
              f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    *reinterpret_cast<bf16x8*>(dx + p * c + v * 8) = pack8(acc);
  }
}

// out[n, c] = sum over the hw rows of sample n (the time-embedding bias gradient of a conv epilogue)
__global__ void rowgroup_sum_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int hw,
                                    int c) {
  pdl_entry();
  __shared__ float red[8][32][8];
  const int n = blockIdx.x;
  const int v = blockIdx.y * 32 + threadIdx.x;  // 8-channel vector column
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (v * 8 < c) {
    for (int r = threadIdx.y; r < hw; r += 8) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + ((size_t)n * hw + r) * c + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && v * 8 < c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][j];
      acc[j] = s;
    }
    *reinterpret_cast<bf16x8*>(out + (size_t)n * c + v * 8) = pack8(acc);
  }
}

// batched 2-D transpose with arbitrary strides: out[b1,b0,c,r] = in[b1,b0,r,c]
// (zero-fills out columns r in [rows, rows_pad))
__global__ void transpose_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                 int rows, int cols, int rows_pad, long long in_ld, long long in_bs0,
                                 long long in_bs1, long long out_ld, long long out_bs0, long long out_bs1,
                                 int batch0) {
  pdl_entry();
  __shared__ __nv_bfloat16 tile[32][34];
  const int b = blockIdx.z;
  const int b1 = b / batch0, b0 = b % batch0;
  const __nv_bfloat16* src = in + b0 * in_bs0 + b1 * in_bs1;
  __nv_bfloat16* dst = out + b0 * out_bs0 + b1 * out_bs1;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * in_ld + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows_pad) dst[(long long)c * out_ld + r] = tile[threadIdx.x][i];
  }
}

// row softmax: s fp32 [rows, ld_s] (first n_valid columns valid) -> p bf16 [rows, ld_p], columns
// [n_valid, n_pad) written as 0.  One warp per row.
// causal_sq > 0: row r only sees columns <= r % causal_sq (the text encoder's causal mask, CLIPTextTransformer).
__global__ void softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, long long rows,
                                    int n_valid_all, int n_pad, long long ld_s, long long ld_p, int causal_sq) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* sr = s + r * ld_s;
    int n_valid = n_valid_all;
    if (causal_sq > 0) n_valid = min(n_valid_all, static_cast<int>(r % causal_sq) + 1);
    float mx = -INFINITY;
    for (int j = lane; j < n_valid; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < n_valid; j += 32) sum += __expf(sr[j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    __nv_bfloat16* pr = p + r * ld_p;
    for (int j = lane; j < n_pad; j += 32)
      pr[j] = __float2bfloat16(j < n_valid ? __expf(sr[j] - mx) * inv : 0.f);
  }
}
// softmax backward: ds = p * (dp - sum_j p_j dp_j) * scale ; dp fp32 [rows, ld], p bf16, ds bf16
__global__ void softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                                        __nv_bfloat16* __restrict__ ds, long long rows, int n_valid, int n_pad,
                                        long long ld_p, long long ld_dp, float scale) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const __nv_bfloat16* pr = p + r * ld_p;
    const float* dr = dp + r * ld_dp;
    float dot = 0.f;
    for (int j = lane; j < n_valid; j += 32) dot += __bfloat162float(pr[j]) * dr[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    __nv_bfloat16* dsr = ds + r * ld_p;
    for (int j = lane; j < n_pad; j += 32)
      dsr[j] = __float2bfloat16(j < n_valid ? __bfloat162float(pr[j]) * (dr[j] - dot) * scale : 0.f);
  }
}

}  // namespace leco

using namespace leco;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

extern "C" int leco_conv_in(const void* x, int x_is_fp32, const void* w, const void* bias, void* y, int n, int h,
                            int wd, int cout, void* stream) {
  LECO_REQUIRE(x && w && y && cout % 8 == 0, "leco_conv_in: bad args");
  const long long work = 1LL * n * h * wd * (cout / 8);
  const size_t smem = (size_t)37 * cout * sizeof(float);
  LECO_REQUIRE(smem <= 48 * 1024, "leco_conv_in: cout=%d too wide", cout);
  int grid = grid_for(work, 256);
  if (grid > 148 * 4) grid = 148 * 4;  // each block re-stages the weights: keep the grid modest
  count_launch();
  static const bool quad = [] { const char* e = getenv("LECO_CONV_IN_QUAD"); return !(e && e[0] == '0'); }();
  if (quad && wd % 4 == 0) {           // four pixels per thread (LECO_CONV_IN_QUAD=0: the one-pixel kernel, for A/B runs)
    int grid4 = grid_for(work / 4, 256);
    if (grid4 > 148 * 2) grid4 = 148 * 2;
    if (x_is_fp32)
      LECO_LAUNCH(conv_in4_kernel<float>, grid4, 256, smem, STREAM(stream), reinterpret_cast<const float*>(x), BF(w), BF(bias),
                  BFW(y), n, h, wd, cout);
    else
      LECO_LAUNCH(conv_in4_kernel<__nv_bfloat16>, grid4, 256, smem, STREAM(stream), BF(x), BF(w), BF(bias), BFW(y), n, h, wd,
                  cout);
    LECO_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (x_is_fp32)
    LECO_LAUNCH(conv_in_kernel<float>, grid, 256, smem, STREAM(stream), reinterpret_cast<const float*>(x), BF(w), BF(bias),
                                                              BFW(y), n, h, wd, cout);
  else
    LECO_LAUNCH(conv_in_kernel<__nv_bfloat16>, grid, 256, smem, STREAM(stream), BF(x), BF(w), BF(bias), BFW(y), n, h, wd,
                                                                      cout);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_conv_out(const void* x, const void* w, const void* bias, float* y, int n, int h, int wd, int c,
                             int cout, void* stream) {
  LECO_REQUIRE(x && w && y && c % 8 == 0 && cout >= 1 && cout <= 8, "leco_conv_out: bad args");
  count_launch();
  LECO_LAUNCH(conv_out_kernel, grid_for(1LL * n * h * wd * 32, 256), 256, 0, STREAM(stream), BF(x), BF(w), BF(bias), y, n,
                                                                                   h, wd, c, cout);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_cols_to_nchw(const float* y8, int ld, const void* bias, float* out, int n, int hw, int cout,
                                 void* stream) {
  LECO_REQUIRE(y8 && out && ld >= 8 && ld % 4 == 0 && cout >= 1 && cout <= 8, "leco_cols_to_nchw: bad args");
  count_launch();
  LECO_LAUNCH(cols_to_nchw_kernel, grid_for(1LL * n * hw, 256), 256, 0, STREAM(stream), y8, ld, BF(bias), out, n, hw, cout);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_conv_out_bwd(const float* dy, const void* w, void* dx, int n, int h, int wd, int c, int cout,
                                 void* stream) {
  LECO_REQUIRE(dy && w && dx && c % 8 == 0 && cout >= 1 && cout <= 8, "leco_conv_out_bwd: bad args");
  count_launch();
  LECO_LAUNCH(conv_out_bwd_kernel, grid_for(1LL * n * h * wd * (c / 8), 256), 256, 0, STREAM(stream), dy, BF(w), BFW(dx),
                                                                                           n, h, wd, c, cout);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_timestep_embedding(const float* t, void* out, int n, int dim, void* stream) {
  LECO_REQUIRE(t && out && dim % 2 == 0, "leco_timestep_embedding: bad args");
  count_launch();
  LECO_LAUNCH(timestep_embedding_kernel, (n * dim / 2 + 127) / 128, 128, 0, STREAM(stream), t, BFW(out), n, dim);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_silu(const void* x, void* y, int64_t numel, void* stream) {
  LECO_REQUIRE(x && y && numel % 8 == 0, "leco_silu: numel must be a multiple of 8");
  count_launch();
  LECO_LAUNCH(silu_kernel, grid_for(numel / 8, 256), 256, 0, STREAM(stream), reinterpret_cast<const bf16x8*>(x),
                                                                   reinterpret_cast<bf16x8*>(y), numel / 8);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_activation(const void* x, void* y, int64_t numel, int kind, void* stream) {
  LECO_REQUIRE(x && y && numel % 8 == 0 && kind >= 0 && kind <= 2, "leco_activation: numel % 8 != 0 or unknown kind");
  count_launch();
  LECO_LAUNCH(activation_kernel, grid_for(numel / 8, 256), 256, 0, STREAM(stream), reinterpret_cast<const bf16x8*>(x),
                                                                         reinterpret_cast<bf16x8*>(y), numel / 8, kind);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_embed_tokens(const int* ids, const void* tok, const void* pos, void* out, int64_t rows, int seq,
                                 int dim, int vocab, void* stream) {
  LECO_REQUIRE(ids && tok && pos && out && rows > 0 && seq > 0 && dim % 8 == 0 && vocab > 0, "leco_embed_tokens: bad args");
  count_launch();
  LECO_LAUNCH(embed_tokens_kernel, grid_for(rows * (dim / 8), 256), 256, 0, STREAM(stream), ids,
                                                                           reinterpret_cast<const bf16x8*>(tok),
                                                                           reinterpret_cast<const bf16x8*>(pos),
                                                                           reinterpret_cast<bf16x8*>(out), rows, seq, dim / 8, vocab);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_add_inplace(void* y, const void* x, int64_t numel, void* stream) {
  LECO_REQUIRE(x && y && numel % 8 == 0, "leco_add_inplace: numel must be a multiple of 8");
  count_launch();
  LECO_LAUNCH(add_inplace_kernel, grid_for(numel / 8, 256), 256, 0, STREAM(stream), 
      reinterpret_cast<bf16x8*>(y), reinterpret_cast<const bf16x8*>(x), numel / 8);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_geglu_fwd(const void* pre, void* out, int64_t M, int H, void* stream) {
  LECO_REQUIRE(pre && out && H % 8 == 0, "leco_geglu_fwd: bad args");
  count_launch();
  LECO_LAUNCH(geglu_fwd_kernel, grid_for(M * (H / 8), 256), 256, 0, STREAM(stream), BF(pre), BFW(out), M, H);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_geglu_bwd(const void* pre, const void* dout, void* dpre, int64_t M, int H, void* stream) {
  LECO_REQUIRE(pre && dout && dpre && H % 8 == 0, "leco_geglu_bwd: bad args");
  count_launch();
  LECO_LAUNCH(geglu_bwd_kernel, grid_for(M * (H / 8), 256), 256, 0, STREAM(stream), BF(pre), BF(dout), BFW(dpre), M, H);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_copy_cols(const void* src, int64_t lds, int scol0, void* dst, int64_t ldd, int dcol0,
                              int64_t M, int ncols, void* stream) {
  LECO_REQUIRE(src && dst && ncols % 8 == 0 && scol0 % 8 == 0 && dcol0 % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0,
               "leco_copy_cols: columns / strides must be multiples of 8");
  count_launch();
  LECO_LAUNCH(copy_cols_kernel, grid_for(M * (ncols / 8), 256), 256, 0, STREAM(stream), BF(src), lds, scol0, BFW(dst),
                                                                               ldd, dcol0, M, ncols);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_upsample2x(const void* x, void* y, int n, int h, int w, int c, void* stream) {
  LECO_REQUIRE(x && y && c % 8 == 0, "leco_upsample2x: bad args");
  count_launch();
  LECO_LAUNCH(upsample2x_kernel, grid_for(4LL * n * h * w * (c / 8), 256), 256, 0, STREAM(stream), BF(x), BFW(y), n, h, w,
                                                                                        c);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_upsample2x_bwd(const void* dy, void* dx, int n, int h, int w, int c, void* stream) {
  LECO_REQUIRE(dy && dx && c % 8 == 0, "leco_upsample2x_bwd: bad args");
  count_launch();
  LECO_LAUNCH(upsample2x_bwd_kernel, grid_for(1LL * n * h * w * (c / 8), 256), 256, 0, STREAM(stream), BF(dy), BFW(dx), n,
                                                                                            h, w, c);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_im2col_s2(const void* x, void* col, int n, int h, int w, int c, void* stream) {
  LECO_REQUIRE(x && col && c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "leco_im2col_s2: bad args");
  count_launch();
  LECO_LAUNCH(im2col_s2_kernel, grid_for(1LL * n * (h / 2) * (w / 2) * 9 * (c / 8), 256), 256, 0, STREAM(stream), 
      BF(x), BFW(col), n, h, w, c, 2);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_im2col_s1(const void* x, void* col, int n, int h, int w, int c, void* stream) {
  LECO_REQUIRE(x && col && c % 8 == 0, "leco_im2col_s1: bad args");
  count_launch();
  LECO_LAUNCH(im2col_s2_kernel, grid_for(1LL * n * h * w * 9 * (c / 8), 256), 256, 0, STREAM(stream), BF(x), BFW(col), n, h, w,
                                                                                          c, 1);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_rowgroup_sum(const void* x, void* out, int n, int hw, int c, void* stream) {
  LECO_REQUIRE(x && out && c % 8 == 0, "leco_rowgroup_sum: bad args");
  count_launch();
  LECO_LAUNCH(rowgroup_sum_kernel, dim3(n, (c / 8 + 31) / 32), dim3(32, 8), 0, STREAM(stream), BF(x), BFW(out), hw, c);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_col2im_s2(const void* dcol, void* dx, int n, int h, int w, int c, void* stream) {
  LECO_REQUIRE(dcol && dx && c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "leco_col2im_s2: bad args");
  count_launch();
  LECO_LAUNCH(col2im_s2_kernel, grid_for(1LL * n * h * w * (c / 8), 256), 256, 0, STREAM(stream), BF(dcol), BFW(dx), n, h,
                                                                                       w, c);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_transpose(const void* in, void* out, int rows, int cols, int rows_pad, int64_t in_ld,
                              int64_t in_bs0, int64_t in_bs1, int64_t out_ld, int64_t out_bs0, int64_t out_bs1,
                              int batch0, int batch1, void* stream) {
  LECO_REQUIRE(in && out && rows > 0 && cols > 0 && rows_pad >= rows && batch0 > 0 && batch1 > 0,
               "leco_transpose: bad args");
  LECO_REQUIRE(1LL * batch0 * batch1 <= 65535, "leco_transpose: too many batches");
  dim3 grid((cols + 31) / 32, (rows_pad + 31) / 32, batch0 * batch1), block(32, 8);
  count_launch();
  LECO_LAUNCH(transpose_kernel, grid, block, 0, STREAM(stream), BF(in), BFW(out), rows, cols, rows_pad, in_ld, in_bs0,
                                                       in_bs1, out_ld, out_bs0, out_bs1, batch0);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_softmax_rows(const float* s, void* p, int64_t rows, int n_valid, int n_pad, int64_t ld_s,
                                 int64_t ld_p, void* stream) {
  LECO_REQUIRE(s && p && n_valid > 0 && n_pad >= n_valid, "leco_softmax_rows: bad args");
  count_launch();
  LECO_LAUNCH(softmax_rows_kernel, grid_for(rows * 32, 256), 256, 0, STREAM(stream), s, BFW(p), rows, n_valid, n_pad, ld_s,
                                                                           ld_p, 0);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_softmax_rows_causal(const float* s, void* p, int64_t rows, int n_valid, int n_pad, int64_t ld_s,
                                        int64_t ld_p, int sq, void* stream) {
  LECO_REQUIRE(s && p && n_valid > 0 && n_pad >= n_valid && sq > 0, "leco_softmax_rows_causal: bad args");
  count_launch();
  LECO_LAUNCH(softmax_rows_kernel, grid_for(rows * 32, 256), 256, 0, STREAM(stream), s, BFW(p), rows, n_valid, n_pad, ld_s,
                                                                           ld_p, sq);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int leco_softmax_bwd_rows(const void* p, const float* dp, void* ds, int64_t rows, int n_valid,
                                     int n_pad, int64_t ld_p, int64_t ld_dp, float scale, void* stream) {
  LECO_REQUIRE(p && dp && ds && n_valid > 0 && n_pad >= n_valid, "leco_softmax_bwd_rows: bad args");
  count_launch();
  LECO_LAUNCH(softmax_bwd_rows_kernel, grid_for(rows * 32, 256), 256, 0, STREAM(stream), BF(p), dp, BFW(ds), rows, n_valid,
                                                                               n_pad, ld_p, ld_dp, scale);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
