// Fused attention backward for sm_100a (head dim <= 64): dQ, dK, dV of O = softmax(scale Q K^T) V per (sample, head)
// without ever materialising S, P, dP or dS in HBM.  Replaces the autograd backward of xformers'
// memory_efficient_attention on the reference's one grad pass (train_lora.py:68, :279).
//
// One CTA per (128-key tile j, head, sample); it loops over the 128-query tiles i:
//   S   = Q_i K_j^T           tcgen05.mma M128 N128 K64 -> TMEM            (recomputed, never stored)
//   dP  = dO_i V_j^T          tcgen05.mma M128 N128 K64 -> TMEM
//   P   = exp2(S c - lse_i),  dS = scale P (dP - D_i)                      softmax warps, registers -> bf16 smem tiles
//   dV += P^T dO_i            tcgen05.mma M128(keys) N64 K128(queries), A and B MN-major (the smem tiles as stored)
//   dK += dS^T Q_i            same
//   dQ_i = dS K_j             tcgen05.mma M128(queries) N64 K128(keys) -> TMEM -> red.global.add.f32 into dq_acc
// lse (log2 domain) comes from the forward kernel (leco_flash_attn_fwd_lse), D_i = rowsum(dO_i o O_i) from
// leco_attn_bwd_prep.  dK / dV accumulate in TMEM over the whole query loop and are written once; dQ needs a sum
// over key tiles, i.e. over CTAs: fp32 reductions into a zeroed [batch][heads][sq][64] workspace, cast afterwards
// (leco_attn_dq_cast).  Roles as in the forward kernel: warp 0 lane 0 TMA, warp 1 lane 0 MMA issue, warp 2 TMEM
// allocator, warps 4-11 two softmax warpgroups (thread (wg,row) owns keys [64 wg, 64 wg + 64) of S row `row`).
#include "../../include/leco_b200.h"
#include "common.cuh"

namespace leco {
void count_launch();

constexpr int FB_M = 128, FB_D = 64;
constexpr int FB_TILE = FB_M * FB_D * 2;        // 16 KiB: one Q / dO / K / V tile
constexpr int FB_PS = FB_M * 128 * 2;           // 32 KiB: P or dS tile (two 64-key chunks of [128 rows x 128 B])
constexpr int FB_SMEM = 2 * FB_TILE /*K,V*/ + 4 * FB_TILE /*Q,dO x2*/ + 2 * FB_PS + 1024 + 256;
constexpr int FB_THREADS = 384;
constexpr uint32_t FB_TM_S = 0, FB_TM_DP = 128, FB_TM_DV = 256, FB_TM_DK = 320, FB_TM_DQ = 384;

struct FlashBwdParams {
  CUtensorMap tm_q, tm_k, tm_v, tm_do;
  const float* lse;     // [batch][heads][sq]  log2 domain
  const float* dvec;    // [batch][heads][sq]  rowsum(dO o O)
  float* dq_acc;        // [batch][heads][sq][64] fp32, zeroed by the caller
  __nv_bfloat16* dk;    // may be null
  __nv_bfloat16* dv;    // may be null
  long long ld_dk, ld_dv;
  int sq, skv, heads, d, n_q_tiles;
  float scale, scale_log2;
};

// MN-major operand tile as stored ([K rows][64 MN elements = 128 B], 128B swizzle): LBO = distance between 64-wide MN
// blocks (16 KiB for the two key chunks of P / dS), SBO = 8 K-rows.
__device__ __forceinline__ uint64_t fb_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(FB_TILE >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t fb_idesc(uint32_t n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) |
         ((128u >> 4) << 24);
}
__device__ __forceinline__ float fb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(FB_THREADS, 1) flash_attn_bwd_kernel(const __grid_constant__ FlashBwdParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + FB_TILE;
  uint8_t* sQ = sV + FB_TILE;           // [2]
  uint8_t* sdO = sQ + 2 * FB_TILE;      // [2]
  uint8_t* sP = sdO + 2 * FB_TILE;
  uint8_t* sdS = sP + FB_PS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + FB_PS);
  uint64_t* kv_full = bars;             // [1]
  uint64_t* qdo_full = bars + 1;        // [2]
  uint64_t* qdo_empty = bars + 3;       // [2]
  uint64_t* s_full = bars + 5;          // [1]  S and dP ready
  uint64_t* s_empty = bars + 6;         // [1]  count 8
  uint64_t* p_full = bars + 7;          // [1]  count 8: P and dS staged
  uint64_t* p_empty = bars + 8;         // [1]
  uint64_t* dq_full = bars + 9;         // [1]
  uint64_t* dq_empty = bars + 10;       // [1]  count 8
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nq = p.n_q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
    tma_prefetch_desc(&p.tm_do);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * FB_TILE);
      tma_load_4d(sK, &p.tm_k, kv_full, 0, jt * FB_M, head, b);
      tma_load_4d(sV, &p.tm_v, kv_full, 0, jt * FB_M, head, b);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * FB_TILE);
        tma_load_4d(sQ + st * FB_TILE, &p.tm_q, &qdo_full[st], 0, i * FB_M, head, b);
        tma_load_4d(sdO + st * FB_TILE, &p.tm_do, &qdo_full[st], 0, i * FB_M, head, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t id_s = fb_idesc(128, false, false);    // S, dP : A K-major, B K-major
      const uint32_t id_kv = fb_idesc(64, true, true);      // dV, dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      const uint32_t id_q = fb_idesc(64, false, true);      // dQ    : A = dS (K-major), B = K (MN-major)
      const uint64_t dk_k = umma_desc_k_sw128(smem_u32(sK)), dv_k = umma_desc_k_sw128(smem_u32(sV));
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        mbar_wait(&qdo_full[st], (i >> 1) & 1);
        mbar_wait(s_empty, (i & 1) ^ 1);
        tc_fence_after();
        const uint64_t dq_ = umma_desc_k_sw128(smem_u32(sQ + st * FB_TILE));
        const uint64_t ddo = umma_desc_k_sw128(smem_u32(sdO + st * FB_TILE));
#pragma unroll
        for (int s = 0; s < FB_D / 16; ++s) umma_bf16(tmem_base + FB_TM_S, dq_ + 2 * s, dk_k + 2 * s, id_s, s > 0 ? 1u : 0u);
#pragma unroll
        for (int s = 0; s < FB_D / 16; ++s) umma_bf16(tmem_base + FB_TM_DP, ddo + 2 * s, dv_k + 2 * s, id_s, s > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int i = 0; i < nq; ++i) {
        if (i + 1 < nq) issue_sdp(i + 1);
        const int st = i & 1;
        mbar_wait(p_full, i & 1);
        tc_fence_after();
        const uint32_t pb = smem_u32(sP), sb = smem_u32(sdS);
        const uint32_t qb = smem_u32(sQ + st * FB_TILE), ob = smem_u32(sdO + st * FB_TILE), kb = smem_u32(sK);
#pragma unroll
        for (int s = 0; s < FB_M / 16; ++s)   // dV[key][d] += sum_q P[q][key] dO[q][d]
          umma_bf16(tmem_base + FB_TM_DV, fb_desc_mn(pb + s * 16 * 128), fb_desc_mn(ob + s * 16 * 128), id_kv,
                    (i > 0 || s > 0) ? 1u : 0u);
#pragma unroll
        for (int s = 0; s < FB_M / 16; ++s)   // dK[key][d] += sum_q dS[q][key] Q[q][d]
          umma_bf16(tmem_base + FB_TM_DK, fb_desc_mn(sb + s * 16 * 128), fb_desc_mn(qb + s * 16 * 128), id_kv,
                    (i > 0 || s > 0) ? 1u : 0u);
        mbar_wait(dq_empty, (i & 1) ^ 1);     // the previous step's dQ has been read out of TMEM
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < FB_M / 16; ++s)   // dQ[q][d] = sum_key dS[q][key] K[key][d]
          umma_bf16(tmem_base + FB_TM_DQ, umma_desc_k_sw128(sb + (s >> 2) * (FB_PS / 2)) + 2 * (s & 3),
                    fb_desc_mn(kb + s * 16 * 128), id_q, s > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(p_empty);
        umma_commit(&qdo_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax-gradient warps + outputs
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const long long bh = static_cast<long long>(b) * p.heads + head;
    const int key0 = jt * FB_M + wg * 64;     // first key of this thread's half

    auto readout_dq = [&](int i) {            // dq_acc[b,h,i*128+r, 32 wg .. +32) += dQ_i
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_DQ + wg * 32, raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      const int row = i * FB_M + r;
      if (row < p.sq) {
        float* dst = p.dq_acc + (bh * p.sq + row) * FB_D + wg * 32;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(__uint_as_float(raw[g * 4 + 0])),
                       "f"(__uint_as_float(raw[g * 4 + 1])), "f"(__uint_as_float(raw[g * 4 + 2])),
                       "f"(__uint_as_float(raw[g * 4 + 3]))
                       : "memory");
      }
    };

    for (int i = 0; i < nq; ++i) {
      const int row = i * FB_M + r;
      const bool row_ok = row < p.sq;
      const float lse_r = row_ok ? __ldg(p.lse + bh * p.sq + row) : 0.f;
      const float d_r = row_ok ? __ldg(p.dvec + bh * p.sq + row) : 0.f;
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      // two passes of 32 keys each keep the live registers at 64 TMEM words
      uint32_t s0[32], g0[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_S + wg * 64, s0);
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_DP + wg * 64, g0);
      tmem_ld_wait();
      // P / dS row pieces (this half = 64 keys = one 128-byte row of chunk `wg`)
      mbar_wait(p_empty, (i & 1) ^ 1);
      uint8_t* prow = sP + wg * (FB_PS / 2) + r * 128;
      uint8_t* drow = sdS + wg * (FB_PS / 2) + r * 128;
#define FB_HALF(SRC, GRD, KOFF, CHUNK0)                                                              \
  _Pragma("unroll") for (int t = 0; t < 4; ++t) {                                                    \
    uint32_t pk[4], dk_[4];                                                                          \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                                  \
      const int e = t * 8 + u * 2;                                                                   \
      float p0 = fb_ex2(fmaf(__uint_as_float(SRC[e]), p.scale_log2, -lse_r));                        \
      float p1 = fb_ex2(fmaf(__uint_as_float(SRC[e + 1]), p.scale_log2, -lse_r));                    \
      if (!row_ok || key0 + KOFF + e >= p.skv) p0 = 0.f;                                             \
      if (!row_ok || key0 + KOFF + e + 1 >= p.skv) p1 = 0.f;                                         \
      const float d0 = p0 * (__uint_as_float(GRD[e]) - d_r) * p.scale;                               \
      const float d1 = p1 * (__uint_as_float(GRD[e + 1]) - d_r) * p.scale;                           \
      pk[u] = pack_bf16(p0, p1);                                                                     \
      dk_[u] = pack_bf16(d0, d1);                                                                    \
    }                                                                                                \
    const int off = (((CHUNK0 + t) ^ (r & 7)) << 4);                                                 \
    *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);                  \
    *reinterpret_cast<uint4*>(drow + off) = make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]);              \
  }
      FB_HALF(s0, g0, 0, 0)
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_S + wg * 64 + 32, s0);
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_DP + wg * 64 + 32, g0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);      // S / dP fully read: the next step's S / dP may overwrite them
      FB_HALF(s0, g0, 32, 4)
#undef FB_HALF
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (i > 0) readout_dq(i - 1);
    }
    readout_dq(nq - 1);   // its commit also covers the last dV / dK updates
    // dV / dK: TMEM lanes are KEY rows here
    const int key = jt * FB_M + r;
    {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_DV + wg * 32, raw);
      tmem_ld_wait();
      if (p.dv && key < p.skv) {
        __nv_bfloat16* dst = p.dv + (static_cast<long long>(b) * p.skv + key) * p.ld_dv + head * p.d + wg * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8)
          if (wg * 32 + c8 * 8 < p.d)
            *reinterpret_cast<uint4*>(dst + c8 * 8) =
                make_uint4(pack_bf16(__uint_as_float(raw[c8 * 8 + 0]), __uint_as_float(raw[c8 * 8 + 1])),
                           pack_bf16(__uint_as_float(raw[c8 * 8 + 2]), __uint_as_float(raw[c8 * 8 + 3])),
                           pack_bf16(__uint_as_float(raw[c8 * 8 + 4]), __uint_as_float(raw[c8 * 8 + 5])),
                           pack_bf16(__uint_as_float(raw[c8 * 8 + 6]), __uint_as_float(raw[c8 * 8 + 7])));
      }
      tmem_ld_32x32b_x32(tmem_base + lane_off + FB_TM_DK + wg * 32, raw);
      tmem_ld_wait();
      if (p.dk && key < p.skv) {
        __nv_bfloat16* dst = p.dk + (static_cast<long long>(b) * p.skv + key) * p.ld_dk + head * p.d + wg * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8)
          if (wg * 32 + c8 * 8 < p.d)
            *reinterpret_cast<uint4*>(dst + c8 * 8) =
                make_uint4(pack_bf16(__uint_as_float(raw[c8 * 8 + 0]), __uint_as_float(raw[c8 * 8 + 1])),
                           pack_bf16(__uint_as_float(raw[c8 * 8 + 2]), __uint_as_float(raw[c8 * 8 + 3])),
                           pack_bf16(__uint_as_float(raw[c8 * 8 + 4]), __uint_as_float(raw[c8 * 8 + 5])),
                           pack_bf16(__uint_as_float(raw[c8 * 8 + 6]), __uint_as_float(raw[c8 * 8 + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// D[b][h][row] = sum_c dO[row][h d + c] * O[row][h d + c]   (one thread per (row, head))
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, long long ldo,
                                     const __nv_bfloat16* __restrict__ dout, long long lddo, float* __restrict__ dvec,
                                     int batch, int heads, int sq, int d) {
  pdl_entry();
  const long long total = (long long)batch * heads * sq;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx % sq);
    const long long bh = idx / sq;
    const int h = (int)(bh % heads);
    const long long bb = bh / heads;
    const __nv_bfloat16* po = o + (bb * sq + row) * ldo + h * d;
    const __nv_bfloat16* pd = dout + (bb * sq + row) * lddo + h * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(po + c);
      const uint4 g = *reinterpret_cast<const uint4*>(pd + c);
      acc = fmaf(bf16_lo(a.x), bf16_lo(g.x), acc);
      acc = fmaf(bf16_hi(a.x), bf16_hi(g.x), acc);
      acc = fmaf(bf16_lo(a.y), bf16_lo(g.y), acc);
      acc = fmaf(bf16_hi(a.y), bf16_hi(g.y), acc);
      acc = fmaf(bf16_lo(a.z), bf16_lo(g.z), acc);
      acc = fmaf(bf16_hi(a.z), bf16_hi(g.z), acc);
      acc = fmaf(bf16_lo(a.w), bf16_lo(g.w), acc);
      acc = fmaf(bf16_hi(a.w), bf16_hi(g.w), acc);
    }
    dvec[idx] = acc;
  }
}

// dq[b*sq + row][h d + c] = bf16(dq_acc[b][h][row][c])
__global__ void attn_dq_cast_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, long long lddq,
                                    int batch, int heads, int sq, int d) {
  pdl_entry();
  const int vpr = d / 8;
  const long long total = (long long)batch * heads * sq * vpr;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % vpr);
    const long long t = idx / vpr;
    const int row = (int)(t % sq);
    const long long bh = t / sq;
    const int h = (int)(bh % heads);
    const long long bb = bh / heads;
    const float4 a = *reinterpret_cast<const float4*>(acc + (bh * sq + row) * FB_D + cv * 8);
    const float4 c = *reinterpret_cast<const float4*>(acc + (bh * sq + row) * FB_D + cv * 8 + 4);
    *reinterpret_cast<uint4*>(dq + (bb * sq + row) * lddq + h * d + cv * 8) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
  }
}

}  // namespace leco

using namespace leco;

extern "C" int leco_attn_bwd_prep(const void* o, int64_t ldo, const void* dout, int64_t lddo, float* dvec, int batch,
                                  int heads, int sq, int d, void* stream) {
  LECO_REQUIRE(o && dout && dvec && d % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0, "leco_attn_bwd_prep: bad arguments");
  const long long total = (long long)batch * heads * sq;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  LECO_LAUNCH(attn_bwd_prep_kernel, (int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream),
              reinterpret_cast<const __nv_bfloat16*>(o), (long long)ldo, reinterpret_cast<const __nv_bfloat16*>(dout),
              (long long)lddo, dvec, batch, heads, sq, d);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int leco_attn_dq_cast(const float* dq_acc, void* dq, int64_t lddq, int batch, int heads, int sq, int d,
                                 void* stream) {
  LECO_REQUIRE(dq_acc && dq && d % 8 == 0 && lddq % 8 == 0, "leco_attn_dq_cast: bad arguments");
  const long long total = (long long)batch * heads * sq * (d / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  LECO_LAUNCH(attn_dq_cast_kernel, (int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream), dq_acc,
              reinterpret_cast<__nv_bfloat16*>(dq), (long long)lddq, batch, heads, sq, d);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// q/k/v/dout: [rows, ld] bf16 buffers (head h in columns [h*d, h*d+d)); lse / dvec: fp32 [batch][heads][sq];
// dq_acc: ZEROED fp32 [batch][heads][sq][64]; dk / dv: [batch*skv, ld] bf16 outputs (either may be NULL).
extern "C" int leco_flash_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                   const void* dout, int64_t lddo, const float* lse, const float* dvec, float* dq_acc,
                                   void* dk, int64_t lddk, void* dv, int64_t lddv, int batch, int heads, int sq, int skv,
                                   int d, float scale, void* stream) {
  LECO_REQUIRE(q && k && v && dout && lse && dvec && dq_acc, "leco_flash_attn_bwd: null pointer");
  LECO_REQUIRE(d % 8 == 0 && d <= FB_D, "leco_flash_attn_bwd: head dim %d unsupported (<=64, multiple of 8)", d);
  LECO_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
               "leco_flash_attn_bwd: strides must be multiples of 8");
  FlashBwdParams p;
  memset(&p, 0, sizeof(p));
  const uint32_t box[4] = {FB_D, FB_M, 1, 1};
  auto make = [&](CUtensorMap* m, const void* base, int64_t ld, int rows) {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)rows, (uint64_t)heads, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)d * 2, (uint64_t)ld * rows * 2};
    return make_tmap_bf16_4d(m, base, dims, str, box);
  };
  if (make(&p.tm_q, q, ldq, sq) || make(&p.tm_k, k, ldk, skv) || make(&p.tm_v, v, ldv, skv) || make(&p.tm_do, dout, lddo, sq))
    return -3;
  p.lse = lse;
  p.dvec = dvec;
  p.dq_acc = dq_acc;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.ld_dk = lddk;
  p.ld_dv = lddv;
  p.sq = sq;
  p.skv = skv;
  p.heads = heads;
  p.d = d;
  p.n_q_tiles = (sq + FB_M - 1) / FB_M;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
    attr_set = true;
  }
  dim3 grid((skv + FB_M - 1) / FB_M, heads, batch);
  count_launch();
  LECO_LAUNCH(flash_attn_bwd_kernel, grid, FB_THREADS, FB_SMEM, reinterpret_cast<cudaStream_t>(stream), p);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
