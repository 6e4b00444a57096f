// Fused attention forward for sm_100a: O = softmax(scale * Q K^T) V per (sample, head), head dim <= 64 (two pipelined
// kernels below) or <= 192 (flash_attn_fwd_wide_kernel at the end of the file).
// Replaces xformers.memory_efficient_attention on the reference's path (train_lora.py:68).
//
// One CTA per (128-query tile, head, sample).  K/V tiles of 128 keys stream through TMA rings;
//   S_j  = Q K_j^T      tcgen05.mma (M128 N128 K64)  -> TMEM, double buffered
//   P_j  = exp2(S_j*c - m)  softmax warps read S from TMEM (tcgen05.ld), keep running max / sum in
//                            registers, write P (bf16) into 128B-swizzled smem as the next A operand
//   PV_j = P_j V_j      tcgen05.mma (M128 N64 K128)  -> TMEM, double buffered
//   O    = O*alpha + PV_j   folded in registers by the softmax warps one tile late, so the tensor pipe
//                            never waits for a TMEM rescale.
// Nothing of S or P ever touches HBM: algorithmic traffic is Q, K, V, O once (K/V re-reads hit L2).
// Roles: warp 0 lane 0 TMA, warp 1 lane 0 MMA issue, warp 2 TMEM alloc, warps 4-7 and 8-11 = two softmax
// warpgroups: warpgroup w owns the KV tiles j = w (mod 2) with its own online-softmax state per row; the two partial
// results (max, sum, 64 output columns) are merged once per CTA at the end.  exp2 is the MUFU-bound part (16 / clk / SM
// measured, tests/gpu_checks/micro_probe.cu): two decoupled warpgroups keep it fed.
//
// V operand: v_mode 0 = V tile [keys, d] used directly as an MN-major B operand;
//            v_mode 1 = a pre-transposed V^T [d, keys] (K-major B operand, like the GEMM kernel).
#include "../../include/leco_b200.h"
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace leco {
void count_launch();

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64;
constexpr int FA_KS = 3, FA_VS = 3;                       // K / V ring depth
constexpr int FA_Q_BYTES = FA_BM * FA_D * 2;              // 16 KiB
constexpr int FA_KV_BYTES = FA_BN * FA_D * 2;             // 16 KiB
constexpr int FA_P_BYTES = FA_BM * FA_BN * 2;             // 32 KiB (two 64-key chunks)
constexpr int FA_SMEM = FA_Q_BYTES + (FA_KS + FA_VS) * FA_KV_BYTES + 2 * FA_P_BYTES + 1024 + 512 + (2 * 2 + 2) * 128 * 4;
constexpr int FA_TMEM_COLS = 512;                         // S: 3 x 128 (ring), PV: 2 x 64
constexpr int FA_SB = 3;                                  // S ring depth
constexpr int FA_THREADS = 384;                          // 4 control warps + 2 softmax warpgroups

struct FlashParams {
  CUtensorMap tm_q, tm_k, tm_v;
  __nv_bfloat16* out;
  long long ld_out;
  int sq, skv, heads, d, n_kv_tiles, v_mode;
  float scale_log2;  // softmax scale * log2(e)
  float* lse;        // optional [batch][heads][sq]: log2-domain log-sum-exp (max*scale_log2 + log2(sum)) for the backward
};

// MN-major B operand (V tile as stored: keys x d), 128B swizzle: 8 key-rows of 128 B per atom.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((FA_KV_BYTES) >> 4) << 16;  // LBO: next 64-wide MN block (unused: N = 64)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;           // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(FA_THREADS, 1) flash_attn_fwd_kernel(const __grid_constant__ FlashParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;
  uint8_t* sV = sK + FA_KS * FA_KV_BYTES;
  uint8_t* sP = sV + FA_VS * FA_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * FA_P_BYTES);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [KS]
  uint64_t* k_empty = k_full + FA_KS;      // [KS]
  uint64_t* v_full = k_empty + FA_KS;      // [VS]
  uint64_t* v_empty = v_full + FA_VS;      // [VS]
  uint64_t* s_full = v_empty + FA_VS;      // [FA_SB]
  uint64_t* s_empty = s_full + FA_SB;      // [FA_SB]
  uint64_t* p_full = s_empty + FA_SB;      // [2]
  uint64_t* p_empty = p_full + 2;          // [2]
  uint64_t* pv_full = p_empty + 2;         // [2]
  uint64_t* pv_empty = pv_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_empty + 2);
  float* mx_buf = reinterpret_cast<float*>(tmem_slot + 2);  // [2 tiles][2 warpgroups][128 rows] maxima + [2][128] sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < FA_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < FA_SB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);   // one softmax warpgroup (4 warps) consumes a score tile
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
      mbar_init(&pv_full[i], 1);
      mbar_init(&pv_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_s = tmem_base;                    // + (j % 3) * 128: three score tiles in flight, so QK^T of tile
                                                        // j+2 never waits for the warpgroup still reading tile j
  const uint32_t tmem_pv = tmem_base + FA_SB * FA_BN;   // + (j & 1) * 64

  // Each role below is ONE elected lane running its whole loop (tests/gpu_checks/mma_probe.cu: re-electing the warp
  // every iteration costs ~700 cycles per round against ~60 for the single-lane loop).
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, FA_Q_BYTES);
      tma_load_4d(sQ, &p.tm_q, q_full, 0, qt * FA_BM, head, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % FA_KS, vs = j % FA_VS;
        mbar_wait(&k_empty[ks], ((j / FA_KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], FA_KV_BYTES);
        tma_load_4d(sK + ks * FA_KV_BYTES, &p.tm_k, &k_full[ks], 0, j * FA_BN, head, b);
        mbar_wait(&v_empty[vs], ((j / FA_VS) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], FA_KV_BYTES);
        if (p.v_mode == 0) {
          tma_load_4d(sV + vs * FA_KV_BYTES, &p.tm_v, &v_full[vs], 0, j * FA_BN, head, b);
        } else {  // V^T [d, keys]: two 64-key chunks, each [64 d-rows x 128 B]
          tma_load_4d(sV + vs * FA_KV_BYTES, &p.tm_v, &v_full[vs], j * FA_BN, 0, head, b);
          tma_load_4d(sV + vs * FA_KV_BYTES + FA_KV_BYTES / 2, &p.tm_v, &v_full[vs], j * FA_BN + 64, 0, head, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(FA_BN, false);
      const uint32_t idesc_pv = umma_idesc_bf16(FA_D, p.v_mode == 0);
      const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const int ks = j % FA_KS, sb = j % FA_SB;
        mbar_wait(&k_full[ks], (j / FA_KS) & 1);
        mbar_wait(&s_empty[sb], ((j / FA_SB) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + ks * FA_KV_BYTES));
#pragma unroll
        for (int s = 0; s < FA_D / 16; ++s) umma_bf16(tmem_s + sb * FA_BN, dq + 2 * s, dk + 2 * s, idesc_s, s > 0 ? 1u : 0u);
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[ks]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 2 < n_tiles) issue_s(j + 2);
        const int vs = j % FA_VS, pb = j & 1;
        mbar_wait(&p_full[pb], (j >> 1) & 1);
        mbar_wait(&v_full[vs], (j / FA_VS) & 1);
        mbar_wait(&pv_empty[pb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t pbase = smem_u32(sP + pb * FA_P_BYTES);
        const uint32_t vbase = smem_u32(sV + vs * FA_KV_BYTES);
#pragma unroll
        for (int s = 0; s < FA_BN / 16; ++s) {
          // A = P: K-major, two 64-key chunks of [128 rows x 128 B]
          const uint64_t da = umma_desc_k_sw128(pbase + (s >> 2) * (FA_P_BYTES / 2)) + 2 * (s & 3);
          uint64_t db;
          if (p.v_mode == 0)
            db = umma_desc_mn_sw128(vbase + s * 16 * 128);                       // 16 key rows per k-step
          else
            db = umma_desc_k_sw128(vbase + (s >> 2) * (FA_KV_BYTES / 2)) + 2 * (s & 3);
          umma_bf16(tmem_pv + pb * FA_D, da, db, idesc_pv, s > 0 ? 1u : 0u);
        }
        umma_commit(&pv_full[pb]);
        umma_commit(&v_empty[vs]);
        umma_commit(&p_empty[pb]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output
    // Warpgroup w owns the KV tiles j = w, w+2, w+4, ... (all 128 keys of them): thread (w, row) keeps its OWN online-
    // softmax state (running max, row sum, 64 output columns) over those tiles, so the two warpgroups never talk to
    // each other inside the loop (the per-tile max exchange through a named barrier was the top stall of the previous
    // version: ncu `stalled_barrier` 1.56 per issue) and each has two tile periods to turn a tile around.  The two
    // partial results are merged once at the end (split-KV combine).
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;                  // TMEM lane quadrant
    const int r = q * 32 + lane;             // row of the query tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o[FA_D];
#pragma unroll
    for (int i = 0; i < FA_D; ++i) o[i] = 0.f;

    auto fold_pv = [&](int j, float alpha) {  // o = (o + PV_j) * alpha   (buffer j & 1 == wg; alpha == 1 almost always)
      const int pb = j & 1;
      mbar_wait(&pv_full[pb], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(tmem_pv + lane_off + pb * FA_D + h * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[h * 32 + i] += __uint_as_float(raw[i]);
      }
      if (alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < FA_D; ++i) o[i] *= alpha;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pv_empty[pb]);
    };

    // One KV tile of the online softmax.  RG (= the tile runs past skv) and FIRST (= this warpgroup's first tile) are
    // compile-time flags: as run-time tests the per-score masks become predicated instructions that take issue slots.
    // LAZY MAX: only the first tile runs a separate max pass.  Every later tile exponentiates against the running max
    // of the tiles BEFORE it (p may exceed 1 when the tile raises the max: harmless in bf16 / fp32) while tracking its
    // own max in the same loop; the state is then rescaled once by alpha = 2^((m_old - m_new) c), which is exactly 1
    // whenever the max did not move.  No second pass over TMEM, no max -> exp dependency inside a tile.
    auto softmax_tile = [&](int j, auto rg_tag, auto first_tag) {
      constexpr bool ragged = decltype(rg_tag)::value;
      constexpr bool first = decltype(first_tag)::value;
      const int sb = j & 1;          // P / PV buffer of this tile
      const int s3 = j % FA_SB;      // score buffer of this tile
      const int kv0 = j * FA_BN;
      // this warpgroup's previous tile first: frees the PV accumulator before P_j is even staged
      if (j >= 2) fold_pv(j - 2, alpha_prev);
      mbar_wait(&s_full[s3], (j / FA_SB) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_s + lane_off + s3 * FA_BN;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      float m_ref = m_run;
      if constexpr (first) {
        // row max over the 128 scores, 32 at a time (four independent running maxima)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]), a2 = __uint_as_float(sv[i + 2]),
                  a3 = __uint_as_float(sv[i + 3]);
            if constexpr (ragged) {
              if (kv0 + c * 32 + i >= p.skv) a0 = -INFINITY;
              if (kv0 + c * 32 + i + 1 >= p.skv) a1 = -INFINITY;
              if (kv0 + c * 32 + i + 2 >= p.skv) a2 = -INFINITY;
              if (kv0 + c * 32 + i + 3 >= p.skv) a3 = -INFINITY;
            }
            mx0 = fmaxf(mx0, a0);
            mx1 = fmaxf(mx1, a1);
            mx2 = fmaxf(mx2, a2);
            mx3 = fmaxf(mx3, a3);
          }
        }
        m_ref = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      const float mb = m_ref * p.scale_log2;
      // p = 2^(s*c - m_ref*c) -> bf16 -> swizzled A-operand tile (two 64-key chunks of [128 rows x 128 B])
      mbar_wait(&p_empty[sb], ((j >> 1) & 1) ^ 1);
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;  // four partial sums: no long dependent FADD chain
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, sv);
        tmem_ld_wait();
        if (c == 3) {               // S fully consumed: the next QK^T of this buffer may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[s3]);
        }
        uint8_t* prow = sP + sb * FA_P_BYTES + (c >> 1) * (FA_P_BYTES / 2) + r * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint32_t pk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = t * 8 + u * 2;
            float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]);
            float p0 = ex2_approx(fmaf(a0, p.scale_log2, -mb));
            float p1 = ex2_approx(fmaf(a1, p.scale_log2, -mb));
            if constexpr (ragged) {
              if (kv0 + c * 32 + i >= p.skv) { p0 = 0.f; a0 = -INFINITY; }
              if (kv0 + c * 32 + i + 1 >= p.skv) { p1 = 0.f; a1 = -INFINITY; }
            }
            if constexpr (!first) {
              if (u == 0) mx0 = fmaxf(mx0, fmaxf(a0, a1)); else if (u == 1) mx1 = fmaxf(mx1, fmaxf(a0, a1));
              else if (u == 2) mx2 = fmaxf(mx2, fmaxf(a0, a1)); else mx3 = fmaxf(mx3, fmaxf(a0, a1));
            }
            if (u == 0) rs0 += p0 + p1; else if (u == 1) rs1 += p0 + p1; else if (u == 2) rs2 += p0 + p1; else rs3 += p0 + p1;
            pk[u] = pack_bf16(p0, p1);
          }
          *reinterpret_cast<uint4*>(prow + ((((c & 1) * 4 + t) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
      float alpha = 1.0f, m_new = m_ref;
      if constexpr (!first) {
        m_new = fmaxf(fmaxf(m_run, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
        alpha = ex2_approx((m_run - m_new) * p.scale_log2);   // exactly 1 when the max did not move
      }
      l_run = (l_run + ((rs0 + rs1) + (rs2 + rs3))) * alpha;
      m_run = m_new;
      alpha_prev = alpha;       // applied to o when this tile's PV is folded: o = (o + PV_j) * alpha
    };
    int last = -1;
    for (int j = wg; j < n_tiles; j += 2) {
      const bool rg = j * FA_BN + FA_BN > p.skv;
      if (j == wg) {
        if (rg) softmax_tile(j, std::true_type{}, std::true_type{});
        else softmax_tile(j, std::false_type{}, std::true_type{});
      } else {
        if (rg) softmax_tile(j, std::true_type{}, std::false_type{});
        else softmax_tile(j, std::false_type{}, std::false_type{});
      }
      last = j;
    }
    if (last >= 0) fold_pv(last, alpha_prev);
    // ---- merge the two partial softmax results: each warpgroup finishes 32 of the 64 output columns.  The partner's
    // half travels through this warpgroup's own P buffer (free: its last PV has been folded), column-major so that
    // consecutive lanes hit consecutive words.  (Static register indices only: `o` must stay in registers.)
    {
      float* xo = reinterpret_cast<float*>(sP + wg * FA_P_BYTES);         // [32 cols][128 rows]
      if (wg == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i * 128 + r] = o[32 + i];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i * 128 + r] = o[i];
      }
      mx_buf[wg * 128 + r] = m_run;
      mx_buf[256 + wg * 128 + r] = l_run;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the two softmax warpgroups only, once per CTA
    {
      const float m_o = mx_buf[(wg ^ 1) * 128 + r], l_o = mx_buf[256 + (wg ^ 1) * 128 + r];
      const float m_tot = fmaxf(m_run, m_o);          // finite: tile 0 has at least one valid key
      const float a_me = ex2_approx((m_run - m_tot) * p.scale_log2);
      const float a_ot = ex2_approx((m_o - m_tot) * p.scale_log2);
      const float l_tot = l_run * a_me + l_o * a_ot;
      const float* xi = reinterpret_cast<const float*>(sP + (wg ^ 1) * FA_P_BYTES);
      const int row = qt * FA_BM + r;
      const float inv = 1.0f / l_tot;
      float f[32];
      if (wg == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = (o[i] * a_me + xi[i * 128 + r] * a_ot) * inv;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = (o[32 + i] * a_me + xi[i * 128 + r] * a_ot) * inv;
      }
      if (row < p.sq) {
        if (p.lse && wg == 0)
          p.lse[(static_cast<long long>(b) * p.heads + head) * p.sq + row] = m_tot * p.scale_log2 + log2f(l_tot);
        __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.sq + row) * p.ld_out + head * p.d + wg * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          if (wg * 32 + c8 * 8 < p.d) {
            uint4 v;
            v.x = pack_bf16(f[c8 * 8 + 0], f[c8 * 8 + 1]);
            v.y = pack_bf16(f[c8 * 8 + 2], f[c8 * 8 + 3]);
            v.z = pack_bf16(f[c8 * 8 + 4], f[c8 * 8 + 5]);
            v.w = pack_bf16(f[c8 * 8 + 6], f[c8 * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + c8 * 8) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, FA_TMEM_COLS);
}


// ----------------------------------------------------------------------------------------------------------------------
// Variant with P and O kept in tensor memory (default; LECO_FLASH_TS=0 selects the kernel above).
//   * P_j (bf16) is written back INTO the score slot it was computed from (tcgen05.st, two keys per 32-bit column) and
//     the PV product reads it from there as a TMEM A operand: no shared-memory P tile, no st.shared / proxy fence, and
//     no p_empty wait (the slot's next user is a later QK^T, ordered behind this PV by the in-order tensor pipe);
//   * each warpgroup's PV products ACCUMULATE in its own 64 TMEM columns over all of its tiles.  The softmax warps
//     therefore never read a PV result inside the loop (the per-tile fold and its pv_full wait were 17 % of their
//     stall samples, profiles/r2k_ncu_full_flash_fwd.txt).  The running max is lazy with a threshold: exponentials use
//     a reference max that is only advanced when the true max has moved by more than 2^8 (P <= 256, harmless in
//     bf16 / fp32); only then is the accumulator rescaled in TMEM, at the start of the warpgroup's next tile.
// TMEM: 3 score/probability slots x 128 columns, then 2 x 64 output columns.
constexpr int FT_SMEM = FA_Q_BYTES + (FA_KS + FA_VS) * FA_KV_BYTES + 2 * (32 * 128 * 4) + 1024 + 512 + (2 * 2 + 2) * 128 * 4;
constexpr float FT_RESCALE_LOG2 = 8.0f;
// Which of the four score pairs of every 8 go through the FMA-pipe polynomial instead of the MUFU (bit u = pair u):
// 0b1000 = a quarter of the exponentials.  ex2 runs at 16 / clk / SM (profiles/r2_micro_probe.txt) and is the floor of
// this kernel; a degree-3 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below bf16's 2^-9) in packed f32x2
// arithmetic moves part of that load to the FMA pipe, which has issue slots to spare once the rest of the row update
// is packed too (FFMA2 scale/shift, FADD2 row sums, 3-input max).
constexpr int FT_POLY_PAIRS = 0b1000;

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_pack_bits(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// 2^x for a packed pair on the FMA pipe: n = round(x) via the 1.5 * 2^23 trick, 2^(x - n) by a cubic, exponent add.
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& r0, float& r1) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  x2 = f2_pack(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t magic = f2_pack(12582912.0f, 12582912.0f), nmagic = f2_pack(-12582912.0f, -12582912.0f);
  const uint64_t t2 = f2_add(x2, magic);                                    // low mantissa bits = round(x)
  const uint64_t n2 = f2_add(t2, nmagic);
  const uint64_t f2 = f2_fma(n2, f2_pack(-1.0f, -1.0f), x2);                // x - round(x) in [-0.5, 0.5]
  uint64_t q2 = f2_fma(f2, f2_pack(0.05517164617776871f, 0.05517164617776871f), f2_pack(0.2426111251115799f, 0.2426111251115799f));
  q2 = f2_fma(q2, f2, f2_pack(0.6932609677314758f, 0.6932609677314758f));
  q2 = f2_fma(q2, f2, f2_pack(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  f2_unpack(q2, q0, q1);
  f2_unpack(t2, t0, t1);
  r0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

template <int POLY = FT_POLY_PAIRS>
__global__ void __launch_bounds__(FA_THREADS, 1) flash_attn_fwd_ts_kernel(const __grid_constant__ FlashParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;
  uint8_t* sV = sK + FA_KS * FA_KV_BYTES;
  float* sX = reinterpret_cast<float*>(sV + FA_VS * FA_KV_BYTES);   // [2 warpgroups][32 cols][128 rows] merge exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * 32 * 128);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [KS]
  uint64_t* k_empty = k_full + FA_KS;      // [KS]
  uint64_t* v_full = k_empty + FA_KS;      // [VS]
  uint64_t* v_empty = v_full + FA_VS;      // [VS]
  uint64_t* s_full = v_empty + FA_VS;      // [FA_SB]  QK^T of the slot finished
  uint64_t* p_full = s_full + FA_SB;       // [FA_SB]  P written into the slot (and the accumulator rescaled if needed)
  uint64_t* pv_done = p_full + FA_SB;      // [2]      per warpgroup: its latest PV finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* mx_buf = reinterpret_cast<float*>(tmem_slot + 2);  // [2 warpgroups][128 rows] maxima + [2][128] sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < FA_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < FA_SB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);    // the four warps of the warpgroup that owns the tile
    }
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_s = tmem_base;                    // + (j % 3) * 128
  const uint32_t tmem_o = tmem_base + FA_SB * FA_BN;    // + warpgroup * 64

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, FA_Q_BYTES);
      tma_load_4d(sQ, &p.tm_q, q_full, 0, qt * FA_BM, head, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % FA_KS, vs = j % FA_VS;
        mbar_wait(&k_empty[ks], ((j / FA_KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], FA_KV_BYTES);
        tma_load_4d(sK + ks * FA_KV_BYTES, &p.tm_k, &k_full[ks], 0, j * FA_BN, head, b);
        mbar_wait(&v_empty[vs], ((j / FA_VS) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], FA_KV_BYTES);
        if (p.v_mode == 0) {
          tma_load_4d(sV + vs * FA_KV_BYTES, &p.tm_v, &v_full[vs], 0, j * FA_BN, head, b);
        } else {  // V^T [d, keys]: two 64-key chunks, each [64 d-rows x 128 B]
          tma_load_4d(sV + vs * FA_KV_BYTES, &p.tm_v, &v_full[vs], j * FA_BN, 0, head, b);
          tma_load_4d(sV + vs * FA_KV_BYTES + FA_KV_BYTES / 2, &p.tm_v, &v_full[vs], j * FA_BN + 64, 0, head, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(FA_BN, false);
      const uint32_t idesc_pv = umma_idesc_bf16(FA_D, p.v_mode == 0);
      const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
      // QK^T of tile j into slot j % 3.  The slot last held P_{j-3}; PV_{j-3} was issued before this call (program
      // order of this lane) and the tensor pipe executes in issue order, so no barrier guards the reuse.  The softmax
      // warps are done with the slot as well: p_full of tile j-3 was waited for before PV_{j-3} was issued.
      auto issue_s = [&](int j) {
        const int ks = j % FA_KS, sb = j % FA_SB;
        mbar_wait(&k_full[ks], (j / FA_KS) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + ks * FA_KV_BYTES));
#pragma unroll
        for (int s = 0; s < FA_D / 16; ++s) umma_bf16(tmem_s + sb * FA_BN, dq + 2 * s, dk + 2 * s, idesc_s, s > 0 ? 1u : 0u);
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[ks]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 2 < n_tiles) issue_s(j + 2);
        const int vs = j % FA_VS, sb = j % FA_SB, wg = j & 1;
        mbar_wait(&p_full[sb], (j / FA_SB) & 1);
        mbar_wait(&v_full[vs], (j / FA_VS) & 1);
        tc_fence_after();
        const uint32_t vbase = smem_u32(sV + vs * FA_KV_BYTES);
        const uint32_t a_tmem = tmem_s + sb * FA_BN;       // P_j: 64 columns of packed bf16 pairs, 8 per K=16 step
#pragma unroll
        for (int s = 0; s < FA_BN / 16; ++s) {
          uint64_t db;
          if (p.v_mode == 0)
            db = umma_desc_mn_sw128(vbase + s * 16 * 128);                       // 16 key rows per k-step
          else
            db = umma_desc_k_sw128(vbase + (s >> 2) * (FA_KV_BYTES / 2)) + 2 * (s & 3);
          umma_bf16_ts(tmem_o + wg * FA_D, a_tmem + 8 * s, db, idesc_pv, (j >= 2 || s > 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[wg]);
        umma_commit(&v_empty[vs]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;                  // TMEM lane quadrant
    const int r = q * 32 + lane;             // row of the query tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_addr = tmem_o + lane_off + wg * FA_D;
    float m_ref = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto softmax_tile = [&](int j, auto rg_tag, auto first_tag) {
      constexpr bool ragged = decltype(rg_tag)::value;
      constexpr bool first = decltype(first_tag)::value;
      const int s3 = j % FA_SB;
      const int kv0 = j * FA_BN;
      if constexpr (!first) {
        // rare: the previous tile of this warpgroup moved the reference max -> rescale the accumulator in TMEM
        if (__any_sync(0xffffffffu, alpha_prev != 1.0f)) {
          mbar_wait(&pv_done[wg], ((j - 2) >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t raw[32];
            tmem_ld_32x32b_x32(o_addr + h * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha_prev);
            tmem_st_32x32b_x32(o_addr + h * 32, raw);
          }
          tmem_st_wait();
        }
      }
      mbar_wait(&s_full[s3], (j / FA_SB) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_s + lane_off + s3 * FA_BN;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if constexpr (first) {
        // row max over the 128 scores, 32 at a time (four independent running maxima)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]), a2 = __uint_as_float(sv[i + 2]),
                  a3 = __uint_as_float(sv[i + 3]);
            if constexpr (ragged) {
              if (kv0 + c * 32 + i >= p.skv) a0 = -INFINITY;
              if (kv0 + c * 32 + i + 1 >= p.skv) a1 = -INFINITY;
              if (kv0 + c * 32 + i + 2 >= p.skv) a2 = -INFINITY;
              if (kv0 + c * 32 + i + 3 >= p.skv) a3 = -INFINITY;
            }
            mx0 = fmaxf(mx0, a0);
            mx1 = fmaxf(mx1, a1);
            mx2 = fmaxf(mx2, a2);
            mx3 = fmaxf(mx3, a3);
          }
        }
        m_ref = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      const float mb = m_ref * p.scale_log2;
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;  // four partial sums: no long dependent FADD chain
      uint64_t rsp0 = 0, rsp1 = 0, rsp2 = 0, rsp3 = 0;   // the same, as packed f32x2 accumulators (full tiles)
      const uint64_t c2 = f2_pack(p.scale_log2, p.scale_log2), nmb2 = f2_pack(-mb, -mb);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, sv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = t * 8 + u * 2;
            if constexpr (ragged) {
              float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]);
              float p0 = ex2_approx(fmaf(a0, p.scale_log2, -mb));
              float p1 = ex2_approx(fmaf(a1, p.scale_log2, -mb));
              if (kv0 + c * 32 + i >= p.skv) { p0 = 0.f; a0 = -INFINITY; }
              if (kv0 + c * 32 + i + 1 >= p.skv) { p1 = 0.f; a1 = -INFINITY; }
              if constexpr (!first) {
                if (u == 0) mx0 = fmaxf(mx0, fmaxf(a0, a1)); else if (u == 1) mx1 = fmaxf(mx1, fmaxf(a0, a1));
                else if (u == 2) mx2 = fmaxf(mx2, fmaxf(a0, a1)); else mx3 = fmaxf(mx3, fmaxf(a0, a1));
              }
              if (u == 0) rs0 += p0 + p1; else if (u == 1) rs1 += p0 + p1; else if (u == 2) rs2 += p0 + p1; else rs3 += p0 + p1;
              pk[t * 4 + u] = pack_bf16(p0, p1);
            } else {
              // packed path: FFMA2 scale/shift, 2 x MUFU (or the FMA-pipe cubic), FADD2 row sum, 3-input max, F2FP
              const uint64_t x2 = f2_fma(f2_pack_bits(sv[i], sv[i + 1]), c2, nmb2);
              float p0, p1;
              if ((POLY >> u) & 1) {
                ex2_poly2(x2, p0, p1);
              } else {
                float x0, x1;
                f2_unpack(x2, x0, x1);
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
              }
              if constexpr (!first) {
                const float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]);
                if (u == 0) mx0 = max3f(mx0, a0, a1); else if (u == 1) mx1 = max3f(mx1, a0, a1);
                else if (u == 2) mx2 = max3f(mx2, a0, a1); else mx3 = max3f(mx3, a0, a1);
              }
              const uint64_t pp = f2_pack(p0, p1);
              if (u == 0) rsp0 = f2_add(rsp0, pp); else if (u == 1) rsp1 = f2_add(rsp1, pp);
              else if (u == 2) rsp2 = f2_add(rsp2, pp); else rsp3 = f2_add(rsp3, pp);
              pk[t * 4 + u] = pack_bf16(p0, p1);
            }
          }
        }
        // keys [32c, 32c+32) -> columns [16c, 16c+16) of the same slot: only score columns already read are overwritten
        tmem_st_32x32b_x16(s_addr + c * 16, pk);
      }
      if constexpr (!ragged) {
        float lo, hi;
        f2_unpack(f2_add(f2_add(rsp0, rsp1), f2_add(rsp2, rsp3)), lo, hi);
        rs0 = lo + hi;
      }
      // Observe every pv_done phase once and in order (a parity wait is only meaningful while the barrier is in the
      // awaited phase or the next one): the PV of this warpgroup's previous tile was issued a whole softmax period ago,
      // so this returns at once, and the final wait below can no longer alias onto an older phase (the failure mode
      // seen in flash_attn_fwd_wide_kernel's first GPU run, where one warpgroup follows every tile).
      if constexpr (!first) mbar_wait(&pv_done[wg], ((j - 2) >> 1) & 1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[s3]);
      float alpha = 1.0f;
      if constexpr (!first) {
        const float m_true = fmaxf(fmaxf(m_ref, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
        if ((m_true - m_ref) * p.scale_log2 > FT_RESCALE_LOG2) {
          alpha = ex2_approx((m_ref - m_true) * p.scale_log2);
          m_ref = m_true;
        }
      }
      l_run = (l_run + ((rs0 + rs1) + (rs2 + rs3))) * alpha;
      alpha_prev = alpha;       // applied to the accumulator once this tile's PV has been added to it
    };
    int last = -1;
    for (int j = wg; j < n_tiles; j += 2) {
      const bool rg = j * FA_BN + FA_BN > p.skv;
      if (j == wg) {
        if (rg) softmax_tile(j, std::true_type{}, std::true_type{});
        else softmax_tile(j, std::false_type{}, std::true_type{});
      } else {
        if (rg) softmax_tile(j, std::true_type{}, std::false_type{});
        else softmax_tile(j, std::false_type{}, std::false_type{});
      }
      last = j;
    }
    // ---- this warpgroup's accumulator, once: TMEM -> registers (scaled by the pending alpha of its last tile)
    float o[FA_D];
    if (last >= 0) {
      mbar_wait(&pv_done[wg], (last >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(o_addr + h * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[h * 32 + i] = __uint_as_float(raw[i]) * alpha_prev;
      }
    } else {
#pragma unroll
      for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
    }
    // ---- merge the two partial softmax results: each warpgroup finishes 32 of the 64 output columns
    {
      float* xo = sX + wg * 32 * 128;         // [32 cols][128 rows]
      if (wg == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i * 128 + r] = o[32 + i];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i * 128 + r] = o[i];
      }
      mx_buf[wg * 128 + r] = m_ref;
      mx_buf[256 + wg * 128 + r] = l_run;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the two softmax warpgroups only, once per CTA
    {
      const float m_o = mx_buf[(wg ^ 1) * 128 + r], l_o = mx_buf[256 + (wg ^ 1) * 128 + r];
      const float m_tot = fmaxf(m_ref, m_o);          // finite: tile 0 has at least one valid key
      const float a_me = ex2_approx((m_ref - m_tot) * p.scale_log2);
      const float a_ot = ex2_approx((m_o - m_tot) * p.scale_log2);
      const float l_tot = l_run * a_me + l_o * a_ot;
      const float* xi = sX + (wg ^ 1) * 32 * 128;
      const int row = qt * FA_BM + r;
      const float inv = 1.0f / l_tot;
      float f[32];
      if (wg == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = (o[i] * a_me + xi[i * 128 + r] * a_ot) * inv;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = (o[32 + i] * a_me + xi[i * 128 + r] * a_ot) * inv;
      }
      if (row < p.sq) {
        if (p.lse && wg == 0)
          p.lse[(static_cast<long long>(b) * p.heads + head) * p.sq + row] = m_tot * p.scale_log2 + log2f(l_tot);
        __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.sq + row) * p.ld_out + head * p.d + wg * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          if (wg * 32 + c8 * 8 < p.d) {
            uint4 v;
            v.x = pack_bf16(f[c8 * 8 + 0], f[c8 * 8 + 1]);
            v.y = pack_bf16(f[c8 * 8 + 2], f[c8 * 8 + 3]);
            v.z = pack_bf16(f[c8 * 8 + 4], f[c8 * 8 + 5]);
            v.w = pack_bf16(f[c8 * 8 + 6], f[c8 * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + c8 * 8) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, FA_TMEM_COLS);
}


// ----------------------------------------------------------------------------------------------------------------------
// Head dims 64 < d <= 192: SD1.5 keeps 8 heads at every level (BASELINE configs[2]), so its attention runs at d = 80
// (32x32 latents) and d = 160 (16x16 and 8x8) besides d = 40.  Same TMEM-resident pipeline as the kernel above, cut
// down to what fits: the head dim is handled as DC chunks of 64 columns (TMA zero-fills the columns past d, so every
// operand tile is a proven [128 rows x 128 B] SW128 tile), QK^T accumulates over the chunks' K-steps (only the steps
// that hold real columns are issued), PV is one N = 64 product per chunk into its own 64 accumulator columns, and ONE
// softmax warpgroup walks all KV tiles (2 score/probability slots x 128 columns + DC x 64 output columns of TMEM; the
// O tile of a 160-wide head no longer fits twice).  These levels hold <= 1024 tokens, the kernel is not a hot spot;
// it exists so that no S x S tensor is materialised on the denoising path of any supported UNet.
template <int DC>
struct FwCfg {
  static constexpr int KS = DC == 2 ? 3 : 2;               // K / V ring depths that fit 227 KB
  static constexpr int VS = DC == 2 ? 2 : 1;
  static constexpr int TILE = DC * FA_KV_BYTES;            // one K or V tile: DC chunks of [128 keys x 128 B]
  static constexpr int SMEM = DC * FA_Q_BYTES + (KS + VS) * TILE + 1024 + 512;
};
constexpr int FW_THREADS = 256;                            // 4 control warps + 1 softmax warpgroup
constexpr int FW_SB = 2;                                   // score / probability slots

template <int DC>
__global__ void __launch_bounds__(FW_THREADS, 1) flash_attn_fwd_wide_kernel(const __grid_constant__ FlashParams p) {
  using C = FwCfg<DC>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + DC * FA_Q_BYTES;
  uint8_t* sV = sK + C::KS * C::TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::VS * C::TILE);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [KS]
  uint64_t* k_empty = k_full + C::KS;      // [KS]
  uint64_t* v_full = k_empty + C::KS;      // [VS]
  uint64_t* v_empty = v_full + C::VS;      // [VS]
  uint64_t* s_full = v_empty + C::VS;      // [FW_SB]  QK^T of the slot finished
  uint64_t* p_full = s_full + FW_SB;       // [FW_SB]  P written into the slot (and the accumulator rescaled if needed)
  uint64_t* pv_done = p_full + FW_SB;      // [1]      the latest PV finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < C::KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < C::VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < FW_SB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_s = tmem_base;                    // + (j & 1) * 128
  const uint32_t tmem_o = tmem_base + FW_SB * FA_BN;    // + chunk * 64

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, DC * FA_Q_BYTES);
#pragma unroll
      for (int c = 0; c < DC; ++c) tma_load_4d(sQ + c * FA_Q_BYTES, &p.tm_q, q_full, c * 64, qt * FA_BM, head, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % C::KS, vs = j % C::VS;
        mbar_wait(&k_empty[ks], ((j / C::KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], C::TILE);
#pragma unroll
        for (int c = 0; c < DC; ++c)
          tma_load_4d(sK + ks * C::TILE + c * FA_KV_BYTES, &p.tm_k, &k_full[ks], c * 64, j * FA_BN, head, b);
        mbar_wait(&v_empty[vs], ((j / C::VS) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], C::TILE);
#pragma unroll
        for (int c = 0; c < DC; ++c)
          tma_load_4d(sV + vs * C::TILE + c * FA_KV_BYTES, &p.tm_v, &v_full[vs], c * 64, j * FA_BN, head, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(FA_BN, false);
      const uint32_t idesc_pv = umma_idesc_bf16(64, true);
      // QK^T of tile j into slot j & 1, accumulated over the head-dim chunks.  The slot last held P_{j-2}: PV_{j-2}
      // was issued before this call and the tensor pipe executes in issue order; the softmax warps left the slot
      // before that PV was issued (p_full).
      auto issue_s = [&](int j) {
        const int ks = j % C::KS, sb = j & 1;
        mbar_wait(&k_full[ks], (j / C::KS) & 1);
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ + c * FA_Q_BYTES));
          const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + ks * C::TILE + c * FA_KV_BYTES));
          const int nks = min(4, (p.d - c * 64 + 15) >> 4);      // K-steps of this chunk that hold real columns
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (s < nks) {
              umma_bf16(tmem_s + sb * FA_BN, dq + 2 * s, dk + 2 * s, idesc_s, acc);
              acc = 1;
            }
          }
        }
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[ks]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int vs = j % C::VS, sb = j & 1;
        mbar_wait(&p_full[sb], (j >> 1) & 1);
        mbar_wait(&v_full[vs], (j / C::VS) & 1);
        tc_fence_after();
        const uint32_t vbase = smem_u32(sV + vs * C::TILE);
        const uint32_t a_tmem = tmem_s + sb * FA_BN;       // P_j: 64 columns of packed bf16 pairs, 8 per K=16 step
#pragma unroll
        for (int s = 0; s < FA_BN / 16; ++s) {
#pragma unroll
          for (int c = 0; c < DC; ++c)
            umma_bf16_ts(tmem_o + c * 64, a_tmem + 8 * s, umma_desc_mn_sw128(vbase + c * FA_KV_BYTES + s * 16 * 128),
                         idesc_pv, (j > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(pv_done);
        umma_commit(&v_empty[vs]);
        if (j + 2 < n_tiles) issue_s(j + 2);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output (one warpgroup, all tiles)
    const int q = warp & 3;                  // TMEM lane quadrant
    const int r = q * 32 + lane;             // row of the query tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_addr = tmem_o + lane_off;
    float m_ref = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto softmax_tile = [&](int j, auto rg_tag, auto first_tag) {
      constexpr bool ragged = decltype(rg_tag)::value;
      constexpr bool first = decltype(first_tag)::value;
      const int s2 = j & 1;
      const int kv0 = j * FA_BN;
      if constexpr (!first) {
        // rare: the previous tile moved the reference max -> rescale the accumulator in TMEM once its PV has landed
        if (__any_sync(0xffffffffu, alpha_prev != 1.0f)) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int h = 0; h < 2 * DC; ++h) {
            uint32_t raw[32];
            tmem_ld_32x32b_x32(o_addr + h * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha_prev);
            tmem_st_32x32b_x32(o_addr + h * 32, raw);
          }
          tmem_st_wait();
        }
      }
      mbar_wait(&s_full[s2], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_s + lane_off + s2 * FA_BN;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if constexpr (first) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]), a2 = __uint_as_float(sv[i + 2]),
                  a3 = __uint_as_float(sv[i + 3]);
            if constexpr (ragged) {
              if (kv0 + c * 32 + i >= p.skv) a0 = -INFINITY;
              if (kv0 + c * 32 + i + 1 >= p.skv) a1 = -INFINITY;
              if (kv0 + c * 32 + i + 2 >= p.skv) a2 = -INFINITY;
              if (kv0 + c * 32 + i + 3 >= p.skv) a3 = -INFINITY;
            }
            mx0 = fmaxf(mx0, a0);
            mx1 = fmaxf(mx1, a1);
            mx2 = fmaxf(mx2, a2);
            mx3 = fmaxf(mx3, a3);
          }
        }
        m_ref = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      const float mb = m_ref * p.scale_log2;
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(s_addr + c * 32, sv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = t * 8 + u * 2;
            float a0 = __uint_as_float(sv[i]), a1 = __uint_as_float(sv[i + 1]);
            float p0 = ex2_approx(fmaf(a0, p.scale_log2, -mb));
            float p1 = ex2_approx(fmaf(a1, p.scale_log2, -mb));
            if constexpr (ragged) {
              if (kv0 + c * 32 + i >= p.skv) { p0 = 0.f; a0 = -INFINITY; }
              if (kv0 + c * 32 + i + 1 >= p.skv) { p1 = 0.f; a1 = -INFINITY; }
            }
            if constexpr (!first) {
              if (u == 0) mx0 = max3f(mx0, a0, a1); else if (u == 1) mx1 = max3f(mx1, a0, a1);
              else if (u == 2) mx2 = max3f(mx2, a0, a1); else mx3 = max3f(mx3, a0, a1);
            }
            if (u == 0) rs0 += p0 + p1; else if (u == 1) rs1 += p0 + p1; else if (u == 2) rs2 += p0 + p1; else rs3 += p0 + p1;
            pk[t * 4 + u] = pack_bf16(p0, p1);
          }
        }
        // keys [32c, 32c+32) -> columns [16c, 16c+16) of the same slot: only score columns already read are overwritten
        tmem_st_32x32b_x16(s_addr + c * 16, pk);
      }
      // Keep in step with pv_done: a parity wait is only meaningful while the barrier is in the awaited phase or the
      // one after it, so every phase is observed exactly once and in order.  Phase j-1 is awaited HERE: PV_j cannot
      // be issued before the arrival below (the barrier cannot run ahead), and PV_{j-1} was issued a whole softmax
      // period ago (no stall in steady state).  Skipping phases let the final wait alias onto an older phase and read
      // the accumulator before the last two PV products had landed (first GPU run of this kernel, 8-tile cases).
      if constexpr (!first) mbar_wait(pv_done, (j - 1) & 1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[s2]);
      float alpha = 1.0f;
      if constexpr (!first) {
        const float m_true = fmaxf(fmaxf(m_ref, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
        if ((m_true - m_ref) * p.scale_log2 > FT_RESCALE_LOG2) {
          alpha = ex2_approx((m_ref - m_true) * p.scale_log2);
          m_ref = m_true;
        }
      }
      l_run = (l_run + ((rs0 + rs1) + (rs2 + rs3))) * alpha;
      alpha_prev = alpha;       // applied to the accumulator once this tile's PV has been added to it
    };
    for (int j = 0; j < n_tiles; ++j) {
      const bool rg = j * FA_BN + FA_BN > p.skv;
      if (j == 0) {
        if (rg) softmax_tile(j, std::true_type{}, std::true_type{});
        else softmax_tile(j, std::false_type{}, std::true_type{});
      } else {
        if (rg) softmax_tile(j, std::true_type{}, std::false_type{});
        else softmax_tile(j, std::false_type{}, std::false_type{});
      }
    }
    // ---- the accumulator, once: TMEM -> registers -> bf16 rows (scaled by the pending alpha and 1 / row sum)
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const int row = qt * FA_BM + r;
    const float fin = alpha_prev / l_run;          // l_run already carries alpha_prev
    if (row < p.sq && p.lse)
      p.lse[(static_cast<long long>(b) * p.heads + head) * p.sq + row] = m_ref * p.scale_log2 + log2f(l_run);
    __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.sq + row) * p.ld_out + head * p.d;
#pragma unroll 1
    for (int h = 0; h < 2 * DC; ++h) {
      if (h * 32 >= p.d) break;
      uint32_t raw[32];
      tmem_ld_32x32b_x32(o_addr + h * 32, raw);
      tmem_ld_wait();
      if (row < p.sq) {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          if (h * 32 + c8 * 8 < p.d) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(raw[c8 * 8 + 0]) * fin, __uint_as_float(raw[c8 * 8 + 1]) * fin);
            v.y = pack_bf16(__uint_as_float(raw[c8 * 8 + 2]) * fin, __uint_as_float(raw[c8 * 8 + 3]) * fin);
            v.z = pack_bf16(__uint_as_float(raw[c8 * 8 + 4]) * fin, __uint_as_float(raw[c8 * 8 + 5]) * fin);
            v.w = pack_bf16(__uint_as_float(raw[c8 * 8 + 6]) * fin, __uint_as_float(raw[c8 * 8 + 7]) * fin);
            *reinterpret_cast<uint4*>(dst + h * 32 + c8 * 8) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

}  // namespace leco

using namespace leco;

// q/k/v: [rows, ld] bf16 buffers whose columns [head*d, head*d+d) hold that head (strides in elements).
// v_t != NULL selects v_mode 1: v_t is V^T laid out [batch][heads][d][skv_pad] (skv_pad multiple of 8).
static int flash_fwd_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          const void* v_t, int64_t skv_pad, void* out, int64_t ldo, float* lse, int batch, int heads,
                          int sq, int skv, int d, float scale, void* stream);
extern "C" int leco_flash_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                   const void* v_t, int64_t skv_pad, void* out, int64_t ldo, int batch, int heads,
                                   int sq, int skv, int d, float scale, void* stream) {
  return flash_fwd_impl(q, ldq, k, ldk, v, ldv, v_t, skv_pad, out, ldo, nullptr, batch, heads, sq, skv, d, scale, stream);
}
// same, additionally writing lse[batch][heads][sq] (fp32, log2 domain) for leco_flash_attn_bwd
extern "C" int leco_flash_attn_fwd_lse(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                       void* out, int64_t ldo, float* lse, int batch, int heads, int sq, int skv, int d,
                                       float scale, void* stream) {
  LECO_REQUIRE(lse, "leco_flash_attn_fwd_lse: null lse");
  return flash_fwd_impl(q, ldq, k, ldk, v, ldv, nullptr, 0, out, ldo, lse, batch, heads, sq, skv, d, scale, stream);
}
static int flash_fwd_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          const void* v_t, int64_t skv_pad, void* out, int64_t ldo, float* lse, int batch, int heads,
                          int sq, int skv, int d, float scale, void* stream) {
  LECO_REQUIRE(q && k && out && (v || v_t), "leco_flash_attn_fwd: null pointer");
  LECO_REQUIRE(d % 8 == 0 && d > 0 && d <= 192, "leco_flash_attn_fwd: head dim %d unsupported (<=192, multiple of 8)", d);
  LECO_REQUIRE(d <= FA_D || !v_t, "leco_flash_attn_fwd: head dims above 64 take V in place (v_t must be NULL)");
  LECO_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldo % 8 == 0, "leco_flash_attn_fwd: strides must be multiples of 8");
  FlashParams p;
  memset(&p, 0, sizeof(p));
  const uint32_t box_q[4] = {FA_D, FA_BM, 1, 1};
  {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)sq, (uint64_t)heads, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)ldq * 2, (uint64_t)d * 2, (uint64_t)ldq * sq * 2};
    if (make_tmap_bf16_4d(&p.tm_q, q, dims, str, box_q)) return -3;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)skv, (uint64_t)heads, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)ldk * 2, (uint64_t)d * 2, (uint64_t)ldk * skv * 2};
    if (make_tmap_bf16_4d(&p.tm_k, k, dims, str, box_q)) return -3;
  }
  if (v_t) {
    LECO_REQUIRE(skv_pad % 8 == 0 && skv_pad >= skv, "leco_flash_attn_fwd: bad skv_pad");
    const uint64_t dims[4] = {(uint64_t)skv_pad, (uint64_t)d, (uint64_t)heads, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)skv_pad * 2, (uint64_t)skv_pad * d * 2, (uint64_t)skv_pad * d * heads * 2};
    const uint32_t box_v[4] = {64, FA_D, 1, 1};
    if (make_tmap_bf16_4d(&p.tm_v, v_t, dims, str, box_v)) return -3;
    p.v_mode = 1;
  } else {
    LECO_REQUIRE(ldv % 8 == 0, "leco_flash_attn_fwd: ldv must be a multiple of 8");
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)skv, (uint64_t)heads, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)ldv * 2, (uint64_t)d * 2, (uint64_t)ldv * skv * 2};
    if (make_tmap_bf16_4d(&p.tm_v, v, dims, str, box_q)) return -3;
    p.v_mode = 0;
  }
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ldo;
  p.sq = sq;
  p.skv = skv;
  p.heads = heads;
  p.d = d;
  p.n_kv_tiles = (skv + FA_BN - 1) / FA_BN;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  static bool attr_set = false;
  if (!attr_set) {
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_ts_kernel<FT_POLY_PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_ts_kernel<0b0000>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_ts_kernel<0b1010>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_ts_kernel<0b1110>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwCfg<2>::SMEM));
    LECO_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_fwd_wide_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwCfg<3>::SMEM));
    attr_set = true;
  }
  // LECO_FLASH_TS=0: the variant that stages P through shared memory and folds PV in registers
  static const bool use_ts = [] { const char* e = getenv("LECO_FLASH_TS"); return !(e && e[0] == '0'); }();
  dim3 grid((sq + FA_BM - 1) / FA_BM, heads, batch);
  count_launch();
  if (d > 128)
    LECO_LAUNCH(flash_attn_fwd_wide_kernel<3>, grid, FW_THREADS, FwCfg<3>::SMEM, reinterpret_cast<cudaStream_t>(stream), p);
  else if (d > FA_D)
    LECO_LAUNCH(flash_attn_fwd_wide_kernel<2>, grid, FW_THREADS, FwCfg<2>::SMEM, reinterpret_cast<cudaStream_t>(stream), p);
  else if (use_ts)
  {
    // LECO_FLASH_POLY = 0 / 1 / 2 / 3: that many quarters of the exponentials on the FMA pipe (A/B switch; default 1/4)
    static const int poly = [] { const char* e = getenv("LECO_FLASH_POLY"); return e ? atoi(e) : 1; }();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (poly == 0) LECO_LAUNCH(flash_attn_fwd_ts_kernel<0b0000>, grid, FA_THREADS, FT_SMEM, st, p);
    else if (poly == 2) LECO_LAUNCH(flash_attn_fwd_ts_kernel<0b1010>, grid, FA_THREADS, FT_SMEM, st, p);
    else if (poly == 3) LECO_LAUNCH(flash_attn_fwd_ts_kernel<0b1110>, grid, FA_THREADS, FT_SMEM, st, p);
    else LECO_LAUNCH(flash_attn_fwd_ts_kernel<FT_POLY_PAIRS>, grid, FA_THREADS, FT_SMEM, st, p);
  }
  else
    LECO_LAUNCH(flash_attn_fwd_kernel, grid, FA_THREADS, FA_SMEM, reinterpret_cast<cudaStream_t>(stream), p);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
