// Shared pieces of the 1-CTA and 2-CTA tcgen05 GEMM kernels: tile constants, the kernel parameter block
// and the TMEM -> registers -> global epilogue (bias / per-sample row bias / residual / GEGLU / bf16|fp32).
#pragma once
#include "../../include/leco_b200.h"
#include "common.cuh"

namespace leco {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;               // 64 bf16 = 128 B = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int GEMM_THREADS = 256;  // 4 role warps + 4 epilogue warps (8 epilogue warps measured no faster: the
                                   // epilogue is instruction-issue bound per scheduler, not latency bound)

struct GemmParams {
  CUtensorMap tm_a, tm_b, tm_a2, tm_b2;
  int mode, M, N;
  int chunks1, ksteps_last1, cin_chunks;
  int has_seg2, ksteps2;
  int tiles_m, tiles_n, batch0, batch1;
  int cn, ch, cw, hb, nb, tiles_per_img, rows_per_tile;
  void* d;
  long long ldd, d_bs0, d_bs1;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* rowbias;
  int rows_per_group;
  long long ld_rowbias;
  const __nv_bfloat16* residual;
  long long ldr;
  int epilogue;
  float alpha;
  int out_fp32;
  unsigned a_tx_bytes;
  // split-K: tile index carries a K-slice id; partial sums are red.add'ed (fp32) into `ws` [M][ldws]
  int dbg;  // perf triage only: 1 = skip the MMAs (TMA-only pipeline), 2 = skip the TMA loads (MMA-only)
  int k_splits;
  float* ws;
  long long ldws;
  // split-K without a finalize launch: per-output-tile arrival counters (persistent, zero, self-resetting); the LAST
  // K-slice CTA to arrive reads the summed tile back, applies the epilogue and re-zeroes the workspace tile
  unsigned int* sk_counters;
  // in-kernel LoRA (lora.py:102-106 in ONE kernel): the stacked lora_down rows `Ad` ride along as fl_kl extra
  // B rows, so accumulator columns [BN, BN+fl_kl) hold T_raw = x.Ad^T; the epilogue adds fl_scale*T_raw.Bup^T
  // epilogue output through TMA (tm_d: D as {N, M, batch0, batch1}, box 32 x 32, 64-byte swizzle): each epilogue warp stages
  // its 32 rows x 32 columns (bf16) in shared memory and issues ONE bulk tensor store per chunk -> full 64-byte row
  // segments instead of one 16-byte store per lane and row.  0 = direct stores (fp32 out, split-K, ragged conv tiles).
  CUtensorMap tm_d;
  int tma_store;
  CUtensorMap tm_ad, tm_bup;      // stacked lora_down rows [fl_kl][K]; stacked lora_up [N][fl_kl] (box BN x 64, zero filled)
  int fl_kl;                      // 0 = off; else 16/32/48/64 padded stacked rank
  int fl_rank;                    // real stacked rank (columns >= fl_rank of T are zero)
  float fl_scale;                 // alpha/rank * multiplier
  const __nv_bfloat16* fl_bup;    // [N][fl_ld_bup] stacked lora_up (block structure)
  long long fl_ld_bup;
  __nv_bfloat16* fl_t_out;        // optional [M][fl_ld_t]: fl_scale*T (saved for the backward)
  long long fl_ld_t;
};
constexpr int FL_MAX_KL = 64;
constexpr int EPI_SLAB_BYTES = 32 * 32 * 2;            // one warp's staging slab: 32 rows x 32 bf16 columns
constexpr int EPI_STAGE_BYTES = 4 * EPI_SLAB_BYTES;    // four epilogue warps

template <int BN, bool FL = false>
struct GemmCfg {
  static constexpr int B_ROWS = BN + (FL ? FL_MAX_KL : 0);  // W rows (+ room for the stacked lora_down rows)
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // as many stages as fit 227 KB: with the role loops at the tensor floor the main loop is bound by bytes in flight
  // in-kernel LoRA: T = fl_scale * x.Ad^T is re-staged by the epilogue warps as a bf16 A-operand tile (128 rows x 128 B,
  // K-major SW128) and the lora_up rows of the N tile arrive by TMA (BN rows x 128 B, double buffered): the up
  // projection T.Bup^T is then ONE more UMMA K-step on the same accumulator
  static constexpr int T_TILE_BYTES = FL ? BLOCK_M * BLOCK_K * 2 : 0;
  static constexpr int BUP_TILE_BYTES = FL ? BN * BLOCK_K * 2 : 0;
  static constexpr int FL_BYTES = T_TILE_BYTES + 2 * BUP_TILE_BYTES;
  static constexpr int STAGES = FL ? (BN >= 160 ? 3 : (BN >= 128 ? 4 : 5)) : ((BN >= 256) ? 4 : (BN >= 160 ? 6 : (BN >= 128 ? 6 : 9)));
  static_assert(STAGES * STAGE_BYTES + FL_BYTES + EPI_STAGE_BYTES + 1024 + 256 <= 232448, "stage ring exceeds 227 KB");
  // accumulator stage stride in TMEM columns (power of two so a stage never straddles an
  // alignment boundary): 64 / 128 / 256
  static constexpr int ACC_STRIDE = FL ? 256 : ((BN <= 64) ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FL_BYTES + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epi_store_bf16(__nv_bfloat16* dst, const float (&v)[32], int ncols_valid) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g * 8 < ncols_valid) {
      uint4 o;
      o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
      o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
      o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
      o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
      *reinterpret_cast<uint4*>(dst + g * 8) = o;
    }
  }
}
__device__ __forceinline__ void epi_store_f32(float* dst, const float (&v)[32], int ncols_valid) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (g * 4 < ncols_valid) {
      *reinterpret_cast<float4*>(dst + g * 4) =
          make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    }
  }
}
// v[j] += src[j] for 32 bf16 (16-byte vector loads), guarded in groups of 8 columns
__device__ __forceinline__ void epi_add_bf16(float (&v)[32], const __nv_bfloat16* src, int ncols_valid) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g * 8 < ncols_valid) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + g * 8));
      v[g * 8 + 0] += bf16_lo(q.x);
      v[g * 8 + 1] += bf16_hi(q.x);
      v[g * 8 + 2] += bf16_lo(q.y);
      v[g * 8 + 3] += bf16_hi(q.y);
      v[g * 8 + 4] += bf16_lo(q.z);
      v[g * 8 + 5] += bf16_hi(q.z);
      v[g * 8 + 6] += bf16_lo(q.w);
      v[g * 8 + 7] += bf16_hi(q.w);
    }
  }
}

// One warp's 32 rows x 32 columns of bf16 output -> its shared-memory slab (row = 64 B, 16-byte chunk c of row l at
// ((c ^ ((l >> 1) & 3)) << 4): the 64-byte swizzle of tm_d, conflict-free for a warp-wide 16-byte store) -> ONE TMA
// store.  The slab is single-buffered: the lane that issues the stores first waits until its previous bulk group has
// finished reading shared memory.  Rows / columns outside the tensor are clipped by the TMA unit.
__device__ __forceinline__ void epi_store_tma(const GemmParams& p, uint8_t* slab, int lane, const float (&v)[32], int col0,
                                              long long row0, int b0, int b1) {
  if (lane == 0) bulk_wait_group_read<0>();
  __syncwarp();
  uint8_t* rowp = slab + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 o;
    o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
    o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
    o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
    o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
    *reinterpret_cast<uint4*>(rowp + ((g ^ sw) << 4)) = o;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(&p.tm_d, smem_u32(slab), col0, (int)row0, b0, b1);
    bulk_commit_group();
  }
}

// The epilogue's global operands of one 32-column group (bias, per-row-group bias, residual), fetched one group
// AHEAD of the TMEM data they are added to: an epilogue iteration is a latency chain (TMEM load -> global loads ->
// math -> store), and the residual comes from L2/HBM; prefetching overlaps that latency with the previous group.
struct EpiGlobals {
  uint4 b[4], rb[4], rs[4];
};
__device__ __forceinline__ void epi_load_globals(EpiGlobals& G, const __nv_bfloat16* bias, const __nv_bfloat16* rb_row,
                                                 const __nv_bfloat16* res_row, int col0, int nvalid, bool ok,
                                                 bool keep_rs = false) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const bool on = ok && g * 8 < nvalid;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (keep_rs) {  // this call refreshes bias / row bias only
      G.b[g] = (on && bias) ? __ldg(reinterpret_cast<const uint4*>(bias + col0 + g * 8)) : z;
      G.rb[g] = (on && rb_row) ? __ldg(reinterpret_cast<const uint4*>(rb_row + col0 + g * 8)) : z;
    } else {        // this call fetches the residual only
      G.rs[g] = (on && res_row) ? __ldg(reinterpret_cast<const uint4*>(res_row + col0 + g * 8)) : z;
    }
  }
}
__device__ __forceinline__ void epi_add_q(float (&v)[32], const uint4 (&q)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    v[g * 8 + 0] += bf16_lo(q[g].x);
    v[g * 8 + 1] += bf16_hi(q[g].x);
    v[g * 8 + 2] += bf16_lo(q[g].y);
    v[g * 8 + 3] += bf16_hi(q[g].y);
    v[g * 8 + 4] += bf16_lo(q[g].z);
    v[g * 8 + 5] += bf16_hi(q[g].z);
    v[g * 8 + 6] += bf16_lo(q[g].w);
    v[g * 8 + 7] += bf16_hi(q[g].w);
  }
}

// Output row of tile row r: (global row m, row inside the problem?)
__device__ __forceinline__ void gemm_epi_row(const GemmParams& p, int r, int mt, long long& m, bool& row_ok) {
  if (p.mode == 0) {
    m = static_cast<long long>(mt) * BLOCK_M + r;
    row_ok = m < p.M;
  } else {
    int img_n0 = 0, img_h0 = 0;
    if (p.nb > 1) {
      img_n0 = mt * p.nb;
      row_ok = (r < p.rows_per_tile) && (img_n0 + r / (p.ch * p.cw) < p.cn);
    } else {
      img_n0 = mt / p.tiles_per_img;
      img_h0 = (mt - img_n0 * p.tiles_per_img) * p.hb;
      row_ok = (r < p.rows_per_tile) && (img_h0 + r / p.cw < p.ch);
    }
    m = static_cast<long long>(img_n0 * p.ch + img_h0) * p.cw + r;
  }
}

// In-kernel LoRA, epilogue half: this thread's row of T_raw = x.Ad^T (accumulator columns [BN, BN+fl_kl), fp32) is
// scaled, rounded to bf16 (the same rounding point as the separate T GEMM) and written into the K-major SW128 smem
// tile the up-projection UMMA reads as its A operand: row r = 128 B, 16-byte chunk c at ((c ^ (r & 7)) << 4).
// With fl_t_out the row is also saved to global memory (the backward needs T: dB = dY^T T).
template <int BN>
__device__ __forceinline__ void gemm_fl_stage_t(const GemmParams& p, uint32_t trow, int r, int mt, int nt,
                                                uint8_t* s_t) {
  long long m;
  bool row_ok;
  gemm_epi_row(p, r, mt, m, row_ok);
  const bool save = p.fl_t_out && nt == 0 && row_ok;
  for (int g = 0; g < p.fl_kl / 8; ++g) {
    uint32_t traw[8];
    tmem_ld_32x32b_x8(trow + BN + g * 8, traw);
    tmem_ld_wait();
    uint4 o;
    o.x = pack_bf16(__uint_as_float(traw[0]) * p.fl_scale, __uint_as_float(traw[1]) * p.fl_scale);
    o.y = pack_bf16(__uint_as_float(traw[2]) * p.fl_scale, __uint_as_float(traw[3]) * p.fl_scale);
    o.z = pack_bf16(__uint_as_float(traw[4]) * p.fl_scale, __uint_as_float(traw[5]) * p.fl_scale);
    o.w = pack_bf16(__uint_as_float(traw[6]) * p.fl_scale, __uint_as_float(traw[7]) * p.fl_scale);
    *reinterpret_cast<uint4*>(s_t + r * 128 + ((g ^ (r & 7)) << 4)) = o;
    if (save) *reinterpret_cast<uint4*>(p.fl_t_out + m * p.fl_ld_t + g * 8) = o;
  }
}

// One output tile: TMEM accumulator (this thread's row r of lane quadrant q) -> epilogue -> global.
// `trow` = TMEM address of the row's first accumulator column; (mt, nt, b0, b1) identify the tile.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, uint32_t trow, int r, int mt, int nt, int b0,
                                                   int b1, uint8_t* slab, uint32_t* sk_ticket = nullptr) {
  const bool geglu = (p.epilogue == 1);
  const int n0 = nt * BN;
  long long m;
  bool row_ok;
  gemm_epi_row(p, r, mt, m, row_ok);
  const long long boff = b0 * p.d_bs0 + b1 * p.d_bs1;
  const int lane = r & 31;
  const long long m_warp = m - lane;      // first output row of this warp's 32-row slab
  const bool tma_out = p.tma_store != 0;  // kernel-uniform

  if (p.k_splits > 1) {
    // partial accumulator of one K-slice -> fp32 workspace (bias / residual are applied by the finalize kernel)
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(trow + c * 32, raw);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      const int nvalid = p.N - col0;
      const bool ok = row_ok && nvalid > 0;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * p.alpha;
      if (ok) {
        float* dst = p.ws + m * p.ldws + col0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g * 4 < nvalid)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(v[g * 4 + 0]),
                         "f"(v[g * 4 + 1]), "f"(v[g * 4 + 2]), "f"(v[g * 4 + 3])
                         : "memory");
        }
      }
    }
    if (p.sk_counters) {
      // ---- in-kernel finalize: the four epilogue warps of this CTA meet, one thread takes a ticket for the output
      // tile; the CTA holding the last ticket sees every K-slice's contribution (fence / atomic / fence) and finishes
      __threadfence();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (r == 0) *sk_ticket = atomicAdd(&p.sk_counters[nt * p.tiles_m + mt], 1u);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const bool last = *sk_ticket == (unsigned)p.k_splits - 1u;
      asm volatile("bar.sync 2, 128;" ::: "memory");   // everyone has read the ticket before it can be rewritten
      if (last) {
        __threadfence();
        if (r == 0) p.sk_counters[nt * p.tiles_m + mt] = 0;
        const __nv_bfloat16* rb_row = p.rowbias ? p.rowbias + (m / p.rows_per_group) * p.ld_rowbias : nullptr;
        const __nv_bfloat16* res_row = p.residual ? p.residual + m * p.ldr : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          const int nvalid = p.N - col0;
          const bool ok = row_ok && nvalid > 0;
          float v[32];
          if (ok) {
            float* src = p.ws + m * p.ldws + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (g * 4 < nvalid) {
                q4 = __ldcg(reinterpret_cast<const float4*>(src + g * 4));
                __stcg(reinterpret_cast<float4*>(src + g * 4), make_float4(0.f, 0.f, 0.f, 0.f));   // leave it zeroed
              }
              v[g * 4 + 0] = q4.x;
              v[g * 4 + 1] = q4.y;
              v[g * 4 + 2] = q4.z;
              v[g * 4 + 3] = q4.w;
            }
            if (p.bias) epi_add_bf16(v, p.bias + col0, nvalid);
            if (rb_row) epi_add_bf16(v, rb_row + col0, nvalid);
            if (res_row) epi_add_bf16(v, res_row + col0, nvalid);
          }
          if (tma_out) {
            if (nvalid > 0) epi_store_tma(p, slab, lane, v, col0, m_warp, b0, b1);
          } else if (ok) {
            epi_store_bf16(reinterpret_cast<__nv_bfloat16*>(p.d) + m * p.ldd + col0, v, nvalid);
          }
        }
      }
    }
  } else if (!geglu) {
    const __nv_bfloat16* rb_row = p.rowbias ? p.rowbias + (m / p.rows_per_group) * p.ld_rowbias : nullptr;
    const __nv_bfloat16* res_row = p.residual ? p.residual + boff + m * p.ldr : nullptr;
    // the residual (L2 / HBM latency) is fetched one group ahead; bias and row bias (L1-resident, shared by all
    // rows) are fetched at the top of the iteration, before the wait on the TMEM load
    EpiGlobals cur, nxt;
    epi_load_globals(cur, nullptr, nullptr, res_row, n0, p.N - n0, row_ok);
#pragma unroll 1  // (fully unrolled this loop bloats the 11 kernel instantiations by 50% and measured slower)
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(trow + c * 32, raw);
      const int col0 = n0 + c * 32;
      const int nvalid = p.N - col0;  // may be <= 0 or > 32
      const bool ok = row_ok && nvalid > 0;
      epi_load_globals(cur, p.bias, rb_row, nullptr, col0, nvalid, row_ok, /*keep_rs=*/true);
      if (c + 1 < BN / 32) epi_load_globals(nxt, nullptr, nullptr, res_row, col0 + 32, nvalid - 32, row_ok);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * p.alpha;
      if (ok) {
        if (p.bias) epi_add_q(v, cur.b);
        if (p.rowbias) epi_add_q(v, cur.rb);
        if (p.residual) epi_add_q(v, cur.rs);
      }
      if (tma_out) {
        if (nvalid > 0) epi_store_tma(p, slab, lane, v, col0, m_warp, b0, b1);   // warp-uniform condition
      } else if (ok) {
        if (p.out_fp32)
          epi_store_f32(reinterpret_cast<float*>(p.d) + boff + m * p.ldd + col0, v, nvalid);
        else
          epi_store_bf16(reinterpret_cast<__nv_bfloat16*>(p.d) + boff + m * p.ldd + col0, v, nvalid);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) cur.rs[g] = nxt.rs[g];
    }
  } else {
    // tile columns [0,BN/2) = hidden block, [BN/2,BN) = matching gate block
#pragma unroll 1
    for (int c = 0; c < BN / 64; ++c) {
      uint32_t rh[32], rg[32];
      tmem_ld_32x32b_x32(trow + c * 32, rh);
      tmem_ld_32x32b_x32(trow + BN / 2 + c * 32, rg);
      tmem_ld_wait();
      const int ocol0 = nt * (BN / 2) + c * 32;
      const int nvalid = p.N / 2 - ocol0;
      const bool ok = row_ok && nvalid > 0;
      float h[32], g[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        h[j] = __uint_as_float(rh[j]) * p.alpha;
        g[j] = __uint_as_float(rg[j]) * p.alpha;
      }
      if (ok) {
        if (p.bias) {
          epi_add_bf16(h, p.bias + ocol0, 32);
          epi_add_bf16(g, p.bias + p.N / 2 + ocol0, 32);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = h[j] * gelu_erf_fast(g[j]);
        if (!tma_out) epi_store_bf16(reinterpret_cast<__nv_bfloat16*>(p.d) + boff + m * p.ldd + ocol0, h, nvalid);
      }
      if (tma_out && nvalid > 0) epi_store_tma(p, slab, lane, h, ocol0, m_warp, b0, b1);
    }
  }
}

// (m0, image n0, image h0) of M-tile `mt`: matrix mode rows mt*128..; conv mode = Hb full image rows of
// Nb images (see leco_gemm_bf16).
__device__ __forceinline__ void gemm_tile_origin(const GemmParams& p, int mt, int& m0, int& img_n0, int& img_h0) {
  m0 = mt * BLOCK_M;
  img_n0 = 0;
  img_h0 = 0;
  if (p.mode == 1) {
    if (p.nb > 1) {
      img_n0 = mt * p.nb;
    } else {
      img_n0 = mt / p.tiles_per_img;
      img_h0 = (mt - img_n0 * p.tiles_per_img) * p.hb;
    }
    m0 = (img_n0 * p.ch + img_h0) * p.cw;
  }
}

int launch_gemm_2cta(const GemmParams& p, int bn, cudaStream_t stream);

}  // namespace leco
