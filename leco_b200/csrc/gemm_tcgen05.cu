// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D = epi( alpha * ( sum_k A[m,k] B[n,k]  +  sum_k2 A2[m,k2] B2[n,k2] ) )
//
// * operands are bf16, K-major; tiles are staged HBM -> smem by TMA (128B swizzle),
//   multiplied by tcgen05.mma (one elected thread), accumulated in TMEM (fp32) and read
//   back with tcgen05.ld by four epilogue warps while the next tile's MMAs already run
//   (two TMEM accumulator stages).
// * mode 0: plain / batched matrices.  mode 1: implicit-GEMM 3x3 convolution over an NHWC
//   image: the A tile for filter tap (kh,kw) is ONE 4-D TMA box {64 ch, W, Hb, Nb} at
//   coordinates (c0, kw-1, h0+kh-1, n0); out-of-image pixels are zero-filled by TMA, so
//   padding costs nothing and no im2col buffer exists.
// * the optional second K segment is the LoRA residual of lora.py:102-106: A2 is
//   scale*lora_down(x) (K2 <= 64), B2 the lora_up weight: one extra UMMA K-step instead
//   of three extra kernels.
// * epilogue: bias, per-row-group bias (the ResnetBlock2D time-embedding add), residual
//   add, GEGLU, bf16 or fp32 store.
//
// Roles: one elected lane of warp 0 = TMA producer, one elected lane of warp 1 = MMA issuer (each runs its whole
// persistent loop inside a single elect.sync region), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane
// quadrant = warp % 4).
#include <stdio.h>

#include "gemm_common.cuh"

namespace leco {

template <int BN, bool FL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, FL>;
  constexpr int STAGES = Cfg::STAGES;
  pdl_launch_dependents();  // the next kernel may start its prologue while this one runs
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_t = smem + STAGES * Cfg::STAGE_BYTES;                  // FL: T tile (A operand of the up projection)
  uint8_t* smem_bup = smem_t + Cfg::T_TILE_BYTES;                      // FL: 2 x lora_up tile
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES + Cfg::FL_BYTES;  // 4 x one 32x32 bf16 slab (TMA-store staging)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]       MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
  // in-kernel LoRA (FL): T staged (epilogue -> MMA), T consumed (MMA -> epilogue), lora_up tile landed / consumed,
  // accumulator complete INCLUDING the up projection (MMA -> epilogue)
  uint64_t* t_full = bars + 2 * STAGES + 4;     // [1] count 4
  uint64_t* t_empty = t_full + 1;               // [1]
  uint64_t* bup_full = t_full + 2;              // [2]
  uint64_t* bup_empty = t_full + 4;             // [2]
  uint64_t* acc2_full = t_full + 6;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    if (FL) {
      tma_prefetch_desc(&p.tm_ad);
      tma_prefetch_desc(&p.tm_bup);
    }
    if (p.has_seg2) {
      tma_prefetch_desc(&p.tm_a2);
      tma_prefetch_desc(&p.tm_b2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    if (FL) {
      mbar_init(t_full, 4);
      mbar_init(t_empty, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bup_full[i], 1);
        mbar_init(&bup_empty[i], 1);
        mbar_init(&acc2_full[i], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel; from here on global memory is touched

  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int splits = p.k_splits;
  const int n_batches = p.batch0 * p.batch1;
  const int total_tiles = tiles_mn * n_batches * splits;
  // Tile walk.  Plain kernels: tile = blockIdx.x + i*gridDim.x, M fastest (neighbouring CTAs share the B tile in L2).
  // FL (in-kernel LoRA): each CTA owns a CONTIGUOUS run of tiles with N fastest, so consecutive tiles share their M
  // rows and with them T = s*x.A^T: only the first tile of such a run ("fresh") computes the extra accumulator columns
  // and re-stages T; the others just add T.Bup^T from the T tile already in shared memory (no extra MMA columns, no
  // lora_down rows in their TMA traffic, no epilogue -> MMA round trip).  K-slices of a split-K problem each own a
  // partial T, so there every tile is fresh.
  const int tile_begin = FL ? (int)((long long)blockIdx.x * total_tiles / gridDim.x) : (int)blockIdx.x;
  const int tile_end = FL ? (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x) : total_tiles;
  const int tile_step = FL ? 1 : (int)gridDim.x;

  // Both role loops below are latency chains of ONE warp: every instruction per K chunk counts (the tensor
  // pipe needs a 4-MMA chunk every ~220-550 cycles).  So: 32-bit shared addresses computed with one IMAD per
  // stage, incremental filter-tap counters instead of divisions, and a fixed-shape fast path for full chunks.
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
  // Measured with tests/gpu_checks/mma_probe.cu: a role loop that runs ENTIRELY inside one elect.sync region issues
  // a 4-MMA chunk at the tensor floor (320 cycles at N=160); re-electing / re-converging the warp every chunk costs
  // ~700 cycles per chunk.  So each role is one elected lane running its whole persistent loop.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_plain = p.a_tx_bytes + BN * BLOCK_K * 2;
    const int dbg = p.dbg;
    int itp = 0, prev_mt = -1;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++itp) {
      // per-tile index math runs on ONE lane's latency chain: skip every division the common case does not need
      int ks = 0, t2 = tile, bidx = 0, b0 = 0, b1 = 0;
      if (splits > 1) {
        t2 = tile / splits;
        ks = tile - t2 * splits;
      }
      if (n_batches > 1) {
        bidx = t2 / tiles_mn;
        b1 = bidx / p.batch0;
        b0 = bidx - b1 * p.batch0;
      }
      const int rem = t2 - bidx * tiles_mn;
      int nt, mt;
      if (FL) {
        mt = rem / p.tiles_n;
        nt = rem - mt * p.tiles_n;
      } else {
        nt = rem / p.tiles_m;
        mt = rem - nt * p.tiles_m;
      }
      const bool fresh = FL && (splits > 1 || mt != prev_mt);
      prev_mt = mt;
      const uint32_t tx1 = tx_plain + (fresh ? p.fl_kl * BLOCK_K * 2 : 0);
      const int n0 = nt * BN;
      int m0, img_n0, img_h0;
      gemm_tile_origin(p, mt, m0, img_n0, img_h0);
      // this tile's K-slice: segment-1 chunks [c_begin, c_end1) (+ the LoRA segment on the last slice)
      int c_begin = 0, c_end1 = p.chunks1;
      if (splits > 1) {
        c_begin = (int)((long long)ks * p.chunks1 / splits);
        c_end1 = (int)((long long)(ks + 1) * p.chunks1 / splits);
      }
      const bool seg2 = p.has_seg2 && ks == splits - 1;
      if (FL) {
        // stacked lora_up rows of this N tile ([BN] x fl_kl, zero filled to 64 columns) for the up-projection UMMA
        const int bb = itp & 1;
        const uint32_t bf = smem_u32(&bup_full[bb]);
        const uint32_t sbup = smem_u32(smem_bup) + bb * Cfg::BUP_TILE_BYTES;
        mbar_wait(&bup_empty[bb], ((itp >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx_u32(bf, Cfg::BUP_TILE_BYTES);
        if (p.epilogue == 1) {
          tma_load_4d_u32(sbup, &p.tm_bup, bf, 0, nt * (BN / 2), 0, 0);
          tma_load_4d_u32(sbup + (BN / 2) * BLOCK_K * 2, &p.tm_bup, bf, 0, p.N / 2 + nt * (BN / 2), 0, 0);
        } else {
          tma_load_4d_u32(sbup, &p.tm_bup, bf, 0, n0, 0, 0);
        }
      }
      // filter-tap counters of the implicit conv (mode 1): chunk c = (tap, cc), tap = kh*3 + kw
      int cc = 0, kw = 0, kh = 0;
      if (p.mode != 0) {
        const int tap = c_begin / p.cin_chunks;
        cc = c_begin - tap * p.cin_chunks;
        kh = tap / 3;
        kw = tap - kh * 3;
      }
      for (int c = c_begin; c < c_end1; ++c) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
        {
          const uint32_t fb = full0 + stage * 8;
          const uint32_t sa = sa0 + stage * A_STAGE_BYTES;
          const uint32_t sb = sb0 + stage * Cfg::B_STAGE_BYTES;
          if (dbg == 2) {
            mbar_arrive_u32(fb);
          } else {
            mbar_arrive_expect_tx_u32(fb, tx1);
            if (p.mode == 0) {
              tma_load_4d_u32(sa, &p.tm_a, fb, c * BLOCK_K, m0, b0, b1);
            } else {
              tma_load_4d_u32(sa, &p.tm_a, fb, cc * BLOCK_K, kw - 1, img_h0 + kh - 1, img_n0);
            }
            if (p.epilogue == 1) {
              // GEGLU: tile columns [0,BN/2) <- hidden rows, [BN/2,BN) <- the matching gate rows
              tma_load_4d_u32(sb, &p.tm_b, fb, c * BLOCK_K, nt * (BN / 2), b0, b1);
              tma_load_4d_u32(sb + (BN / 2) * BLOCK_K * 2, &p.tm_b, fb, c * BLOCK_K, p.N / 2 + nt * (BN / 2), b0, b1);
            } else {
              tma_load_4d_u32(sb, &p.tm_b, fb, c * BLOCK_K, n0, b0, b1);
            }
            if (FL && fresh)  // stacked lora_down rows for this K chunk land right behind the W rows of the B tile
              tma_load_4d_u32(sb + BN * BLOCK_K * 2, &p.tm_ad, fb, c * BLOCK_K, 0, 0, 0);
          }
        }
        if (++cc == p.cin_chunks) {
          cc = 0;
          if (++kw == 3) {
            kw = 0;
            ++kh;
          }
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (seg2) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
        {
          const uint32_t fb = full0 + stage * 8;
          const uint32_t sa = sa0 + stage * A_STAGE_BYTES;
          const uint32_t sb = sb0 + stage * Cfg::B_STAGE_BYTES;
          if (dbg == 2) {
            mbar_arrive_u32(fb);
          } else {
            mbar_arrive_expect_tx_u32(fb, A_STAGE_BYTES + BN * BLOCK_K * 2);
            tma_load_4d_u32(sa, &p.tm_a2, fb, 0, m0, 0, 0);
            if (p.epilogue == 1) {
              tma_load_4d_u32(sb, &p.tm_b2, fb, 0, nt * (BN / 2), 0, 0);
              tma_load_4d_u32(sb + (BN / 2) * BLOCK_K * 2, &p.tm_b2, fb, 0, p.N / 2 + nt * (BN / 2), 0, 0);
            } else {
              tma_load_4d_u32(sb, &p.tm_b2, fb, 0, n0, 0, 0);
            }
          }
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (one elected lane)
    if (elect_one()) {
    const uint32_t idesc_fresh = umma_idesc_bf16_m128(BN + (FL ? p.fl_kl : 0));  // FL: extra columns = x.Ad^T
    const uint32_t idesc_plain = umma_idesc_bf16_m128(BN);
    const uint32_t a_lo0 = umma_desc_lo(sa0), b_lo0 = umma_desc_lo(sb0);
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    const int dbg = p.dbg;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    // FL: the up projection of tile i (acc += T.Bup^T, one UMMA per 16 ranks) can only be issued once the epilogue
    // warps have re-staged T; it is slipped in after the first chunks of tile i+1's main loop so the tensor pipe
    // never waits for that round trip
    int pend_it = -1, pend_fresh_idx = -1;   // pend_fresh_idx >= 0: the pending tile staged a new T (wait for it)
    int n_fresh = 0, prev_mt = -1;
    const uint32_t idesc_up = umma_idesc_bf16_m128(BN);
    auto issue_up = [&](int it_, int fresh_idx) {
      const int as_ = it_ & 1, bb = it_ & 1;
      if (fresh_idx >= 0) mbar_wait(t_full, fresh_idx & 1);
      mbar_wait(&bup_full[bb], (it_ >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = umma_desc_lo(smem_u32(smem_t));
      const uint32_t tb = umma_desc_lo(smem_u32(smem_bup) + bb * Cfg::BUP_TILE_BYTES);
      const uint32_t d = tmem_base + as_ * Cfg::ACC_STRIDE;
      for (int j = 0; j < p.fl_kl / 16; ++j) umma_bf16_lo(d, ta + 2 * j, tb + 2 * j, idesc_up, 1u);
      umma_commit(&acc2_full[as_]);
      umma_commit(t_empty);
      umma_commit(&bup_empty[bb]);
    };
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      int ks = 0, c_begin = 0, c_end1 = p.chunks1;
      if (splits > 1) {
        ks = tile % splits;
        c_begin = (int)((long long)ks * p.chunks1 / splits);
        c_end1 = (int)((long long)(ks + 1) * p.chunks1 / splits);
      }
      bool fresh = false;
      if (FL) {
        const int t2 = tile / splits;
        const int mt = (t2 % tiles_mn) / p.tiles_n;
        fresh = splits > 1 || mt != prev_mt;
        prev_mt = mt;
      }
      const uint32_t idesc = fresh ? idesc_fresh : idesc_plain;
      const bool seg2 = p.has_seg2 && ks == splits - 1;
      // chunks [c_begin, c_fast) are full 4-k-step chunks; the K tail and the LoRA segment take the generic path
      const int c_fast = (dbg == 1) ? c_begin : ((c_end1 == p.chunks1 && p.ksteps_last1 != 4) ? c_end1 - 1 : c_end1);
      const int c_end = c_end1 + (seg2 ? 1 : 0);
      mbar_wait_u32(tempty0 + as * 8, aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
      uint32_t acc = 0;
      for (int c = c_begin; c < c_fast; ++c) {
        mbar_wait_u32(full0 + stage * 8, phase);
        tc_fence_after();
        {
          // +32 bytes (16 bf16) along K inside the 128B swizzle row: +2 in the >>4 address field
          const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (Cfg::B_STAGE_BYTES >> 4);
          umma_bf16_lo(tmem_d, a_lo, b_lo, idesc, acc);
          umma_bf16_lo(tmem_d, a_lo + 2, b_lo + 2, idesc, 1u);
          umma_bf16_lo(tmem_d, a_lo + 4, b_lo + 4, idesc, 1u);
          umma_bf16_lo(tmem_d, a_lo + 6, b_lo + 6, idesc, 1u);
          umma_commit_u32(empty0 + stage * 8);  // smem slot reusable once these MMAs retire
        }
        acc = 1;
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (FL && pend_it >= 0 && c - c_begin == 2) {
          issue_up(pend_it, pend_fresh_idx);
          pend_it = -1;
        }
      }
      for (int c = c_fast; c < c_end; ++c) {
        mbar_wait_u32(full0 + stage * 8, phase);
        tc_fence_after();
        const int ksteps = (c < p.chunks1 - 1) ? 4 : (c == p.chunks1 - 1 ? p.ksteps_last1 : p.ksteps2);
        {
          if (dbg == 1) {
            mbar_arrive_u32(empty0 + stage * 8);
          } else {
            const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + stage * (Cfg::B_STAGE_BYTES >> 4);
            for (int j = 0; j < ksteps; ++j) umma_bf16_lo(tmem_d, a_lo + 2 * j, b_lo + 2 * j, idesc, (acc | j) != 0 ? 1u : 0u);
            umma_commit_u32(empty0 + stage * 8);
          }
        }
        acc = 1;
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (FL && pend_it >= 0) {   // tile shorter than three chunks
        issue_up(pend_it, pend_fresh_idx);
        pend_it = -1;
      }
      umma_commit_u32(tfull0 + as * 8);  // accumulator complete (FL: up to the up projection) -> epilogue
      if (FL) {
        pend_it = it;
        pend_fresh_idx = fresh ? n_fresh++ : -1;
      }
    }
    if (FL && pend_it >= 0) issue_up(pend_it, pend_fresh_idx);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;                  // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;             // row inside the tile
    int it = 0, prev_mt = -1;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int t2 = tile / splits;
      const int bidx = t2 / tiles_mn;
      const int rem = t2 - bidx * tiles_mn;
      int nt, mt;
      if (FL) {
        mt = rem / p.tiles_n;
        nt = rem - mt * p.tiles_n;
      } else {
        nt = rem / p.tiles_m;
        mt = rem - nt * p.tiles_m;
      }
      const bool fresh = FL && (splits > 1 || mt != prev_mt);
      prev_mt = mt;
      const int b1 = bidx / p.batch0;
      const int b0 = bidx - b1 * p.batch0;
      mbar_wait_backoff(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
      if (FL) {
        if (fresh) {
          mbar_wait(t_empty, (it & 1) ^ 1);          // the previous tile's up projection has consumed the T tile
          gemm_fl_stage_t<BN>(p, trow, r, mt, nt, smem_t);
          tc_fence_before();
          fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(t_full);
        }
        mbar_wait(&acc2_full[as], aphase);           // accumulator now holds x.W^T + T.Bup^T
        tc_fence_after();
      }
      if (p.dbg != 3) gemm_epilogue_tile<BN>(p, trow, r, mt, nt, b0, b1, smem_epi + q * EPI_SLAB_BYTES, tmem_slot + 1);  // dbg 3: no epilogue
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
    if (lane == 0) bulk_wait_group_read<0>();   // the last TMA stores must have read their slab before smem is released
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// split-K finalize: out = bf16(ws + bias + rowbias + residual); the workspace is zeroed again on the way.
// Two vectors per thread with every load issued before the first store (the stores to ws / d could alias the later
// loads as far as the compiler knows), 32-bit index arithmetic: the in-graph timeline showed 6-11 us per call for
// 1-5 MB of workspace, i.e. latency, not bandwidth.
__device__ __forceinline__ void splitk_fin_load(const GemmParams& p, int i, int vpr, int& m, int& col, float4& v, uint2& qr,
                                                uint2& qrb) {
  m = i / vpr;
  col = (i - m * vpr) * 4;
  v = *reinterpret_cast<const float4*>(p.ws + (size_t)m * p.ldws + col);
  qr = make_uint2(0u, 0u);
  qrb = make_uint2(0u, 0u);
  if (p.residual) qr = *reinterpret_cast<const uint2*>(p.residual + (size_t)m * p.ldr + col);
  if (p.rowbias) qrb = *reinterpret_cast<const uint2*>(p.rowbias + (size_t)(m / p.rows_per_group) * p.ld_rowbias + col);
}
__device__ __forceinline__ void splitk_fin_store(const GemmParams& p, int m, int col, const float4& v, const uint2& qr,
                                                 const uint2& qrb) {
  *reinterpret_cast<float4*>(p.ws + (size_t)m * p.ldws + col) = make_float4(0.f, 0.f, 0.f, 0.f);
  float f[4] = {v.x, v.y, v.z, v.w};
  if (p.bias) {
    const uint2 q = *reinterpret_cast<const uint2*>(p.bias + col);
    f[0] += bf16_lo(q.x); f[1] += bf16_hi(q.x); f[2] += bf16_lo(q.y); f[3] += bf16_hi(q.y);
  }
  f[0] = (f[0] + bf16_lo(qrb.x)) + bf16_lo(qr.x);   // same order as the in-kernel epilogue; absent terms were loaded
  f[1] = (f[1] + bf16_hi(qrb.x)) + bf16_hi(qr.x);   // as zeros
  f[2] = (f[2] + bf16_lo(qrb.y)) + bf16_lo(qr.y);
  f[3] = (f[3] + bf16_hi(qrb.y)) + bf16_hi(qr.y);
  uint2 o;
  o.x = pack_bf16(f[0], f[1]);
  o.y = pack_bf16(f[2], f[3]);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d) + (size_t)m * p.ldd + col) = o;
}
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const GemmParams p) {
  pdl_entry();
  const int vpr = p.N / 4;
  const int total = p.M * vpr;                          // < 2^31: checked by the launcher
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < total; i += 2 * stride) {
    int m0, c0, m1, c1;
    float4 v0, v1;
    uint2 r0, r1, b0, b1;
    splitk_fin_load(p, i, vpr, m0, c0, v0, r0, b0);
    splitk_fin_load(p, i + stride, vpr, m1, c1, v1, r1, b1);
    splitk_fin_store(p, m0, c0, v0, r0, b0);
    splitk_fin_store(p, m1, c1, v1, r1, b1);
  }
  if (i < total) {
    int m0, c0;
    float4 v0;
    uint2 r0, b0;
    splitk_fin_load(p, i, vpr, m0, c0, v0, r0, b0);
    splitk_fin_store(p, m0, c0, v0, r0, b0);
  }
}
void count_launch();
static int launch_splitk_finalize(const GemmParams& p, cudaStream_t stream) {
  const long long total = (long long)p.M * (p.N / 4);
  LECO_REQUIRE(total < (1LL << 31), "split-K finalize: %lld vectors", total);
  long long blocks = (total + 511) / 512;               // two vectors per thread
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  count_launch();
  LECO_LAUNCH(splitk_finalize_kernel, (int)blocks, 256, 0, stream, p);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ host side
// Chooses the N tile and the K split together: score = tensor efficiency of the tile shape x how much of the
// machine the (tiles x splits) grid fills.  `max_split` = 1 disables split-K.
static int pick_block_n(const leco_gemm_args& a, int tiles_m, int batches, int nsm, int chunks, int max_split,
                        int* split_out) {
  *split_out = 1;
  if (a.epilogue == 1) return 128;
  const int cand[4] = {256, 160, 128, 64};
  const double rate[4] = {1.0, 0.95, 0.9, 0.6};
  double best = -1;
  int best_bn = 128;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (a.block_n && a.block_n != bn) continue;
    if (a.fl_ad && bn == 256) continue;  // BN + stacked rank must fit one UMMA (N <= 256)
    const int tn = (a.N + bn - 1) / bn;
    const double util_n = double(a.N) / double(tn * bn);  // ragged last tile wastes MMA
    const long long tiles = 1LL * tiles_m * tn * batches;
    int sp = 1;
    if (max_split > 1 && tiles * 2 <= nsm && chunks >= 36) {  // only long-K problems repay the finalize pass
      sp = (int)(nsm / tiles);
      if (sp > max_split) sp = max_split;
      while (sp > 1 && chunks / sp < 8) --sp;
    }
    const long long ctas = tiles * sp;
    const long long waves = (ctas + nsm - 1) / nsm;
    const double eff = double(ctas) / double(waves * nsm);
    const double score = rate[i] * util_n * eff * (sp > 1 ? 0.97 : 1.0);  // split-K pays a finalize pass
    if (score > best + 1e-9) {
      best = score;
      best_bn = bn;
      *split_out = sp;
    }
  }
  return best_bn;
}

template <int BN, bool FL>
static int launch_gemm(const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, FL>;
  static bool attr_set = false;
  if (!attr_set) {
    LECO_CHECK_CUDA(cudaFuncSetAttribute((gemm_tcgen05_kernel<BN, FL>),
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  LECO_LAUNCH((gemm_tcgen05_kernel<BN, FL>), grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream, p);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace leco

extern "C" int leco_gemm_bf16(const leco_gemm_args* a, void* stream_) {
  using namespace leco;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LECO_REQUIRE(a && a->a && a->b && a->d, "leco_gemm_bf16: null operand");
  LECO_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "leco_gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  LECO_REQUIRE(a->N % 8 == 0 && a->K % 8 == 0, "leco_gemm_bf16: N%%8 / K%%8 violated (N=%d K=%d)", a->N, a->K);
  LECO_REQUIRE(a->ldd % 8 == 0 && a->ldb % 8 == 0, "leco_gemm_bf16: ldd/ldb must be multiples of 8");
  const int batch0 = a->batch0 > 0 ? a->batch0 : 1, batch1 = a->batch1 > 0 ? a->batch1 : 1;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.mode = a->mode;
  p.M = a->M;
  p.N = a->N;
  p.batch0 = batch0;
  p.batch1 = batch1;
  p.chunks1 = (a->K + BLOCK_K - 1) / BLOCK_K;
  p.ksteps_last1 = ((a->K - 1) % BLOCK_K) / 16 + 1;  // K%16==8: the tail k-step reads TMA zero fill
  uint32_t box_a[4] = {BLOCK_K, BLOCK_M, 1, 1};
  if (a->mode == 0) {
    LECO_REQUIRE(a->lda % 8 == 0, "leco_gemm_bf16: lda must be a multiple of 8");
    p.tiles_m = (a->M + BLOCK_M - 1) / BLOCK_M;
    const uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)a->M, (uint64_t)batch0, (uint64_t)batch1};
    const uint64_t bs0 = batch0 > 1 ? (uint64_t)a->a_bs0 : (uint64_t)a->lda * a->M;
    const uint64_t bs1 = batch1 > 1 ? (uint64_t)a->a_bs1 : bs0 * batch0;
    const uint64_t str[3] = {(uint64_t)a->lda * 2, bs0 * 2, bs1 * 2};
    if (make_tmap_bf16_4d(&p.tm_a, a->a, dims, str, box_a)) return -3;
    p.a_tx_bytes = A_STAGE_BYTES;
  } else if (a->mode == 1) {
    LECO_REQUIRE(batch0 == 1 && batch1 == 1, "leco_gemm_bf16: conv mode is not batched");
    LECO_REQUIRE(a->cc % 64 == 0 && a->K == 9 * a->cc, "leco_gemm_bf16: conv needs C%%64==0 and K==9C (C=%d K=%d)", a->cc, a->K);
    LECO_REQUIRE(a->cw >= 1 && a->cw <= 128 && a->M == a->cn * a->ch * a->cw, "leco_gemm_bf16: conv geometry (n=%d h=%d w=%d M=%d)", a->cn, a->ch, a->cw, a->M);
    p.cn = a->cn;
    p.ch = a->ch;
    p.cw = a->cw;
    p.cin_chunks = a->cc / 64;
    p.ksteps_last1 = 4;
    int hb = 128 / a->cw;
    if (hb > a->ch) hb = a->ch;
    int nb = 1;
    if (hb == a->ch) {
      nb = 128 / (a->ch * a->cw);
      if (nb < 1) nb = 1;
      if (nb > 256) nb = 256;
    }
    p.hb = hb;
    p.nb = nb;
    p.tiles_per_img = (a->ch + hb - 1) / hb;
    p.rows_per_tile = a->cw * hb * nb;
    p.tiles_m = nb > 1 ? (a->cn + nb - 1) / nb : a->cn * p.tiles_per_img;
    const uint64_t dims[4] = {(uint64_t)a->cc, (uint64_t)a->cw, (uint64_t)a->ch, (uint64_t)a->cn};
    const uint64_t str[3] = {(uint64_t)a->cc * 2, (uint64_t)a->cc * a->cw * 2, (uint64_t)a->cc * a->cw * a->ch * 2};
    box_a[1] = a->cw;
    box_a[2] = hb;
    box_a[3] = nb;
    if (make_tmap_bf16_4d(&p.tm_a, a->a, dims, str, box_a)) return -3;
    p.a_tx_bytes = (unsigned)p.rows_per_tile * BLOCK_K * 2;
  } else {
    LECO_REQUIRE(false, "leco_gemm_bf16: unknown mode %d", a->mode);
  }

  const int nsm = sm_count();
  // cta_pair: 0 = 1-CTA kernel, 1 = 2-CTA kernel, 2 = choose.  Measured (profiles/r1_gemm_shape_modes_2cta.json): the
  // CTA pair shares the B tile, so the TMA-bound long-K GEMMs gain 11-17 % (3x3 convs at the 64x64 / 32x32 levels, ff2);
  // short-K GEMMs (K <= 640) lose to the extra cluster handshakes, and small-M GEMMs are better served by split-K.
  const bool pair = a->cta_pair == 1 ||
                    (a->cta_pair == 2 && p.chunks1 >= 16 && p.tiles_m >= 16 && a->epilogue == 0 && batch0 * batch1 == 1 &&
                     !a->fl_ad && !a->out_fp32);
  const bool splitk_ok = a->splitk_ws && !deterministic() && !pair && !a->out_fp32 && a->epilogue == 0 && batch0 * batch1 == 1 &&
                         (size_t)a->M * a->N * 4 <= (size_t)a->splitk_ws_bytes && a->N % 4 == 0 && !a->fl_t_out;
  int k_split = 1;
  const int bn = pick_block_n(*a, p.tiles_m, batch0 * batch1, nsm, p.chunks1, splitk_ok ? 8 : 1, &k_split);
  LECO_REQUIRE(bn == 64 || bn == 128 || bn == 160 || bn == 256, "leco_gemm_bf16: unsupported block_n %d", bn);
  if (a->epilogue == 1) LECO_REQUIRE(a->N % 128 == 0 && !a->out_fp32 && !a->residual && !a->rowbias, "leco_gemm_bf16: GEGLU needs N%%128==0, bf16 out, no residual/rowbias");
  p.tiles_n = (a->N + bn - 1) / bn;
  {
    const uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)(a->b_rows > 0 ? a->b_rows : a->N), (uint64_t)batch0, (uint64_t)batch1};
    const uint64_t bs0 = batch0 > 1 ? (uint64_t)a->b_bs0 : (uint64_t)a->ldb * a->N;
    const uint64_t bs1 = batch1 > 1 ? (uint64_t)a->b_bs1 : bs0 * batch0;
    const uint64_t str[3] = {(uint64_t)a->ldb * 2, bs0 * 2, bs1 * 2};
    const uint32_t box_b[4] = {BLOCK_K, (uint32_t)((a->epilogue == 1 || pair) ? bn / 2 : bn), 1, 1};
    if (make_tmap_bf16_4d(&p.tm_b, a->b, dims, str, box_b)) return -3;
  }
  if (a->a2) {
    LECO_REQUIRE(a->b2 && a->K2 > 0 && a->K2 <= 64 && a->K2 % 16 == 0, "leco_gemm_bf16: LoRA segment needs K2 in {16,32,48,64} (K2=%d)", a->K2);
    LECO_REQUIRE(a->lda2 % 8 == 0 && a->ldb2 % 8 == 0 && batch0 == 1 && batch1 == 1, "leco_gemm_bf16: LoRA segment strides / batching");
    p.has_seg2 = 1;
    p.ksteps2 = a->K2 / 16;
    const uint64_t dims_a[4] = {(uint64_t)a->K2, (uint64_t)a->M, 1, 1};
    const uint64_t str_a[3] = {(uint64_t)a->lda2 * 2, (uint64_t)a->lda2 * a->M * 2, (uint64_t)a->lda2 * a->M * 2};
    const uint32_t box2[4] = {BLOCK_K, BLOCK_M, 1, 1};
    if (make_tmap_bf16_4d(&p.tm_a2, a->a2, dims_a, str_a, box2)) return -3;
    const uint64_t dims_b[4] = {(uint64_t)a->K2, (uint64_t)a->N, 1, 1};
    const uint64_t str_b[3] = {(uint64_t)a->ldb2 * 2, (uint64_t)a->ldb2 * a->N * 2, (uint64_t)a->ldb2 * a->N * 2};
    const uint32_t box_b2[4] = {BLOCK_K, (uint32_t)((a->epilogue == 1 || pair) ? bn / 2 : bn), 1, 1};
    if (make_tmap_bf16_4d(&p.tm_b2, a->b2, dims_b, str_b, box_b2)) return -3;
  }
  if (a->fl_ad) {
    LECO_REQUIRE(a->fl_bup && a->fl_kl >= 16 && a->fl_kl <= FL_MAX_KL && a->fl_kl % 16 == 0 && a->fl_rank >= 1 &&
                     a->fl_rank <= a->fl_kl,
                 "leco_gemm_bf16: fused LoRA needs kl in {16..64} and 1 <= rank <= kl (kl=%d rank=%d)", a->fl_kl, a->fl_rank);
    LECO_REQUIRE(!a->a2 && !pair && !a->out_fp32 && batch0 * batch1 == 1 && a->fl_ld_ad % 8 == 0 &&
                     a->fl_ld_bup % 8 == 0,
                 "leco_gemm_bf16: fused LoRA excludes the T segment / 2-CTA / fp32 out / batching");
    const uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)a->fl_kl, 1, 1};
    const uint64_t str[3] = {(uint64_t)a->fl_ld_ad * 2, (uint64_t)a->fl_ld_ad * a->fl_kl * 2, (uint64_t)a->fl_ld_ad * a->fl_kl * 2};
    const uint32_t box[4] = {BLOCK_K, (uint32_t)a->fl_kl, 1, 1};
    if (make_tmap_bf16_4d(&p.tm_ad, a->fl_ad, dims, str, box)) return -3;
    {
      // stacked lora_up [N][fl_kl]: one box = the N tile's rows x 64 columns (columns >= fl_kl are TMA zero fill)
      const uint64_t dims_u[4] = {(uint64_t)a->fl_kl, (uint64_t)a->N, 1, 1};
      const uint64_t str_u[3] = {(uint64_t)a->fl_ld_bup * 2, (uint64_t)a->fl_ld_bup * a->N * 2, (uint64_t)a->fl_ld_bup * a->N * 2};
      const uint32_t box_u[4] = {BLOCK_K, (uint32_t)(a->epilogue == 1 ? bn / 2 : bn), 1, 1};
      if (make_tmap_bf16_4d(&p.tm_bup, a->fl_bup, dims_u, str_u, box_u)) return -3;
    }
    p.fl_kl = a->fl_kl;
    p.fl_rank = a->fl_rank;
    p.fl_scale = a->fl_scale;
    p.fl_bup = reinterpret_cast<const __nv_bfloat16*>(a->fl_bup);
    p.fl_ld_bup = a->fl_ld_bup;
    p.fl_t_out = reinterpret_cast<__nv_bfloat16*>(a->fl_t_out);
    p.fl_ld_t = a->fl_ld_t;
    if (a->fl_t_out) LECO_REQUIRE(a->fl_ld_t % 8 == 0, "leco_gemm_bf16: fl_ld_t must be a multiple of 8");
  }
  p.dbg = a->debug_mode;
  p.d = a->d;
  p.ldd = a->ldd;
  p.d_bs0 = batch0 > 1 ? a->d_bs0 : 0;
  p.d_bs1 = batch1 > 1 ? a->d_bs1 : 0;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
  p.rowbias = reinterpret_cast<const __nv_bfloat16*>(a->rowbias);
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.ld_rowbias = a->ld_rowbias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.ldr = a->ldr;
  if (a->residual) LECO_REQUIRE(a->ldr % 8 == 0, "leco_gemm_bf16: ldr must be a multiple of 8");
  if (a->rowbias) LECO_REQUIRE(a->ld_rowbias % 8 == 0, "leco_gemm_bf16: ld_rowbias must be a multiple of 8");
  p.epilogue = a->epilogue;
  p.alpha = a->alpha == 0.0f ? 1.0f : a->alpha;
  p.out_fp32 = a->out_fp32;

  p.k_splits = 1;
  // TMA-store epilogue: bf16 output whose tile rows are 128 consecutive output rows (matrix mode; conv tiles made of
  // whole image rows that tile the image exactly), no split-K
  static const bool tma_store_enabled = [] { const char* e = getenv("LECO_TMA_STORE"); return !(e && e[0] == '0'); }();
  const bool rows_contig = a->mode == 0 || (p.rows_per_tile == BLOCK_M && (p.nb > 1 || a->ch % p.hb == 0));
  // split-K finalize inside the GEMM (last-arriving K-slice CTA; arrival counters in the last 64 KiB of the caller's
  // workspace).  Parity-tested, but MEASURED SLOWER than the separate finalize launch (239.6 vs 223.0 ms / iteration:
  // only `tiles` CTAs of 128 threads finish the tiles, the finalize kernel uses the whole machine), so opt-in only:
  // LECO_SPLITK_FUSED=1
  static const bool sk_fused = [] { const char* e = getenv("LECO_SPLITK_FUSED"); return e && e[0] == '1'; }();
  constexpr long long SK_COUNTER_BYTES = 65536;
  const bool sk_inkernel = sk_fused && k_split > 1 && (long long)p.tiles_m * ((a->N + bn - 1) / bn) <= SK_COUNTER_BYTES / 4 &&
                           (long long)a->M * a->N * 4 <= a->splitk_ws_bytes - SK_COUNTER_BYTES;
  if (tma_store_enabled && !a->out_fp32 && rows_contig && (k_split <= 1 || sk_inkernel) && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0 &&
      (batch0 == 1 || a->d_bs0 % 8 == 0) && (batch1 == 1 || a->d_bs1 % 8 == 0)) {
    const int n_out = a->epilogue == 1 ? a->N / 2 : a->N;
    const uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)a->M, (uint64_t)batch0, (uint64_t)batch1};
    const uint64_t bs0 = batch0 > 1 ? (uint64_t)a->d_bs0 : (uint64_t)a->ldd * a->M;
    const uint64_t bs1 = batch1 > 1 ? (uint64_t)a->d_bs1 : bs0 * batch0;
    const uint64_t str[3] = {(uint64_t)a->ldd * 2, bs0 * 2, bs1 * 2};
    const uint32_t box[4] = {32, 32, 1, 1};
    p.tma_store = make_tmap_bf16_4d_sw64(&p.tm_d, a->d, dims, str, box) == 0 ? 1 : 0;
  }
  if (pair) return launch_gemm_2cta(p, bn, stream);
  long long total_tiles = 1LL * p.tiles_m * p.tiles_n * batch0 * batch1;
  // split-K: few output tiles but a long K (the 16x16 / 8x8 UNet levels at small batch): spread the K range
  // over idle SMs; partial sums meet in an fp32 workspace, splitk_finalize_kernel applies the epilogue
  if (k_split > 1) {
    p.k_splits = k_split;
    p.ws = reinterpret_cast<float*>(a->splitk_ws);
    p.ldws = a->N;
    total_tiles *= k_split;
    if (sk_inkernel)
      p.sk_counters = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(a->splitk_ws) + a->splitk_ws_bytes - SK_COUNTER_BYTES);
  }
  const int grid = (int)(total_tiles < nsm ? total_tiles : nsm);
  count_launch();
  int rc;
  if (p.fl_kl) {
    switch (bn) {
      case 64: rc = launch_gemm<64, true>(p, grid, stream); break;
      case 128: rc = launch_gemm<128, true>(p, grid, stream); break;
      default: rc = launch_gemm<160, true>(p, grid, stream); break;
    }
  } else {
    switch (bn) {
      case 64: rc = launch_gemm<64, false>(p, grid, stream); break;
      case 128: rc = launch_gemm<128, false>(p, grid, stream); break;
      case 160: rc = launch_gemm<160, false>(p, grid, stream); break;
      default: rc = launch_gemm<256, false>(p, grid, stream); break;
    }
  }
  if (rc == 0 && p.k_splits > 1 && !p.sk_counters) rc = launch_splitk_finalize(p, stream);
  return rc;
}
