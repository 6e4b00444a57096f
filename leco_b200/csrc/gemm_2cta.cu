// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a cluster of two CTAs on one TPC computes a
// 256 x BN output tile.  Each CTA stages ITS OWN 128 A rows and HALF of the B rows (N split); one
// `tcgen05.mma.cta_group::2` issued by the leader CTA reads both CTAs' shared memory, so per-SM smem
// fill per K-chunk drops from (128 + BN) to (128 + BN/2) rows — the 1-CTA kernel is bound by per-SM
// operand ingest (~50 B/clk measured: 1250 TF at 8k^3, BN=256), not by the tensor pipe.
//
// Protocol (per smem stage s, accumulator stage a):
//   both CTAs' TMA  --complete_tx-->  leader.full[s]   (leader arrive.expect_tx of BOTH CTAs' bytes)
//   leader MMA      --tcgen05.commit multicast 0b11--> empty[s] in both CTAs, tmem_full[a] in both CTAs
//   both epilogues  --(remote) arrive-->               leader.tmem_empty[a]   (count 8 = 4 warps x 2 CTAs)
#include "gemm_common.cuh"

namespace leco {
void count_launch();

// ---- cluster / cta_group::2 PTX ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();  // .aligned barrier: the warp must be converged (role branches diverge lane 0)
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit cleared).
__device__ __forceinline__ void tma_load_4d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2, int c3) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta_u32(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                     int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// descriptors by their low words (see umma_bf16_lo in common.cuh)
__device__ __forceinline__ void umma_bf16_2cta_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_SW128_HI)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_mc_u32(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the barrier at the same smem offset in every CTA of `mask` once prior MMAs retire
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
__device__ __host__ __forceinline__ uint32_t umma_idesc_bf16_m256(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

template <int BN>
struct Gemm2Cfg {
  static constexpr int B_STAGE_BYTES = (BN / 2) * BLOCK_K * 2;  // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 6 : (BN >= 160 ? 8 : (BN >= 128 ? 9 : 10));  // fill 227 KB
  static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "stage ring exceeds 227 KB");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;   // 4 x one 32x32 bf16 slab (TMA-store staging)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  used in the leader only
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  per CTA
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]       per CTA
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]       used in the leader only (count 8)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
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
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // barriers of both CTAs exist before any remote signal / allocation handshake
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_m2 = (p.tiles_m + 1) / 2;  // pairs of M tiles
  const int tiles_mn = tiles_m2 * p.tiles_n;
  const int total_tiles = tiles_mn * p.batch0 * p.batch1;
  const int total_chunks = p.chunks1 + p.has_seg2;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const bool geglu = p.epilogue == 1;

  // role loops: same lean single-elected-lane structure as gemm_tcgen05.cu
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; one elected lane each)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx1 = 2 * (p.a_tx_bytes + Cfg::B_STAGE_BYTES);
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int bidx = tile / tiles_mn;
        const int rem = tile - bidx * tiles_mn;
        const int nt = rem / tiles_m2;
        const int mt = (rem - nt * tiles_m2) * 2 + (int)rank;  // may be == tiles_m (dummy tile: all rows OOB)
        const int b1 = bidx / p.batch0;
        const int b0 = bidx - b1 * p.batch0;
        // B rows held by this CTA: its half of the BN-wide N tile (GEGLU: rank 0 = hidden rows, rank 1 = gate rows)
        const int nb_row = geglu ? (int)rank * (p.N / 2) + nt * (BN / 2) : nt * BN + (int)rank * (BN / 2);
        int m0, img_n0, img_h0;
        gemm_tile_origin(p, mt, m0, img_n0, img_h0);
        int cc = 0, kw = 0, kh = 0;  // filter-tap counters of the implicit conv (chunk c = (tap, cc), tap = kh*3 + kw)
        for (int c = 0; c < p.chunks1; ++c) {
          mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
          const uint32_t fb = full0 + stage * 8;
          const uint32_t sa = sa0 + stage * A_STAGE_BYTES;
          const uint32_t sb = sb0 + stage * Cfg::B_STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx_u32(fb, tx1);
          if (p.mode == 0) {
            tma_load_4d_2cta_u32(sa, &p.tm_a, fb, c * BLOCK_K, m0, b0, b1);
          } else {
            tma_load_4d_2cta_u32(sa, &p.tm_a, fb, cc * BLOCK_K, kw - 1, img_h0 + kh - 1, img_n0);
          }
          tma_load_4d_2cta_u32(sb, &p.tm_b, fb, c * BLOCK_K, nb_row, b0, b1);
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
        if (p.has_seg2) {
          mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
          const uint32_t fb = full0 + stage * 8;
          if (leader) mbar_arrive_expect_tx_u32(fb, 2 * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
          tma_load_4d_2cta_u32(sa0 + stage * A_STAGE_BYTES, &p.tm_a2, fb, 0, m0, 0, 0);
          tma_load_4d_2cta_u32(sb0 + stage * Cfg::B_STAGE_BYTES, &p.tm_b2, fb, 0, nb_row, 0, 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // -------------------------------------------------------------- MMA issuer (leader CTA only; one elected lane)
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16_m256(BN);
      const uint32_t a_lo0 = umma_desc_lo(sa0), b_lo0 = umma_desc_lo(sb0);
      const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
      const int c_fast = (p.ksteps_last1 != 4) ? p.chunks1 - 1 : p.chunks1;  // full 4-k-step chunks
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_u32(tempty0 + as * 8, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
        uint32_t acc = 0;
        for (int c = 0; c < c_fast; ++c) {
          mbar_wait_u32(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (Cfg::B_STAGE_BYTES >> 4);
          umma_bf16_2cta_lo(tmem_d, a_lo, b_lo, idesc, acc);
          umma_bf16_2cta_lo(tmem_d, a_lo + 2, b_lo + 2, idesc, 1u);
          umma_bf16_2cta_lo(tmem_d, a_lo + 4, b_lo + 4, idesc, 1u);
          umma_bf16_2cta_lo(tmem_d, a_lo + 6, b_lo + 6, idesc, 1u);
          umma_commit_2cta_mc_u32(empty0 + stage * 8, 0b11);
          acc = 1;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        for (int c = c_fast; c < total_chunks; ++c) {
          mbar_wait_u32(full0 + stage * 8, phase);
          tc_fence_after();
          const int ksteps = (c < p.chunks1 - 1) ? 4 : (c == p.chunks1 - 1 ? p.ksteps_last1 : p.ksteps2);
          const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (Cfg::B_STAGE_BYTES >> 4);
          for (int j = 0; j < ksteps; ++j) umma_bf16_2cta_lo(tmem_d, a_lo + 2 * j, b_lo + 2 * j, idesc, (acc | j) != 0 ? 1u : 0u);
          umma_commit_2cta_mc_u32(empty0 + stage * 8, 0b11);
          acc = 1;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta_mc_u32(tfull0 + as * 8, 0b11);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int bidx = tile / tiles_mn;
      const int rem = tile - bidx * tiles_mn;
      const int nt = rem / tiles_m2;
      const int mt = (rem - nt * tiles_m2) * 2 + (int)rank;
      const int b1 = bidx / p.batch0;
      const int b0 = bidx - b1 * p.batch0;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
      if (mt < p.tiles_m) {
        gemm_epilogue_tile<BN>(p, trow, r, mt, nt, b0, b1, smem_epi + q * EPI_SLAB_BYTES);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[as], 0);
    }
    if (lane == 0) bulk_wait_group_read<0>();   // the last TMA stores must have read their slab before smem is released
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
}

template <int BN>
static int launch_2cta(const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    LECO_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const long long pair_tiles = 1LL * ((p.tiles_m + 1) / 2) * p.tiles_n * p.batch0 * p.batch1;
  const int max_clusters = sm_count() / 2;
  const int clusters = (int)(pair_tiles < max_clusters ? pair_tiles : max_clusters);
  LECO_LAUNCH(gemm_tcgen05_2cta_kernel<BN>, 2 * clusters, GEMM_THREADS, Cfg::SMEM_BYTES, stream, p);
  LECO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_gemm_2cta(const GemmParams& p, int bn, cudaStream_t stream) {
  count_launch();
  switch (bn) {
    case 64: return launch_2cta<64>(p, stream);
    case 128: return launch_2cta<128>(p, stream);
    case 160: return launch_2cta<160>(p, stream);
    default: return launch_2cta<256>(p, stream);
  }
}

}  // namespace leco
