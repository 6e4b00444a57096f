// Stand-alone micro-benchmarks that answer the open design questions of DESIGN.md section 8 (perf triage only; nothing
// here is product code and no result is checked).  NOT YET RUN ON A GPU when committed (round-1 GPU budget was spent):
// run it first thing in round 2 under `timeout 120`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leco_b200/csrc \
//        -o tests/gpu_checks/build/micro_probe tests/gpu_checks/micro_probe.cu
//
//   1. tcgen05.ld bandwidth TMEM -> registers per SM for 4 and 8 reading warps (bounds every GEMM epilogue and the
//      flash-attention softmax: S is 64 KB per 128x128 tile)
//   2. ex2.approx throughput per SM for 4 / 8 / 16 warps (flash-attention floor)
//   3. output-tile store patterns for a [16384 x 320] bf16 result: (a) one 16-byte store per lane and row (what the
//      GEMM epilogue does today), (b) staged through shared memory and written as full rows by consecutive lanes
//   4. global -> shared bulk-copy bandwidth per SM as a function of the bytes in flight (the TMA-side bound of the
//      GEMM main loop): cp.async.bulk with 1..8 buffers of 16 / 32 KB from an L2-resident and an HBM-sized source
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "common.cuh"

using namespace leco;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------------ 1. TMEM read bandwidth
// warps 0..nw-1 read their lane quadrant (warp % 4) `iters` times, 32 columns x 32 lanes x 4 B = 4 KB per instruction.
__global__ void __launch_bounds__(256) tmem_read_probe(int nw, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nw) {
    const uint32_t addr = base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(addr + (i & 7) * 32, v);
      tmem_ld_wait();
      acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);  // keep the load alive (fixed indices: no local array)
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;  // never true for uninitialised-but-finite data; defeats dead-code elimination
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

// ------------------------------------------------------------------------------------------------ 2. ex2 throughput
__global__ void ex2_probe(int iters, long long* out, float* sink) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {  // four independent chains: throughput, not latency
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
    x0 -= 1.0f;
    x1 -= 1.0f;
    x2 -= 1.0f;
    x3 -= 1.0f;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (x0 + x1 + x2 + x3 == 12345.678f) sink[0] = x0;
}

// ------------------------------------------------------------------------------------------------ 3. store patterns
constexpr int ST_COLS = 320;  // bf16 columns of the output (640 B per row)
// (a) lane = row (as after a 32x32b TMEM load): every lane writes its row's 320 columns as 40 16-byte stores.
__global__ void __launch_bounds__(128) store_rowwise(uint4* dst, int rows) {
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= rows) return;
  uint4 v = make_uint4(r, r + 1, r + 2, r + 3);
  uint4* p = dst + (size_t)r * (ST_COLS / 8);
#pragma unroll 8
  for (int c = 0; c < ST_COLS / 8; ++c) p[c] = v;
}
// (b) the same tile staged through shared memory: lane = row writes 16 bytes into a padded smem row, then
// consecutive lanes copy consecutive 16-byte pieces of one row (full 128-byte lines, 5 lines per row).
__global__ void __launch_bounds__(128) store_staged(uint4* dst, int rows) {
  constexpr int VPR = ST_COLS / 8;               // 40 vectors per row
  __shared__ uint4 tile[32][VPR + 1];            // one warp's 32 rows at a time, +1 vector of padding
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128;
  for (int w = 0; w < 4; ++w) {                  // the four row groups of the 128-row tile, one warp-sized group each
    if (warp == 0) {
      const int r = r0 + w * 32 + lane;
      const uint4 v = make_uint4(r, r + 1, r + 2, r + 3);
#pragma unroll 8
      for (int c = 0; c < VPR; ++c) tile[lane][c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * VPR; i += 128) {
      const int rr = i / VPR, cc = i - rr * VPR;
      const int r = r0 + w * 32 + rr;
      if (r < rows) dst[(size_t)r * VPR + cc] = tile[rr][cc];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ 4. bulk-copy bandwidth
// One elected lane keeps `nbuf` buffers of `bytes` in flight with cp.async.bulk (1-D TMA) and re-issues a buffer as
// soon as it has landed; every CTA streams its own slice of `src` (slice_bytes, wrapped) `rounds` times.
__global__ void __launch_bounds__(128) bulk_probe(const uint8_t* src, size_t slice_bytes, int nbuf, int bytes, int rounds,
                                                  long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nbuf * bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < nbuf; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint8_t* base = src + (size_t)blockIdx.x * slice_bytes;
  const int per_slice = (int)(slice_bytes / bytes);
  const long long t0 = clock64();
  if (threadIdx.x < 32) {
    if (elect_one()) {
      for (int i = 0; i < rounds + nbuf; ++i) {
        const int b = i % nbuf;
        if (i >= nbuf) mbar_wait(&bars[b], ((i / nbuf) - 1) & 1);  // the copy issued nbuf iterations ago has landed
        if (i < rounds) {
          mbar_arrive_expect_tx(&bars[b], bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(smem + (size_t)b * bytes)),
                       "l"(base + (size_t)(i % per_slice) * bytes), "r"(bytes), "r"(smem_u32(&bars[b]))
                       : "memory");
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d_out;
  float* d_sink;
  CK(cudaMalloc(&d_out, 8));
  CK(cudaMalloc(&d_sink, 4));
  long long cyc = 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  printf("== 1. tcgen05.ld (32x32b.x32) bandwidth per SM\n");
  for (int nw : {1, 4, 8}) {
    const int iters = 4096;
    tmem_read_probe<<<sms, 256>>>(nw, 64, d_out, d_sink);
    tmem_read_probe<<<sms, 256>>>(nw, iters, d_out, d_sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
    printf("   %d warps: %.1f B/clk/SM  (%.0f cycles per 4 KB load instruction per warp)\n", nw,
           (double)nw * iters * 4096.0 / cyc, (double)cyc / iters);
  }

  printf("== 2. ex2.approx throughput per SM\n");
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 8192;
    ex2_probe<<<sms, warps * 32>>>(64, d_out, d_sink);
    ex2_probe<<<sms, warps * 32>>>(iters, d_out, d_sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
    printf("   %2d warps: %.2f ex2/clk/SM\n", warps, (double)warps * 32 * iters * 4.0 / cyc);
  }

  printf("== 3. [16384 x 320] bf16 tile stores (10.5 MB)\n");
  {
    const int rows = 16384;
    uint4* d_dst;
    CK(cudaMalloc(&d_dst, (size_t)rows * ST_COLS * 2));
    for (int variant = 0; variant < 2; ++variant) {
      for (int rep = 0; rep < 3; ++rep) {
        if (variant == 0) store_rowwise<<<rows / 128, 128>>>(d_dst, rows);
        else store_staged<<<rows / 128, 128>>>(d_dst, rows);
      }
      CK(cudaEventRecord(e0));
      for (int rep = 0; rep < 10; ++rep) {
        if (variant == 0) store_rowwise<<<rows / 128, 128>>>(d_dst, rows);
        else store_staged<<<rows / 128, 128>>>(d_dst, rows);
      }
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("   %-28s %.2f us per tile set, %.0f GB/s\n", variant == 0 ? "lane = row, 16 B per store" : "staged, full rows per warp",
             ms * 100.0, rows * ST_COLS * 2.0 / (ms / 10 * 1e-3) / 1e9);
    }
    CK(cudaFree(d_dst));
  }

  printf("== 4. cp.async.bulk global -> shared bandwidth per SM vs bytes in flight\n");
  {
    const size_t big = (size_t)1 << 30;  // 1 GiB source (HBM); the L2 case re-reads a 256 KB slice per CTA (37 MB total)
    uint8_t* d_src;
    CK(cudaMalloc(&d_src, big));
    CK(cudaMemset(d_src, 1, big));
    CK(cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int hbm = 0; hbm < 2; ++hbm) {
      const size_t slice = hbm ? (big / sms) & ~(size_t)0xFFFF : (size_t)256 * 1024;
      for (int bytes : {16384, 32768}) {
        for (int nbuf : {1, 2, 4, 6, 8, 12}) {
          if ((size_t)nbuf * bytes + 1024 + 256 > 227 * 1024) continue;
          const int rounds = 2048;
          const size_t smem_bytes = (size_t)nbuf * bytes + 1024 + 256;
          bulk_probe<<<sms, 128, smem_bytes>>>(d_src, slice, nbuf, bytes, 64, d_out);
          bulk_probe<<<sms, 128, smem_bytes>>>(d_src, slice, nbuf, bytes, rounds, d_out);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
          printf("   %s  %2d x %2d KB in flight: %6.1f B/clk/SM\n", hbm ? "HBM" : "L2 ", nbuf, bytes / 1024,
                 (double)rounds * bytes / cyc);
        }
      }
    }
    CK(cudaFree(d_src));
  }
  return 0;
}
