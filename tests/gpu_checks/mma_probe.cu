// Standalone tcgen05.mma issue-rate probe (perf triage only; no results are checked).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I leco_b200/csrc -o tests/gpu_checks/build/mma_probe \
//        tests/gpu_checks/mma_probe.cu
// One CTA per SM; one elected lane issues `iters` groups of 4 K=16 MMAs back to back on fixed smem
// operands and commits once at the end.  Prints cycles per MMA for several instruction shapes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "common.cuh"

using namespace leco;

// variant: 0 = SS (A, B from smem), 1 = TS (A from TMEM), 2 = SS + commit after every group of 4,
//          3 = SS, every MMA re-reads the SAME K-slice (no descriptor advance), 4 = SS, M = 64
__global__ void __launch_bounds__(128) probe(int n, int variant, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 128 rows x 128 B
  uint8_t* sb = smem + 16384;         // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    uint32_t idesc = umma_idesc_bf16_m128(n);
    if (variant == 4) idesc = (idesc & ~(0x1Fu << 24)) | (4u << 24);
    const uint64_t da = umma_desc_k_sw128(smem_u32(sa));
    const uint64_t db = umma_desc_k_sw128(smem_u32(sb));
    const uint32_t a_tmem = tmem + 256;  // columns 256.. hold a (garbage) A operand for the TS form
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (variant == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                "r"(a_tmem + 8 * j), "l"(db + 2 * j), "r"(idesc), "r"(1u)
                : "memory");
        } else if (variant == 3) {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(tmem, da, db, idesc, 1u);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(tmem, da + 2 * j, db + 2 * j, idesc, 1u);
        }
        if (variant == 2) umma_commit(bar2);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    long long dt = __shfl_sync(0xffffffffu, t1, 0) - 0;  // every lane sees completion at about the same time
    (void)dt;
    if (t0 != 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- probe 2: the GEMM kernel's producer / MMA / epilogue-waiter role skeleton without any data movement ----
// flags: bit0 = two chunks per barrier round (8 MMAs per wait), bit1 = 4 extra warps spin on a barrier that
// completes only at the end (epilogue waiters), bit2 = those waiters sleep between polls, bit3 = single-lane
// (lane == 0) issue instead of elect.sync, bit4 = try_wait of the next stage issued before this chunk's MMAs
template <int STAGES>
__global__ void __launch_bounds__(256) probe_loop(int n, int flags, int chunks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_stage = 16384, b_stage = 256 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (a_stage + b_stage));  // operands: 2 stages reused
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* done_bar = bars + 2 * STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  for (int i = threadIdx.x; i < 2 * (a_stage + b_stage) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const int group = (flags & 1) ? 2 : 1;
  const int rounds = chunks / group;
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < rounds; ++c) {
      mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
      if (elect_one()) mbar_arrive_u32(full0 + stage * 8);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_m128(n);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem)), b_lo0 = umma_desc_lo(smem_u32(smem + 2 * a_stage));
    int stage = 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    bool ready = (flags & 16) ? mbar_try_wait_u32(full0, 0) : false;
    for (int c = 0; c < rounds; ++c) {
      if (flags & 16) {
        if (!ready) mbar_wait_u32(full0 + stage * 8, phase);
        const int ns = (stage + 1 == STAGES) ? 0 : stage + 1;
        ready = mbar_try_wait_u32(full0 + ns * 8, (stage + 1 == STAGES) ? phase ^ 1 : phase);
      } else if (flags & 32) {
        // no wait at all: free-running issue loop
      } else if (flags & 64) {
        mbar_wait_u32(smem_u32(done_bar), 1);  // a wait that always succeeds at once (fresh barrier, parity 1)
      } else if (flags & 256) {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(full0 + stage * 8), "r"(phase) : "memory");
      } else {
        mbar_wait_u32(full0 + stage * 8, phase);
      }
      if (!(flags & 128)) tc_fence_after();
      const bool issuer = (flags & 8) ? (lane == 0) : elect_one();
      if (issuer) {
        const uint32_t a_lo = a_lo0 + (stage & 1) * (a_stage >> 4);
        const uint32_t b_lo = b_lo0 + (stage & 1) * (b_stage >> 4);
        umma_bf16_lo(tmem, a_lo, b_lo, idesc, 1u);
        umma_bf16_lo(tmem, a_lo + 2, b_lo + 2, idesc, 1u);
        umma_bf16_lo(tmem, a_lo + 4, b_lo + 4, idesc, 1u);
        umma_bf16_lo(tmem, a_lo + 6, b_lo + 6, idesc, 1u);
        if (flags & 1) {
          const uint32_t a2 = a_lo0 + ((stage & 1) ^ 1) * (a_stage >> 4), b2 = b_lo0 + ((stage & 1) ^ 1) * (b_stage >> 4);
          umma_bf16_lo(tmem, a2, b2, idesc, 1u);
          umma_bf16_lo(tmem, a2 + 2, b2 + 2, idesc, 1u);
          umma_bf16_lo(tmem, a2 + 4, b2 + 4, idesc, 1u);
          umma_bf16_lo(tmem, a2 + 6, b2 + 6, idesc, 1u);
        }
        umma_commit_u32(empty0 + stage * 8);
      }
      if (flags & 8) __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
    mbar_wait(done_bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 4 && (flags & 2)) {
    if (flags & 4) mbar_wait_backoff(done_bar, 0); else mbar_wait(done_bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---- probe 3: the whole MMA role loop inside ONE elect.sync region (single-thread loop) ----
// variant 0: stage-indexed addresses (IMAD from the stage counter); 1: addresses carried incrementally;
// 2: as 1 plus the stage loop unrolled by STAGES (compile-time stage inside the body)
template <int STAGES, int VAR>
__global__ void __launch_bounds__(256) probe_single(int n, int chunks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t a_stage = 16384, b_stage = 256 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (a_stage + b_stage));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* done_bar = bars + 2 * STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  for (int i = threadIdx.x; i < 2 * (a_stage + b_stage) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < chunks; ++c) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
        mbar_arrive_u32(full0 + stage * 8);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16_m128(n);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem)), b_lo0 = umma_desc_lo(smem_u32(smem + 2 * a_stage));
    const long long t0 = clock64();
    if (elect_one()) {
      if (VAR == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait_u32(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (stage & 1) * (a_stage >> 4);
          const uint32_t b_lo = b_lo0 + (stage & 1) * (b_stage >> 4);
          umma_bf16_lo(tmem, a_lo, b_lo, idesc, 1u);
          umma_bf16_lo(tmem, a_lo + 2, b_lo + 2, idesc, 1u);
          umma_bf16_lo(tmem, a_lo + 4, b_lo + 4, idesc, 1u);
          umma_bf16_lo(tmem, a_lo + 6, b_lo + 6, idesc, 1u);
          umma_commit_u32(empty0 + stage * 8);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      } else if (VAR == 1) {
        int stage = 0;
        uint32_t phase = 0, fb = full0, eb = empty0, a_lo = a_lo0, b_lo = b_lo0;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait_u32(fb, phase);
          tc_fence_after();
          umma_bf16_lo(tmem, a_lo, b_lo, idesc, 1u);
          umma_bf16_lo(tmem, a_lo + 2, b_lo + 2, idesc, 1u);
          umma_bf16_lo(tmem, a_lo + 4, b_lo + 4, idesc, 1u);
          umma_bf16_lo(tmem, a_lo + 6, b_lo + 6, idesc, 1u);
          umma_commit_u32(eb);
          fb += 8; eb += 8;
          a_lo ^= (a_stage >> 4); b_lo ^= (b_stage >> 4);  // 2 operand buffers in this probe
          if (++stage == STAGES) { stage = 0; phase ^= 1; fb = full0; eb = empty0; }
        }
      } else {
        uint32_t phase = 0;
        for (int c = 0; c < chunks; c += STAGES) {
#pragma unroll
          for (int s = 0; s < STAGES; ++s) {
            mbar_wait_u32(full0 + s * 8, phase);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + (s & 1) * (a_stage >> 4);
            const uint32_t b_lo = b_lo0 + (s & 1) * (b_stage >> 4);
            umma_bf16_lo(tmem, a_lo, b_lo, idesc, 1u);
            umma_bf16_lo(tmem, a_lo + 2, b_lo + 2, idesc, 1u);
            umma_bf16_lo(tmem, a_lo + 4, b_lo + 4, idesc, 1u);
            umma_bf16_lo(tmem, a_lo + 6, b_lo + 6, idesc, 1u);
            umma_commit_u32(empty0 + s * 8);
          }
          phase ^= 1;
        }
      }
      umma_commit(done_bar);
    }
    __syncwarp();
    mbar_wait(done_bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int VAR>
static void run_single(const char* name, long long* d_out) {
  const int smem_bytes = 2 * (16384 + 32768) + 256 + 1024;
  cudaFuncSetAttribute(probe_single<6, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int chunks = 8190;  // multiple of 6
  for (int n : {64, 160, 256}) {
    long long cyc = 0;
    for (int rep = 0; rep < 2; ++rep) probe_single<6, VAR><<<148, 256, smem_bytes>>>(n, chunks, d_out);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%s n %d: %s\n", name, n, cudaGetErrorString(err));
      return;
    }
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-26s N=%3d  %7.1f cyc/chunk (4 MMAs; tensor floor %5.1f)\n", name, n, (double)cyc / chunks, 4 * 128.0 * n / 256.0);
  }
}

static void run_loop_probe(long long* d_out) {
  run_single<0>("single-thread stage-indexed", d_out);
  run_single<1>("single-thread incremental", d_out);
  run_single<2>("single-thread unrolled", d_out);
  const int smem_bytes = 2 * (16384 + 32768) + 256 + 1024;
  cudaFuncSetAttribute(probe_loop<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(probe_loop<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int chunks = 8192;
  struct V { const char* name; int flags; int stages; };
  const V vs[] = {{"loop elect", 0, 6},        {"loop lane0", 8, 6},          {"no wait", 32, 6},
                  {"no wait no fence", 32 + 128, 6}, {"always-true wait", 64, 6}, {"always-true, no fence", 64 + 128, 6},
                  {"real wait, no fence", 128, 6}, {"test_wait spin", 256, 6},  {"test_wait, no fence", 256 + 128, 6},
                  {"pair elect", 1, 3},         {"pair no fence", 1 + 128, 3},  {"pair test_wait no fence", 1 + 256 + 128, 3}};
  for (const V& v : vs) {
    for (int n : {64, 160, 256}) {
      long long cyc = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (v.stages == 6) probe_loop<6><<<148, 256, smem_bytes>>>(n, v.flags, chunks, d_out);
        else probe_loop<3><<<148, 256, smem_bytes>>>(n, v.flags, chunks, d_out);
      }
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) {
        printf("%s n %d: %s\n", v.name, n, cudaGetErrorString(err));
        return;
      }
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      printf("%-26s N=%3d  %7.1f cyc/chunk (4 MMAs; tensor floor %5.1f)\n", v.name, n, (double)cyc / chunks, 4 * 128.0 * n / 256.0);
    }
  }
}

int main(int argc, char** argv) {
  if (argc > 1) {
    long long* d;
    cudaMalloc(&d, 8);
    run_loop_probe(d);
    return 0;
  }

  long long* d_out;
  cudaMalloc(&d_out, 8);
  const int smem_bytes = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const char* names[] = {"SS", "TS(A in TMEM)", "SS+commit/4", "SS same K-slice", "SS M=64"};
  const int iters = 4096;
  for (int variant = 0; variant < 5; ++variant) {
    for (int n : {64, 128, 160, 256}) {
      for (int grid : {1, 148}) {
        long long cyc = 0;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        probe<<<grid, 128, smem_bytes>>>(n, variant, 64, d_out);  // warm
        cudaEventRecord(e0);
        probe<<<grid, 128, smem_bytes>>>(n, variant, iters, d_out);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
          printf("variant %d n %d: %s\n", variant, n, cudaGetErrorString(err));
          return 1;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        const double per = (double)cyc / (iters * 4.0);
        const double floor_c = 128.0 * n / 256.0 * (variant == 4 ? 0.5 : 1.0);
        printf("%-16s N=%3d grid=%3d  %7.1f cyc/MMA  (floor %5.1f)  %.3f ms  -> %.0f TFLOP/s chip-equivalent\n",
               names[variant], n, grid, per, floor_c, ms,
               (variant == 4 ? 64.0 : 128.0) * n * 16 * 2 * 4.0 * iters * grid / (ms * 1e-3) / 1e12);
      }
    }
  }
  return 0;
}
