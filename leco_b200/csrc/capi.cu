// Error plumbing, device info and the driver-API tensor-map encoder for the C ABI.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/leco_b200.h"
#include "common.cuh"

namespace leco {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<int> g_deterministic{-1};
bool deterministic() {
  int v = g_deterministic.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("LECO_DETERMINISTIC");
    v = (e && e[0] == '1') ? 1 : 0;
    g_deterministic.store(v);
  }
  return v == 1;
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LECO_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// cuTensorMapEncodeTiled is a driver-API symbol; resolve it at run time so that the
// shared object loads on a machine without libcuda (the CPU-only build container).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

static int make_tmap_impl(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                          const uint32_t box[4], CUtensorMapSwizzle swz);
int make_tmap_bf16_4d(CUtensorMap* out, const void* base, const uint64_t dims[4],
                      const uint64_t strides_bytes[3], const uint32_t box[4]) {
  return make_tmap_impl(out, base, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
int make_tmap_bf16_4d_sw64(CUtensorMap* out, const void* base, const uint64_t dims[4],
                           const uint64_t strides_bytes[3], const uint32_t box[4]) {
  return make_tmap_impl(out, base, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_64B);
}
static int make_tmap_impl(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                          const uint32_t box[4], CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tensor map base %p is not 16-byte aligned", base);
    return -1;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < 3; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0) {
      set_error("tensor map stride[%d]=%llu not a multiple of 16 bytes", i, (unsigned long long)gstr[i]);
      return -1;
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims {%llu,%llu,%llu,%llu} strides {%llu,%llu,%llu} box {%u,%u,%u,%u}",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
              (unsigned long long)strides_bytes[2], box[0], box[1], box[2], box[3]);
    return -1;
  }
  return 0;
}

}  // namespace leco

extern "C" const char* leco_last_error(void) { return leco::last_error(); }
extern "C" int leco_abi_version(void) { return 2; }
extern "C" int leco_set_deterministic(int on) {
  leco::g_deterministic.store(on ? 1 : 0);
  return 0;
}
extern "C" int64_t leco_launch_count(void) { return leco::g_launches.load(); }
extern "C" int leco_device_info(int32_t* sm, int32_t* major, int32_t* minor) {
  int dev = 0;
  LECO_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LECO_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm) *sm = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return 0;
}
