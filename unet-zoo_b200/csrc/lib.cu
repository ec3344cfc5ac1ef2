// Library plumbing for libunetzoo_b200.so: error text, driver entry point for tensor-map encoding, device info.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_set>
#include "common.cuh"
#include "unetzoo_b200.h"

namespace {
thread_local char g_err[512] = "";
long long g_launches = 0;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int load_encode() {
  if (g_encode) return UZ_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    uz::set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
    return UZ_ERR_DRIVER;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return UZ_OK;
}
}  // namespace

namespace uz {
int g_pdl = [] {
  // Programmatic dependent launch.  Round 1 (profiles/r01_pdl.md): on for every kernel it gained 4 % on a single stream
  // and LOST 3.5 % with the multi-stream overlap the models use (early-scheduled dependents hold SM resources other
  // streams could use).  Round 2, restructured step (PHiSeg-7/5 B=12, ms per step): off 4.32, every kernel 4.26-4.28,
  // light kernels only 4.48, tensor-core kernels only 4.20-4.26 -- their barrier / TMEM / descriptor prologue is what
  // overlaps the predecessor's tail -- hence default 3.  UZ_PDL=0/1/2/3 selects.
  const char* e = getenv("UZ_PDL");
  if (!e) return 3;
  return (e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
}();

// Shared-memory carveout preference applied to EVERY kernel of the library on its first launch (UZ_CARVEOUT = percent of
// the unified L1/shared array, -1 = leave the driver default).  With the default policy every kernel gets the smallest
// carveout that fits it, so a step that alternates 0-KB elementwise kernels with 100-200 KB tensor-core kernels makes the
// SMs re-partition between launches and keeps kernels of different streams from sharing an SM.
int g_carveout = [] {
  const char* e = getenv("UZ_CARVEOUT");
  return e ? atoi(e) : -1;
}();
static std::mutex g_prep_mutex;
static std::unordered_set<const void*> g_prepared;
void prepare_kernel(const void* fn) {
  if (g_carveout < 0) return;
  std::lock_guard<std::mutex> lock(g_prep_mutex);
  if (g_prepared.insert(fn).second) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout) != cudaSuccess)
      (void)cudaGetLastError();
  }
}

void count_launch() { ++g_launches; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, uint32_t swizzle_bytes) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return UZ_ERR_DRIVER;
  }
  return UZ_OK;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}
}  // namespace uz

extern "C" const char* uz_last_error(void) { return g_err; }

extern "C" int uz_abi_version(void) { return UZ_ABI_VERSION; }
extern "C" int uz_storage_dtype(void) {
#ifdef UZ_ACT_FP16
  return 1;
#else
  return 0;
#endif
}

extern "C" int uz_device_sm_count(void) { return uz::num_sms(); }

extern "C" long long uz_launch_count(void) { return g_launches; }

extern "C" int uz_set_smem_carveout(int percent) {
  UZ_CHECK_ARG(percent >= -1 && percent <= 100, "uz_set_smem_carveout: %d outside [-1, 100]", percent);
  std::lock_guard<std::mutex> lock(uz::g_prep_mutex);
  uz::g_carveout = percent;
  uz::g_prepared.clear();
  return UZ_OK;
}

#ifdef UZ_PROFILE_KNOBS
namespace uz { unsigned long long* g_trace = nullptr; }
#endif
extern "C" int uz_set_trace_buffer(void* device_ptr) {
#ifdef UZ_PROFILE_KNOBS
  uz::g_trace = static_cast<unsigned long long*>(device_ptr);
  return UZ_OK;
#else
  UZ_CHECK_ARG(device_ptr == nullptr, "uz_set_trace_buffer: this library was built without -DUZ_PROFILE_KNOBS");
  return UZ_OK;
#endif
}

extern "C" int uz_get_pdl(void) { return uz::g_pdl; }

extern "C" int uz_set_pdl(int enabled) {
  uz::g_pdl = enabled < 0 ? 0 : (enabled > 3 ? 1 : enabled);   // 0 off, 1 all kernels, 2 light / 3 tensor-core kernels only
  return UZ_OK;
}
