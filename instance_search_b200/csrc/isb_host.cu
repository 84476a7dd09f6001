// Error state, device checks and TMA tensor-map construction.
#include "isb_host.cuh"

#include <atomic>
#include <mutex>
#include <string.h>

namespace isb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Tuning options (include/isb.h, isb_set_option): explicit, process-wide, -1 = default.
static std::atomic<int> g_options[ISB_OPT_COUNT_];

int option(int id, int dflt) {
  if (id < 0 || id >= ISB_OPT_COUNT_) return dflt;
  const int v = g_options[id].load(std::memory_order_relaxed);
  return v == 0 ? dflt : v - 1;   // stored biased by one so that zero-initialised means "unset"
}

int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    return 148;
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_k64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ISB_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%llu)",
              (unsigned long long)ld);
    return ISB_ERR_INVALID_ARGUMENT;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                        box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return ISB_ERR_CUDA;
  }
  return ISB_OK;
}

}  // namespace isb

extern "C" int isb_abi_version(void) { return ISB_ABI_VERSION; }

extern "C" int isb_set_option(int option_id, int value) {
  ISB_CHECK_ARG(option_id >= 0 && option_id < ISB_OPT_COUNT_, "isb_set_option: unknown option %d", option_id);
  ISB_CHECK_ARG(value >= -1, "isb_set_option: value must be >= 0, or -1 to restore the default");
  isb::g_options[option_id].store(value + 1, std::memory_order_relaxed);
  return ISB_OK;
}

extern "C" int isb_get_option(int option_id) {
  if (option_id < 0 || option_id >= ISB_OPT_COUNT_) return -1;
  return isb::option(option_id, -1);
}

extern "C" const char* isb_last_error(void) { return isb::g_err; }

extern "C" int isb_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    isb::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return ISB_ERR_UNSUPPORTED_DEVICE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    isb::set_error("libisb is built for sm_100a only; device %d has compute capability major %d", dev, major);
    return ISB_ERR_UNSUPPORTED_DEVICE;
  }
  return ISB_OK;
}
