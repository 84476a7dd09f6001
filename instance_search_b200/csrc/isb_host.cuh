// Host-side helpers shared by the C-ABI entry points: error reporting,
// tensor-map construction (driver entry point fetched through the runtime, so
// the library does not link libcuda), launch checks.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/isb.h"

namespace isb {

void set_error(const char* fmt, ...);
int device_sm_count();
// value of a tuning option (ISB_OPT_*), dflt when it was never set / reset with -1
int option(int id, int dflt);

#define ISB_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::isb::set_error(__VA_ARGS__);        \
      return ISB_ERR_INVALID_ARGUMENT;      \
    }                                       \
  } while (0)

#define ISB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ::isb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                       __LINE__);                                                        \
      return ISB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// 2-D bf16 row-major [rows, cols] (leading dimension ld elements) -> tensor map
// with a {64, box_rows} box and 128-byte swizzle.  Out-of-bounds box elements
// are zero-filled by TMA, which is what makes ragged M / N / K safe.
int make_tmap_bf16_k64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t ld, uint32_t box_rows);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace isb
