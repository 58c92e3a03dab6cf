// Shared host/device helpers for libnawsod (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdarg>
#include <cstdio>
#include "nawsod.h"

namespace nawsod {

void set_error(const char* fmt, ...);
int64_t get_tuning(const char* key, int64_t dflt);
int sm_count();

#define NAWSOD_REQUIRE(cond, code, ...)            \
  do {                                             \
    if (!(cond)) {                                 \
      ::nawsod::set_error(__VA_ARGS__);            \
      return (code);                               \
    }                                              \
  } while (0)

#define NAWSOD_CUDA_OK(expr)                                                         \
  do {                                                                               \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      ::nawsod::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                          __FILE__, __LINE__);                                       \
      return NAWSOD_ERR_CUDA;                                                        \
    }                                                                                \
  } while (0)

#define NAWSOD_LAUNCH_OK()                                                           \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) {                                                         \
      ::nawsod::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),\
                          __FILE__, __LINE__);                                       \
      return NAWSOD_ERR_CUDA;                                                        \
    }                                                                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

}  // namespace nawsod
