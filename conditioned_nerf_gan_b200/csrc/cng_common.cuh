// Shared helpers for the cng_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "cng_b200.h"

namespace cng {

void set_error(const char* fmt, ...);

// Records the message and returns `code` (used as `return fail(CNG_ERR_..., "...")`).
int fail(int code, const char* fmt, ...);

// Launch check: cudaGetLastError() -> message + code.
int check_launch(const char* what);

inline cudaStream_t as_stream(cng_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached).
int sm_count();

#define CNG_REQUIRE(cond, code, ...)          \
  do {                                        \
    if (!(cond)) return ::cng::fail(code, __VA_ARGS__); \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cng
