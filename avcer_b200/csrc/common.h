// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avcer_b200.h"

namespace avcer {

int set_error(const char* fmt, ...);   // stores the message, returns 1

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

#define AVCER_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) return avcer::set_error("%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define AVCER_REQUIRE(cond, ...)                           \
  do {                                                     \
    if (!(cond)) return avcer::set_error(__VA_ARGS__);     \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Storage-type helpers for kernels templated on float / __nv_bfloat16.
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

}  // namespace avcer
