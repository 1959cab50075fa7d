// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/avcer_b200.h"

// ---- 16-bit storage type of this build.  The library is compiled twice from the same sources: libavcer_b200.so stores
// activations / weights as bfloat16 (the default "bf16" precision), libavcer_b200_fp16.so (-DAVCER_HALF) as IEEE half
// ("fp16" precision: same kernels, same throughput, 11 instead of 8 mantissa bits -- the wide-init probability error drops
// from 6-8e-3 to ~1e-3; the price is the 65504 range, which the path's activations, <= ~200, do not come near).
// In the half build the bf16 spellings below are redirected to their half twins, so the kernel sources are shared verbatim;
// the few places that depend on the bit layout (UMMA instruction descriptor, mma.sync type, a shift-based unpack) test
// AVCER_HALF explicitly.  The C ABI is unchanged: dtype code AVCER_BF16 means "the 16-bit storage type of this library".
#ifdef AVCER_HALF
#define __nv_bfloat16 __half
#define __nv_bfloat162 __half2
#define __floats2bfloat162_rn __floats2half2_rn
#define __bfloat1622float2 __half22float2
#define __float2bfloat16_rn __float2half_rn
#define __bfloat162float __half2float
#define AVCER_STORAGE_NAME "fp16"
#else
#define AVCER_STORAGE_NAME "bf16"
#endif

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

// SM count persistent kernels size their grids with: the device's (cached; 148 on B200), or the caller's smaller
// share of it (avcer_set_sm_limit) when two branches of the pipeline run side by side on different streams.
int& sm_limit_ref();
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  const int lim = sm_limit_ref();
  return (lim > 0 && lim < n) ? lim : n;
}

// Programmatic dependent launch (PDL).  Hot-path kernels are launched with the programmatic-stream-serialization
// attribute: the next kernel's CTAs may become resident and run their prologue (barrier init, TMEM allocation,
// descriptor prefetch, constant staging) while the previous kernel drains, and block in pdl_wait() until that kernel
// has completed and flushed its writes.  Every kernel launched through launch_pdl() must call pdl_wait() before its
// first access to memory another kernel may have written, and before its own first global write.  Works inside CUDA
// graph capture (programmatic edges).  AVCER_PDL=0 launches the same kernels fully serialised (3 % slower audio
// forward, 1.5 % slower VS forward).
// History: PDL's simultaneous release of all CTAs exposed a missing generic->async proxy fence in the residual ring of
// the barrier epilogue (tc_gemm.cuh / tc_gemm2.cuh): the TMA refill of a slot could overtake the last ld.shared of the
// previous occupant (about 2 % of the audio encoder's residual GEMM launches showed 16-byte-granular corruptions);
// tests/test_gpu_nets.py::test_forwards_are_bit_stable_run_to_run and scripts/probe_layer_race.py guard it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Off while the caller runs with an SM share (avcer_set_sm_limit): an early-resident dependent kernel would sit on the
// SMs the share leaves free for the other branch's stream.
inline bool pdl_enabled() {
  static const int on = getenv("AVCER_PDL") ? atoi(getenv("AVCER_PDL")) : 1;
  return on != 0 && sm_limit_ref() == 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// launch_pdl for persistent single-CTA kernels.  While the caller runs with an SM share (avcer_set_sm_limit) the CTAs are
// launched as clusters of two, which the hardware places on the two SMs of one TPC: a kernel of one branch then leaves
// whole TPCs free, so the two-SM (cta_group::2) kernels of the other branch can still be placed beside it.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_tpc(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (sm_limit_ref() > 0 && grid.x % 2 == 0 && grid.y == 1 && grid.z == 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Storage-type helpers for kernels templated on float / __nv_bfloat16.
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

}  // namespace avcer
