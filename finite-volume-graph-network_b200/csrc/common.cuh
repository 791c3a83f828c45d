// common.cuh — shared helpers for libfvgn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fvgn_b200.h"

namespace fvgn {

void set_error(const char* fmt, ...);

#define FVGN_CHECK_ARG(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::fvgn::set_error("%s: bad argument: %s", __func__, msg);     \
      return FVGN_E_BADARG;                                         \
    }                                                               \
  } while (0)

#define FVGN_CUDA(call)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::fvgn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));  \
      return (int)e__;                                                                          \
    }                                                                                           \
  } while (0)

#define FVGN_LAUNCH_CHECK()                                                                     \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) {                                                                   \
      ::fvgn::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));     \
      return (int)e__;                                                                          \
    }                                                                                           \
  } while (0)

static inline cudaStream_t as_stream(fvgn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int sm_count();

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float silu_f(float v) {
  // x * sigmoid(x) with IEEE division; expf (not __expf) keeps the strict-fp32 path within ~1 ulp of torch's SiLU
  return v / (1.0f + expf(-v));
}

}  // namespace fvgn
