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
// cudaFuncSetAttribute (the dynamic shared-memory opt-in) is per DEVICE, not per process: first() is true once per device
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// CellBlock second aggregation (GN_blocks.py:110-138), 4 columns of one cell's 64-wide aggregate.  idx != NULL (mp_times >= 2): mean of
// the cell's three node aggregates, ((a + b) + c) / 3.0 in the reference's order (add, add, true division: no FMA contraction possible).
// idx == NULL (mp_times = 1): `na` already is the per-cell scatter-add [C,64] (GN_blocks.py:130-138) and is read as it is.
__device__ __forceinline__ float4 cell_agg4(const float* na, const int32_t* idx, size_t C, int c, int col) {
  if (idx == nullptr) return ldg4(na + (size_t)c * 64 + col);
  const int n0 = __ldg(idx + c), n1 = __ldg(idx + C + c), n2 = __ldg(idx + 2 * C + c);
  const float4 a = ldg4(na + (size_t)n0 * 64 + col), b = ldg4(na + (size_t)n1 * 64 + col), d = ldg4(na + (size_t)n2 * 64 + col);
  return make_float4(__fdiv_rn(__fadd_rn(__fadd_rn(a.x, b.x), d.x), 3.0f), __fdiv_rn(__fadd_rn(__fadd_rn(a.y, b.y), d.y), 3.0f),
                     __fdiv_rn(__fadd_rn(__fadd_rn(a.z, b.z), d.z), 3.0f), __fdiv_rn(__fadd_rn(__fadd_rn(a.w, b.w), d.w), 3.0f));
}
// tri / quad hybrid meshes ([4,C] table, 4th row -1 on triangles; oracle/fvgn_oracle.py::cell_block): ((a + b) + c) [+ d], then a true
// division by k = 3 or 4.  Bit-identical to cell_agg4 on triangles.
__device__ __forceinline__ float4 cell_agg4_poly(const float* na, const int32_t* idx4, size_t C, int c, int col) {
  const int n0 = __ldg(idx4 + c), n1 = __ldg(idx4 + C + c), n2 = __ldg(idx4 + 2 * C + c), n3 = __ldg(idx4 + 3 * C + c);
  const float4 a = ldg4(na + (size_t)n0 * 64 + col), b = ldg4(na + (size_t)n1 * 64 + col), d = ldg4(na + (size_t)n2 * 64 + col);
  float4 s = make_float4(__fadd_rn(__fadd_rn(a.x, b.x), d.x), __fadd_rn(__fadd_rn(a.y, b.y), d.y), __fadd_rn(__fadd_rn(a.z, b.z), d.z),
                         __fadd_rn(__fadd_rn(a.w, b.w), d.w));
  float k = 3.0f;
  if (n3 >= 0) {
    const float4 e = ldg4(na + (size_t)n3 * 64 + col);
    s = make_float4(__fadd_rn(s.x, e.x), __fadd_rn(s.y, e.y), __fadd_rn(s.z, e.z), __fadd_rn(s.w, e.w));
    k = 4.0f;
  }
  return make_float4(__fdiv_rn(s.x, k), __fdiv_rn(s.y, k), __fdiv_rn(s.z, k), __fdiv_rn(s.w, k));
}
__device__ __forceinline__ float silu_f(float v) {
  // x * sigmoid(x) with IEEE division; expf (not __expf) keeps the strict-fp32 path within ~1 ulp of torch's SiLU
  return v / (1.0f + expf(-v));
}

}  // namespace fvgn
