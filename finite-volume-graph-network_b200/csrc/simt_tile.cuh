// simt_tile.cuh — shared building blocks of the strict-fp32 (FFMA) tile kernels: 64-row CTA tile, 256 threads,
// thread tile 4 rows x 8 cols, weights streamed from L2 in 16-wide K chunks and transposed on the fly.
#pragma once
#include "common.cuh"

namespace fvgn {

constexpr int BM = 64;    // rows per CTA
constexpr int NT = 256;   // threads per CTA
constexpr int HID = 128;  // latent width
constexpr int KC = 16;    // K chunk of the streamed weights
constexpr int WLD = 132;  // leading dim of Wt[k][n]
constexpr int HLD = 132;  // leading dim of the hidden tiles

struct WRegs { float4 v[2]; };

__device__ __forceinline__ void load_w_chunk(const float* __restrict__ W, int ldw, int k_valid, int n_valid, int k0, int tid,
                                             bool vec, WRegs& r) {
  const int kq = tid & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = (tid >> 2) + 64 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < n_valid) {
      const float* p = W + (size_t)n * ldw + k0 + kq * 4;
      if (vec) {
        v = ldg4(p);
      } else {
        const int k = k0 + kq * 4;
        if (k + 0 < k_valid) v.x = __ldg(p + 0);
        if (k + 1 < k_valid) v.y = __ldg(p + 1);
        if (k + 2 < k_valid) v.z = __ldg(p + 2);
        if (k + 3 < k_valid) v.w = __ldg(p + 3);
      }
    }
    r.v[i] = v;
  }
}

__device__ __forceinline__ void store_w_chunk(float* __restrict__ wt, int tid, const WRegs& r) {
  const int kq = tid & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = (tid >> 2) + 64 * i;
    float* p = wt + (kq * 4) * WLD + n;
    p[0 * WLD] = r.v[i].x;
    p[1 * WLD] = r.v[i].y;
    p[2 * WLD] = r.v[i].z;
    p[3 * WLD] = r.v[i].w;
  }
}

// acc[4][8] = As[64][KPAD] * W[n_valid][k_valid]^T   (zero padded to KPAD x 128)
template <int KPAD, int LDA>
__device__ __forceinline__ void gemm_layer(const float* __restrict__ As, const float* __restrict__ W, int ldw, int k_valid,
                                           int n_valid, float* __restrict__ Wt, float (&acc)[4][8], int tid) {
  const int tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  constexpr int NCH = KPAD / KC;
  const bool vec = ((ldw & 3) == 0) && (k_valid % KC == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  WRegs wr;
  load_w_chunk(W, ldw, k_valid, n_valid, 0, tid, vec, wr);
  store_w_chunk(Wt, tid, wr);
  __syncthreads();
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch + 1 < NCH) load_w_chunk(W, ldw, k_valid, n_valid, (ch + 1) * KC, tid, vec, wr);
    const float* wt = Wt + (ch & 1) * KC * WLD;
    const float* a0 = As + (tr * 4) * LDA + ch * KC;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * LDA + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + (kk + q) * WLD + tc * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (kk + q) * WLD + 64 + tc * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = (q == 0) ? a[i].x : (q == 1) ? a[i].y : (q == 2) ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, w0.x, acc[i][0]);
          acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]);
          acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]);
          acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]);
          acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
    if (ch + 1 < NCH) store_w_chunk(Wt + ((ch + 1) & 1) * KC * WLD, tid, wr);
    __syncthreads();
  }
}

// hidden = silu(acc + bias) -> Hs[64][HLD]
__device__ __forceinline__ void store_hidden(const float (&acc)[4][8], const float* __restrict__ bias, float* __restrict__ Hs, int tid) {
  const int tr = tid >> 4, tc = tid & 15;
  const float4 b0 = ldg4(bias + tc * 4);
  const float4 b1 = ldg4(bias + 64 + tc * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 o0, o1;
    o0.x = silu_f(acc[i][0] + b0.x); o0.y = silu_f(acc[i][1] + b0.y); o0.z = silu_f(acc[i][2] + b0.z); o0.w = silu_f(acc[i][3] + b0.w);
    o1.x = silu_f(acc[i][4] + b1.x); o1.y = silu_f(acc[i][5] + b1.y); o1.z = silu_f(acc[i][6] + b1.z); o1.w = silu_f(acc[i][7] + b1.w);
    float* h = Hs + (tr * 4 + i) * HLD;
    *reinterpret_cast<float4*>(h + tc * 4) = o0;
    *reinterpret_cast<float4*>(h + 64 + tc * 4) = o1;
  }
}


}  // namespace fvgn
