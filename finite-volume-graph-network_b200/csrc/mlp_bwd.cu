// mlp_bwd.cu — strict-fp32 backward of the fused gather -> MLP -> LayerNorm blocks (training path).
//
// One persistent CTA per SM walks 64-row tiles.  Per tile it RE-GATHERS the operand rows (nothing but the layer inputs was
// saved by the forward), recomputes the three layers in shared memory, then runs the chain
//   dOut -> LayerNorm' -> (dW3, db3, dh2) -> SiLU' -> (dW2, db2, dh1) -> SiLU' -> (dW1, db1, dA)
// with three tile GEMM flavours on FFMA: NT (x . W^T, forward), NN (dy . W, activation gradients) and TN (dy^T . x, weight
// gradients).  Weight/bias/LayerNorm gradients are accumulated into PER-CTA partial buffers by plain read-modify-write of
// the owning CTA (tiles in a fixed order) and summed over CTAs in a fixed order by reduce_partials_kernel, so the
// backward is run-to-run deterministic: no floating-point atomics anywhere.
// dA (gradient of the virtual concatenated input [x | agg] / [x'_s | x'_r | e]) is written to HBM; the transposes of the
// gathers are deterministic CSR segment sums in segment_sum_rows_kernel.
#include "simt_tile.cuh"

namespace fvgn {

enum { BW_ROWS = 0, BW_CELL = 1, BW_EDGE = 2 };

struct BwdParams {
  fvgn_mlp_t mlp;
  int mode, rows, ntiles, nseg;
  const float* in; int ld_in;                               // ROWS
  const float* src0; const float* src1; const int32_t* idx;  // CELL: x, node_agg, cells_node_T ; EDGE: x', e, neighbour_cell
  const float* d_out; int ld_dout;                           // [rows, n_out]
  float* d_in; int ld_din;                                   // [rows, k_in] or NULL
  const float* d_in_base; int ld_base;                       // optional: added onto columns 0..127 of d_in (residual path)
  float* part;                                               // [grid][part_stride]
  int part_stride;
};

__host__ __device__ inline int part_stride_of(int k_in) { return 128 * k_in + 2 * 128 * 128 + 5 * 128; }

// ---------------------------------------------------------------------------------------------- tile GEMMs
// acc (+)= As[64][0..127] . W[n][0..k_valid)^T   (W points at the first column of the 128-wide K segment)
template <int LDA>
__device__ __forceinline__ void gemm_nt(const float* __restrict__ As, const float* __restrict__ W, int ldw, int k_valid, int n_valid,
                                        float* __restrict__ Wt, float (&acc)[4][8], int tid) {
  const int tr = tid >> 4, tc = tid & 15;
  constexpr int NCH = 128 / KC;
  const bool vec = ((ldw & 3) == 0) && (k_valid == 128) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
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
          acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
    if (ch + 1 < NCH) store_w_chunk(Wt + ((ch + 1) & 1) * KC * WLD, tid, wr);
    __syncthreads();
  }
}

// W chunk for the NN product: Wt[k = o][n = i] = W[o0 + o][col0 + i]   (no transpose; 16 rows x 128 cols)
__device__ __forceinline__ void load_w_rows(const float* __restrict__ W, int ldw, int o_valid, int col0, int ncols_valid, int o0, int tid,
                                            bool vec, WRegs& r) {
  const int o = o0 + (tid >> 4), c4 = tid & 15;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = c4 * 4 + 64 * h;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o < o_valid) {
      const float* p = W + (size_t)o * ldw + col0 + c;
      if (vec) {
        v = ldg4(p);
      } else {
        if (c + 0 < ncols_valid) v.x = __ldg(p + 0);
        if (c + 1 < ncols_valid) v.y = __ldg(p + 1);
        if (c + 2 < ncols_valid) v.z = __ldg(p + 2);
        if (c + 3 < ncols_valid) v.w = __ldg(p + 3);
      }
    }
    r.v[h] = v;
  }
}
__device__ __forceinline__ void store_w_rows(float* __restrict__ wt, int tid, const WRegs& r) {
  float* p = wt + (tid >> 4) * WLD + (tid & 15) * 4;
  *reinterpret_cast<float4*>(p) = r.v[0];
  *reinterpret_cast<float4*>(p + 64) = r.v[1];
}

// acc = As[64][0..127] . W[0..o_valid)[col0 .. col0+128)   (activation gradient: dh = dy . W)
template <int LDA>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ As, const float* __restrict__ W, int ldw, int o_valid, int col0,
                                        int ncols_valid, float* __restrict__ Wt, float (&acc)[4][8], int tid) {
  const int tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  constexpr int NCH = 128 / KC;
  const bool vec = ((ldw & 3) == 0) && ((col0 & 3) == 0) && (ncols_valid == 128) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  WRegs wr;
  load_w_rows(W, ldw, o_valid, col0, ncols_valid, 0, tid, vec, wr);
  store_w_rows(Wt, tid, wr);
  __syncthreads();
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch + 1 < NCH) load_w_rows(W, ldw, o_valid, col0, ncols_valid, (ch + 1) * KC, tid, vec, wr);
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
          acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
    if (ch + 1 < NCH) store_w_rows(Wt + ((ch + 1) & 1) * KC * WLD, tid, wr);
    __syncthreads();
  }
}

// dst[o][i] (+)= sum_r P[r][o] * Q[r][i]  over the tile's 64 rows; thread tile 8 (o) x 8 (i).  dst is this CTA's partial.
template <int LD>
__device__ __forceinline__ void gemm_tn_rmw(const float* __restrict__ P, const float* __restrict__ Q, float* __restrict__ dst, int ld_dst,
                                            int o_valid, int i_valid, bool first, int tid) {
  const int tr = tid >> 4, tc = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int r = 0; r < BM; ++r) {
    const float4 p0 = *reinterpret_cast<const float4*>(P + r * LD + tr * 8);
    const float4 p1 = *reinterpret_cast<const float4*>(P + r * LD + tr * 8 + 4);
    const float4 q0 = *reinterpret_cast<const float4*>(Q + r * LD + tc * 4);
    const float4 q1 = *reinterpret_cast<const float4*>(Q + r * LD + 64 + tc * 4);
    const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(pv[i], qv[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = tr * 8 + i;
    if (o >= o_valid) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = tc * 4 + 64 * h;
      float* d = dst + (size_t)o * ld_dst + c;
      if (c + 3 < i_valid && ((ld_dst & 3) == 0)) {
        float4 old = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(d);
        *reinterpret_cast<float4*>(d) = make_float4(old.x + acc[i][4 * h], old.y + acc[i][4 * h + 1], old.z + acc[i][4 * h + 2], old.w + acc[i][4 * h + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < i_valid) d[j] = (first ? 0.f : d[j]) + acc[i][4 * h + j];
      }
    }
  }
}

__device__ __forceinline__ float colsum64(const float* __restrict__ buf, int c) {
  float s = 0.f;
#pragma unroll 8
  for (int r = 0; r < BM; ++r) s += buf[r * HLD + c];
  return s;
}

__device__ __forceinline__ float dsilu(float a) {
  const float s = 1.0f / (1.0f + expf(-a));
  return s * (1.0f + a * (1.0f - s));
}

// write the thread tile (4 rows x 8 cols) into a [64][HLD] shared buffer
__device__ __forceinline__ void st_tile(float* __restrict__ buf, const float (&v)[4][8], int tid) {
  const int tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* h = buf + (tr * 4 + i) * HLD;
    *reinterpret_cast<float4*>(h + tc * 4) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
    *reinterpret_cast<float4*>(h + 64 + tc * 4) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
  }
}

// gather one 128-wide K segment of the tile's operand rows into S[64][HLD] (zero padded)
__device__ __forceinline__ void gather_seg(const BwdParams& p, int seg, int row0, float* __restrict__ S, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  if (p.mode == BW_EDGE) {
    const int F = p.rows;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i, f = row0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < F) {
        if (seg < 2) {
          const int c = __ldg(p.idx + (size_t)seg * F + f);
          v = ldg4(p.src0 + (size_t)c * HID + lane * 4);
        } else {
          v = ldg4(p.src1 + (size_t)f * HID + lane * 4);
        }
      }
      *reinterpret_cast<float4*>(S + r * HLD + lane * 4) = v;
    }
  } else if (p.mode == BW_CELL) {
    const int C = p.rows;
    if (seg == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 8 + i, c = row0 + r;
        const float4 v = (c < C) ? ldg4(p.src0 + (size_t)c * HID + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(S + r * HLD + lane * 4) = v;
      }
    } else {
      const int hl = lane & 15, hw = lane >> 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = warp * 8 + i * 2 + hw, c = row0 + r;
        const float4 o = (c < C) ? cell_agg4(p.src1, p.idx, (size_t)C, c, hl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(S + r * HLD + hl * 4) = o;
        *reinterpret_cast<float4*>(S + r * HLD + 64 + hl * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else {
    const int k_in = p.mlp.k_in;
    for (int idx = tid; idx < BM * 128; idx += NT) {
      const int r = idx >> 7, k = idx & 127;
      const int g = row0 + r, kk = seg * 128 + k;
      S[r * HLD + k] = (g < p.rows && kk < k_in) ? p.in[(size_t)g * p.ld_in + kk] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(NT, 1) mlp_bwd_simt_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float smem[];
  float* S = smem;
  float* A1 = S + BM * HLD;
  float* A2 = A1 + BM * HLD;
  float* T1 = A2 + BM * HLD;
  float* T2 = T1 + BM * HLD;
  float* Wt = T2 + BM * HLD;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const int k_in = p.mlp.k_in, n_out = p.mlp.n_out;
  const bool has_ln = p.mlp.ln_w != nullptr;
  float* part = p.part + (size_t)blockIdx.x * p.part_stride;
  float* pW1 = part;
  float* pW2 = pW1 + 128 * k_in;
  float* pW3 = pW2 + 128 * 128;
  float* pvec = pW3 + 128 * 128;  // db1 | db2 | db3 | dgamma | dbeta
  int col[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { col[j] = tc * 4 + j; col[4 + j] = 64 + tc * 4 + j; }
  float s_db1 = 0.f, s_db2 = 0.f, s_db3 = 0.f, s_dg = 0.f, s_dbeta = 0.f;  // column tid (threads < 128)
  bool first = true;

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int row0 = tile * BM;
    float acc[4][8];
    const float* zs = p.mlp.z_save;
    if (zs != nullptr) {
      // ---------------------------------------------------------------- saved pre-activations (z1 | z2 | z3) instead of the recompute
      __syncthreads();  // the previous tile's last products may still be reading A1 / T1
      float v[4][8], hh[4][8];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int g = min(row0 + tr * 4 + i, p.rows - 1);
          const float4 a = ldg4(zs + (size_t)g * 384 + 128 * l + tc * 4), b = ldg4(zs + (size_t)g * 384 + 128 * l + 64 + tc * 4);
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
        }
        if (l < 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) hh[i][j] = silu_f(v[i][j]);
          st_tile(l == 0 ? A1 : A2, v, tid);
          st_tile(l == 0 ? T1 : T2, hh, tid);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = v[i][j];  // z3 already holds the bias
        }
      }
    } else {
    // ------------------------------------------------------------------ forward recompute
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int seg = 0; seg < p.nseg; ++seg) {
      gather_seg(p, seg, row0, S, tid);
      const int kv = min(128, k_in - seg * 128);
      gemm_nt<HLD>(S, p.mlp.w1 + seg * 128, k_in, kv, 128, Wt, acc, tid);
    }
    {
      float v[4][8], hh[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = acc[i][j] + __ldg(p.mlp.b1 + col[j]); hh[i][j] = silu_f(v[i][j]); acc[i][j] = 0.f; }
      st_tile(A1, v, tid);
      st_tile(T1, hh, tid);
    }
    gemm_nt<HLD>(T1, p.mlp.w2, 128, 128, 128, Wt, acc, tid);
    {
      float v[4][8], hh[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = acc[i][j] + __ldg(p.mlp.b2 + col[j]); hh[i][j] = silu_f(v[i][j]); acc[i][j] = 0.f; }
      st_tile(A2, v, tid);
      st_tile(T2, hh, tid);
    }
    gemm_nt<HLD>(T2, p.mlp.w3, 128, 128, n_out, Wt, acc, tid);
    }
    // ------------------------------------------------------------------ dOut, LayerNorm backward -> dy (registers)
    float dy[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = row0 + tr * 4 + i;
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[i][j] = (g < p.rows && col[j] < n_out) ? p.d_out[(size_t)g * p.ld_dout + col[j]] : 0.f;
    }
    if (has_ln) {
      float xh[4][8], rstd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8], s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = (zs != nullptr) ? acc[i][j] : acc[i][j] + __ldg(p.mlp.b3 + col[j]); }
        s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 128.0f);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        rstd[i] = 1.0f / sqrtf(q * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 8; ++j) xh[i][j] = (v[j] - mean) * rstd[i];
      }
      // dbeta = colsum(dOut), dgamma = colsum(dOut * xhat): staged through S, one fixed-order column sum per thread
      st_tile(S, dy, tid);
      __syncthreads();
      if (tid < 128) s_dbeta += colsum64(S, tid);
      __syncthreads();
      {
        float t[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) t[i][j] = dy[i][j] * xh[i][j];
        st_tile(S, t, tid);
      }
      __syncthreads();
      if (tid < 128) s_dg += colsum64(S, tid);
      __syncthreads();
      float gam[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) gam[j] = __ldg(p.mlp.ln_w + col[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float dxh[8], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { dxh[j] = dy[i][j] * gam[j]; m1 += dxh[j]; m2 = fmaf(dxh[j], xh[i][j], m2); }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
        m1 *= (1.0f / 128.0f); m2 *= (1.0f / 128.0f);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[i][j] = rstd[i] * (dxh[j] - m1 - xh[i][j] * m2);
      }
    }
    st_tile(S, dy, tid);
    __syncthreads();
    if (tid < 128) s_db3 += colsum64(S, tid);
    // ------------------------------------------------------------------ layer 3
    gemm_tn_rmw<HLD>(S, T2, pW3, 128, n_out, 128, first, tid);                 // dW3[o][i] += dy[.,o] h2[.,i]
    const float* da_in = p.mlp.da_in;  // (da2 | da1) precomputed on the tensor cores: only the TN products remain here
    if (da_in != nullptr) {
      float t[4][8];
      __syncthreads();  // the TN product above still reads T2 (h2)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int g = min(row0 + tr * 4 + i, p.rows - 1);
        const bool ok = row0 + tr * 4 + i < p.rows;
        const float4 a = ldg4(da_in + (size_t)g * 256 + tc * 4), b = ldg4(da_in + (size_t)g * 256 + 64 + tc * 4);
        t[i][0] = ok ? a.x : 0.f; t[i][1] = ok ? a.y : 0.f; t[i][2] = ok ? a.z : 0.f; t[i][3] = ok ? a.w : 0.f;
        t[i][4] = ok ? b.x : 0.f; t[i][5] = ok ? b.y : 0.f; t[i][6] = ok ? b.z : 0.f; t[i][7] = ok ? b.w : 0.f;
      }
      st_tile(T2, t, tid);
    } else {
    gemm_nn<HLD>(S, p.mlp.w3, 128, n_out, 0, 128, Wt, acc, tid);               // dh2 = dy . W3
    {
      float t[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t[i][j] = acc[i][j] * dsilu(A2[(tr * 4 + i) * HLD + tc * 4 + j]);
          t[i][4 + j] = acc[i][4 + j] * dsilu(A2[(tr * 4 + i) * HLD + 64 + tc * 4 + j]);
        }
      st_tile(T2, t, tid);  // da2 (h2 is dead: the TN and NN products that read it are behind a barrier)
    }
    }
    __syncthreads();
    if (tid < 128) s_db2 += colsum64(T2, tid);
    // ------------------------------------------------------------------ layer 2
    gemm_tn_rmw<HLD>(T2, T1, pW2, 128, 128, 128, first, tid);                  // dW2 += da2^T h1
    if (da_in != nullptr) {
      float t[4][8];
      __syncthreads();  // the TN product above still reads T1 (h1)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int g = min(row0 + tr * 4 + i, p.rows - 1);
        const bool ok = row0 + tr * 4 + i < p.rows;
        const float4 a = ldg4(da_in + (size_t)g * 256 + 128 + tc * 4), b = ldg4(da_in + (size_t)g * 256 + 128 + 64 + tc * 4);
        t[i][0] = ok ? a.x : 0.f; t[i][1] = ok ? a.y : 0.f; t[i][2] = ok ? a.z : 0.f; t[i][3] = ok ? a.w : 0.f;
        t[i][4] = ok ? b.x : 0.f; t[i][5] = ok ? b.y : 0.f; t[i][6] = ok ? b.z : 0.f; t[i][7] = ok ? b.w : 0.f;
      }
      st_tile(T1, t, tid);
    } else {
    gemm_nn<HLD>(T2, p.mlp.w2, 128, 128, 0, 128, Wt, acc, tid);                // dh1 = da2 . W2
    {
      float t[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t[i][j] = acc[i][j] * dsilu(A1[(tr * 4 + i) * HLD + tc * 4 + j]);
          t[i][4 + j] = acc[i][4 + j] * dsilu(A1[(tr * 4 + i) * HLD + 64 + tc * 4 + j]);
        }
      st_tile(T1, t, tid);  // da1
    }
    }
    __syncthreads();
    if (tid < 128) s_db1 += colsum64(T1, tid);
    // ------------------------------------------------------------------ layer 1: per 128-wide K segment
    for (int seg = 0; seg < p.nseg; ++seg) {
      gather_seg(p, seg, row0, S, tid);
      __syncthreads();
      const int kv = min(128, k_in - seg * 128);
      gemm_tn_rmw<HLD>(T1, S, pW1 + seg * 128, k_in, 128, kv, first, tid);     // dW1[:, seg] += da1^T A_seg
      if (p.d_in != nullptr && da_in == nullptr) {
        gemm_nn<HLD>(T1, p.mlp.w1, k_in, 128, seg * 128, kv, Wt, acc, tid);   // dA_seg = da1 . W1[:, seg]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int g = row0 + tr * 4 + i;
          if (g >= p.rows) continue;
          float* d = p.d_in + (size_t)g * p.ld_din + seg * 128;
          if (seg == 0 && p.d_in_base != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] += p.d_in_base[(size_t)g * p.ld_base + col[j]];
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = tc * 4 + 64 * h;
            if (c + 3 < kv && ((p.ld_din & 3) == 0)) {
              *reinterpret_cast<float4*>(d + c) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c + j < kv) d[c + j] = acc[i][4 * h + j];
            }
          }
        }
      } else {
        __syncthreads();
      }
    }
    first = false;
  }
  if (tid < 128) {
    pvec[tid] = s_db1; pvec[128 + tid] = s_db2; pvec[256 + tid] = s_db3; pvec[384 + tid] = s_dg; pvec[512 + tid] = s_dbeta;
  }
}

// out[j] = sum over CTAs (fixed order) of part[cta][j]
__global__ void reduce_partials_kernel(const float* __restrict__ part, int n_cta, int stride, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= stride) return;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += part[(size_t)c * stride + j];
  out[j] = s;
}

// Deterministic transposes of the gathers: out[d, 0:W] = base[d] + scale * sum_{pos in CSR(d)} src[pos % n_src, col0 + (pos / n_src) * part_stride + 0:W]
__global__ void __launch_bounds__(256) segment_sum_rows_kernel(const float* __restrict__ src, int ld_src, int n_src, int col0, int part_stride,
                                                               int W4 /* width / 4 */, const int32_t* __restrict__ ptr,
                                                               const int32_t* __restrict__ items, int n_dst, const float* __restrict__ base,
                                                               int ld_base, float scale, float* __restrict__ out, int ld_out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = (int)(gid / W4), q = (int)(gid % W4);
  if (d >= n_dst) return;
  const int beg = __ldg(ptr + d), end = __ldg(ptr + d + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = beg; i < end; ++i) {
    const int pos = __ldg(items + i);
    const int part = pos / n_src, row = pos - part * n_src;
    const float4 v = ldg4(src + (size_t)row * ld_src + col0 + part * part_stride + q * 4);
    acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
  }
  float4 o = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  if (base) {
    const float4 b = ldg4(base + (size_t)d * ld_base + q * 4);
    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
  }
  *reinterpret_cast<float4*>(out + (size_t)d * ld_out + q * 4) = o;
}

// g_e_in[f, j] = g_e_out[f, j] + dA_edge[f, 256 + j] + g_node[face_node[j < 64 ? 0 : 1][f], j % 64]
__global__ void edge_grad_combine_kernel(const float* __restrict__ g_e_out, const float* __restrict__ dA_edge, const float* __restrict__ g_node,
                                         const int32_t* __restrict__ fn, int F, float* __restrict__ g_e_in) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = (int)(gid >> 5), q = (int)(gid & 31);
  if (f >= F) return;
  const int n = __ldg(fn + (size_t)(q >> 4) * F + f);
  const float4 a = ldg4(g_e_out + (size_t)f * 128 + q * 4);
  const float4 b = ldg4(dA_edge + (size_t)f * 384 + 256 + q * 4);
  const float4 c = ldg4(g_node + (size_t)n * 64 + (q & 15) * 4);
  *reinterpret_cast<float4*>(g_e_in + (size_t)f * 128 + q * 4) = make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
}

static size_t bwd_smem_bytes() { return sizeof(float) * (size_t)(5 * BM * HLD + 2 * KC * WLD); }

}  // namespace fvgn

using namespace fvgn;

extern "C" size_t fvgn_mlp_bwd_workspace_floats(int k_in) { return (size_t)(sm_count() + 1) * part_stride_of(k_in); }
extern "C" size_t fvgn_mlp_grad_floats(int k_in) { return (size_t)part_stride_of(k_in); }

static int launch_bwd(BwdParams& p, float* grads, float* workspace, cudaStream_t st) {
  FVGN_CHECK_ARG(p.mlp.w1 && p.mlp.b1 && p.mlp.w2 && p.mlp.b2 && p.mlp.w3 && p.mlp.b3, "mlp weight pointer is NULL");
  FVGN_CHECK_ARG(p.mlp.k_in >= 1 && p.mlp.k_in <= 384 && p.mlp.n_out >= 1 && p.mlp.n_out <= 128, "bad mlp dims");
  FVGN_CHECK_ARG(!p.mlp.ln_w || (p.mlp.ln_b && p.mlp.n_out == 128), "LayerNorm needs n_out == 128");
  FVGN_CHECK_ARG(grads && workspace && p.d_out, "NULL pointer");
  FVGN_CHECK_ARG(p.rows > 0, "rows must be > 0");
  p.ntiles = cdiv(p.rows, BM);
  p.nseg = (p.mlp.k_in + 127) / 128;
  p.part_stride = part_stride_of(p.mlp.k_in);
  p.part = workspace;
  const int grid = p.ntiles < sm_count() ? p.ntiles : sm_count();
  static ::fvgn::PerDeviceOnce configured;
  if (configured.first()) {
    FVGN_CUDA(cudaFuncSetAttribute(mlp_bwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes()));
  }
  mlp_bwd_simt_kernel<<<grid, NT, bwd_smem_bytes(), st>>>(p);
  FVGN_LAUNCH_CHECK();
  reduce_partials_kernel<<<cdiv(p.part_stride, 256), 256, 0, st>>>(workspace, grid, p.part_stride, grads);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_mlp_rows_bwd(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, const float* d_out, int ld_dout, float* d_in,
                                 int ld_din, float* grads, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(mlp && in, "NULL pointer");
  BwdParams p{};
  p.mlp = *mlp; p.mode = BW_ROWS; p.rows = rows; p.in = in; p.ld_in = ld_in; p.d_out = d_out; p.ld_dout = ld_dout; p.d_in = d_in; p.ld_din = ld_din;
  return launch_bwd(p, grads, workspace, as_stream(stream));
}

extern "C" int fvgn_cell_block_bwd(const fvgn_mlp_t* mlp, const float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells,
                                   const float* d_x_prime, const float* d_x_base, int ld_base, float* d_in /*[C,192]*/, float* grads,
                                   float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(mlp && x && node_agg && (d_in || mlp->da_in), "NULL pointer");  // cells_node_T == NULL: mp_times = 1
  FVGN_CHECK_ARG(mlp->k_in == 192 && mlp->ln_w, "CellBlock MLP must be 192 -> 128 with LayerNorm");
  BwdParams p{};
  p.mlp = *mlp; p.mode = BW_CELL; p.rows = n_cells; p.src0 = x; p.src1 = node_agg; p.idx = cells_node_T; p.d_out = d_x_prime; p.ld_dout = 128;
  p.d_in = d_in; p.ld_din = 192; p.d_in_base = d_x_base; p.ld_base = ld_base;
  return launch_bwd(p, grads, workspace, as_stream(stream));
}

extern "C" int fvgn_edge_block_bwd(const fvgn_mlp_t* mlp, const float* x_prime, const float* e, const int32_t* neighbour_cell, int n_faces,
                                   const float* d_e_prime, float* d_in /*[F,384]*/, float* grads, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(mlp && x_prime && e && neighbour_cell && (d_in || mlp->da_in), "NULL pointer");
  FVGN_CHECK_ARG(mlp->k_in == 384 && mlp->ln_w, "EdgeBlock MLP must be 384 -> 128 with LayerNorm");
  BwdParams p{};
  p.mlp = *mlp; p.mode = BW_EDGE; p.rows = n_faces; p.src0 = x_prime; p.src1 = e; p.idx = neighbour_cell; p.d_out = d_e_prime; p.ld_dout = 128;
  p.d_in = d_in; p.ld_din = 384;
  return launch_bwd(p, grads, workspace, as_stream(stream));
}

extern "C" int fvgn_segment_sum_rows(const float* src, int ld_src, int n_src, int col0, int part_stride, int width, const int32_t* ptr,
                                     const int32_t* items, int n_dst, const float* base, int ld_base, float scale, float* out, int ld_out,
                                     fvgn_stream_t stream) {
  FVGN_CHECK_ARG(src && ptr && items && out, "NULL pointer");
  FVGN_CHECK_ARG(width > 0 && (width & 3) == 0 && (ld_src & 3) == 0 && (ld_out & 3) == 0 && (col0 & 3) == 0 && (part_stride & 3) == 0,
                 "width / strides must be multiples of 4 floats");
  FVGN_CHECK_ARG(!base || (ld_base & 3) == 0, "base stride must be a multiple of 4 floats");
  if (n_dst <= 0) return 0;
  const long long threads = (long long)n_dst * (width / 4);
  segment_sum_rows_kernel<<<cdiv(threads, 256), 256, 0, as_stream(stream)>>>(src, ld_src, n_src, col0, part_stride, width / 4, ptr, items, n_dst,
                                                                            base, ld_base, scale, out, ld_out);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_edge_grad_combine(const float* g_e_out, const float* dA_edge, const float* g_node, const int32_t* face_node, int n_faces,
                                      float* g_e_in, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(g_e_out && dA_edge && g_node && face_node && g_e_in, "NULL pointer");
  if (n_faces <= 0) return 0;
  edge_grad_combine_kernel<<<cdiv((long long)n_faces * 32, 256), 256, 0, as_stream(stream)>>>(g_e_out, dA_edge, g_node, face_node, n_faces, g_e_in);
  FVGN_LAUNCH_CHECK();
  return 0;
}
