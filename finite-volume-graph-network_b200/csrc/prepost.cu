// prepost.cu — the small fp32 kernels either side of the network (compiled with --fmad=false so every product and
// sum is rounded separately, in the reference's evaluation order):
//   a2/a3  node->cell / node->face interpolation, edge_attr assembly, create_next_graph   (dataset/Load_mesh.py:904-1115)
//   a4/a13 online Normalizer accumulate / apply / inverse + one-hot feature assembly        (utils/normalization.py, FVGN.py:83-138)
//   a10    face-area feature + BatchNorm1d(1)                                                (FVGN.py:209-227)
//   a11    Intergrator: face flux -> cell divergence                                        (EncoderProcesserDecoder.py:251-331)
//   a12    compute_FVM_loss                                                                  (utils/loss_compute.py:24-115)
// All reductions are two-level with a fixed tree (no floating-point atomics) => run-to-run deterministic.
#include "common.cuh"

namespace fvgn {

constexpr int RED_BLOCKS = 128;   // first-level partials
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    for (int o = 16; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// ------------------------------------------------------------------------------------------------ a2/a3
__global__ void interp_nodes_kernel(const float* __restrict__ nf, int ld, int col0, int width, const int32_t* __restrict__ cnT,
                                    const float* __restrict__ cf, const int32_t* __restrict__ fn, int C, int F,
                                    float* __restrict__ on_cell, float* __restrict__ on_face) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (on_cell && i < C) {
    const int n0 = cnT[i], n1 = cnT[(size_t)C + i], n2 = cnT[2 * (size_t)C + i];
    const float f0 = cf[3 * i], f1 = cf[3 * i + 1], f2 = cf[3 * i + 2];
    for (int w = 0; w < width; ++w) {  // (f0*a + f1*b) + f2*c   (Load_mesh.py:908-915 / 989-996)
      const float a = nf[(size_t)n0 * ld + col0 + w], b = nf[(size_t)n1 * ld + col0 + w], c = nf[(size_t)n2 * ld + col0 + w];
      on_cell[(size_t)i * width + w] = (f0 * a + f1 * b) + f2 * c;
    }
  }
  if (on_face && i < F) {
    const int a = fn[i], b = fn[(size_t)F + i];
    for (int w = 0; w < width; ++w)  // (x[a] + x[b]) / 2.0   (:917-920 / 998-1001)
      on_face[(size_t)i * width + w] = (nf[(size_t)a * ld + col0 + w] + nf[(size_t)b * ld + col0 + w]) / 2.0f;
  }
}

// tri / quad hybrid meshes: [4,C] node table (4th row -1 on triangles) and four weights per cell (oracle/fvgn_oracle.py:
// ((f0 a + f1 b) + f2 c) + f3 d, the last term only where the cell has a 4th node)
__global__ void interp_cells_poly_kernel(const float* __restrict__ nf, int ld, int col0, int width, const int32_t* __restrict__ cn4T,
                                         const float* __restrict__ cf4, int C, float* __restrict__ on_cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const int n0 = cn4T[i], n1 = cn4T[(size_t)C + i], n2 = cn4T[2 * (size_t)C + i], n3 = cn4T[3 * (size_t)C + i];
  const float f0 = cf4[4 * i], f1 = cf4[4 * i + 1], f2 = cf4[4 * i + 2], f3 = cf4[4 * i + 3];
  for (int w = 0; w < width; ++w) {
    const float a = nf[(size_t)n0 * ld + col0 + w], b = nf[(size_t)n1 * ld + col0 + w], c = nf[(size_t)n2 * ld + col0 + w];
    float v = (f0 * a + f1 * b) + f2 * c;
    if (n3 >= 0) v = v + f3 * nf[(size_t)n3 * ld + col0 + w];
    on_cell[(size_t)i * width + w] = v;
  }
}

__device__ __forceinline__ bool is_interior(float t) { return t == 0.0f || t == 5.0f; }  // NORMAL or OUTFLOW (:931-934)

__global__ void assemble_edge_attr_kernel(const float* __restrict__ uv, int ld_uv, const float* __restrict__ cen, const int32_t* __restrict__ nc,
                                          const float* __restrict__ ftl, const float* __restrict__ bc, int ld_bc, const uint8_t* __restrict__ keep,
                                          int F, float* __restrict__ ea, uint8_t* __restrict__ mask) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  int s = nc[f], r = nc[(size_t)F + f];
  if (keep && !keep[f]) { const int t = s; s = r; r = t; }  // random direction: features only (App. D.3)
  const bool interior = is_interior(ftl[2 * f]);
  float du = uv[(size_t)s * ld_uv] - uv[(size_t)r * ld_uv];
  float dv = uv[(size_t)s * ld_uv + 1] - uv[(size_t)r * ld_uv + 1];
  if (!interior) { du = bc[(size_t)f * ld_bc]; dv = bc[(size_t)f * ld_bc + 1]; }  // enforce BC (:968-970 / 1059-1062)
  ea[5 * (size_t)f + 0] = du;
  ea[5 * (size_t)f + 1] = dv;
  ea[5 * (size_t)f + 2] = cen[2 * s] - cen[2 * r];
  ea[5 * (size_t)f + 3] = cen[2 * s + 1] - cen[2 * r + 1];
  ea[5 * (size_t)f + 4] = ftl[2 * f + 1];
  if (mask) mask[f] = interior ? 1 : 0;
}

__global__ void next_graph_kernel(const float* __restrict__ nuv, const int32_t* __restrict__ nc, const float* __restrict__ ftl,
                                  const float* __restrict__ bc, int ld_bc, int C, int F, float* __restrict__ x, float* __restrict__ ea) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { x[4 * (size_t)i + 2] = nuv[2 * i]; x[4 * (size_t)i + 3] = nuv[2 * i + 1]; }  // x = [type, Re, u, v] (:1088)
  if (i < F) {
    const int s = nc[i], r = nc[(size_t)F + i];
    float du = nuv[2 * s] - nuv[2 * r], dv = nuv[2 * s + 1] - nuv[2 * r + 1];
    if (!is_interior(ftl[2 * i])) { du = bc[(size_t)i * ld_bc]; dv = bc[(size_t)i * ld_bc + 1]; }  // (:1110-1112)
    ea[5 * (size_t)i] = du;
    ea[5 * (size_t)i + 1] = dv;
  }
}

// ------------------------------------------------------------------------------------------------ a4/a13
// feature j of row r: j < width_raw -> x[r, j] ; else one_hot(type[r])[j - width_raw]
__device__ __forceinline__ float feat(const float* __restrict__ x, int ld_x, int width_raw, const float* __restrict__ type_col, int ld_type,
                                      int r, int j) {
  if (j < width_raw) return x[(size_t)r * ld_x + j];
  return ((int)type_col[(size_t)r * ld_type] == j - width_raw) ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(RED_THREADS) norm_partial_kernel(const float* __restrict__ x, int ld_x, int width_raw,
                                                                  const float* __restrict__ type_col, int ld_type, int width, int rows,
                                                                  double* __restrict__ part /*[RED_BLOCKS][2*width]*/) {
  __shared__ double sh[32];
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  for (int j = 0; j < width; ++j) {
    double s = 0.0, q = 0.0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      const float v = feat(x, ld_x, width_raw, type_col, ld_type, r, j);
      s += (double)v;
      q += (double)(v * v);  // batched_data**2 is rounded to fp32 before the sum (normalization.py:62)
    }
    s = block_reduce_sum(s, sh);
    q = block_reduce_sum(q, sh);
    if (threadIdx.x == 0) { part[(size_t)blockIdx.x * 2 * width + j] = s; part[(size_t)blockIdx.x * 2 * width + width + j] = q; }
  }
}

__global__ void norm_final_kernel(const double* __restrict__ part, int nblocks, int width, int rows, float max_acc, float* acc_count,
                                  float* num_acc, float* acc_sum, float* acc_sq) {
  const int j = threadIdx.x;
  if (*num_acc >= max_acc) return;  // normalization.py:40 (uniform across the block)
  if (j < width) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblocks; ++b) { s += part[(size_t)b * 2 * width + j]; q += part[(size_t)b * 2 * width + width + j]; }
    acc_sum[j] = acc_sum[j] + (float)s;
    acc_sq[j] = acc_sq[j] + (float)q;
  }
  __syncthreads();
  if (j == 0) { *acc_count = *acc_count + (float)rows; *num_acc = *num_acc + 1.0f; }
}

__global__ void norm_apply_kernel(const float* __restrict__ x, int ld_x, int width_raw, const float* __restrict__ type_col, int ld_type,
                                  int width, int rows, const float* __restrict__ acc_count, const float* __restrict__ acc_sum,
                                  const float* __restrict__ acc_sq, int inverse, const float* __restrict__ addend, int ld_add,
                                  float* __restrict__ out, int ld_out) {
  __shared__ float s_mean[32], s_std[32];
  if (threadIdx.x < width) {
    const float cnt = fmaxf(*acc_count, 1.0f);              // normalization.py:73,81
    const float mean = acc_sum[threadIdx.x] / cnt;
    float sd = sqrtf(acc_sq[threadIdx.x] / cnt - mean * mean);
    if (sd < 1e-8f) sd = 1.0f;                              // :84 (NaN compares false, like torch)
    s_mean[threadIdx.x] = mean;
    s_std[threadIdx.x] = sd;
  }
  __syncthreads();
  const long long total = (long long)rows * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / width), j = (int)(i - (long long)r * width);
    const float v = feat(x, ld_x, width_raw, type_col, ld_type, r, j);
    float o;
    if (!inverse) o = (v - s_mean[j]) / s_std[j];           // :43-45
    else {
      o = v * s_std[j] + s_mean[j];                         // :52-54
      if (addend) o = o + addend[(size_t)r * ld_add + j];   // FVGN.py:318-321
    }
    out[(size_t)r * ld_out + j] = o;
  }
}

// ------------------------------------------------------------------------------------------------ a10
__global__ void __launch_bounds__(RED_THREADS) face_area_raw_kernel(const float* __restrict__ ftl, const float* __restrict__ area,
                                                                   const int32_t* __restrict__ nc, int F, float dt, float* __restrict__ raw,
                                                                   double* __restrict__ part /*[grid][2] or NULL*/) {
  __shared__ double sh[32];
  const int per = (F + gridDim.x - 1) / gridDim.x;
  const int f0 = blockIdx.x * per, f1 = min(F, f0 + per);
  double s = 0.0, q = 0.0;
  for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
    const float a = (area[nc[f]] + area[nc[(size_t)F + f]]) / 2.0f;  // FVGN.py:214-223
    const float v = ftl[2 * f + 1] * (dt / a);
    raw[f] = v;
    s += (double)v;
    q += (double)v * (double)v;
  }
  if (part) {
    s = block_reduce_sum(s, sh);
    q = block_reduce_sum(q, sh);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s; part[2 * blockIdx.x + 1] = q; }
  }
}

// stats[0]=mean, stats[1]=invstd used by the apply kernel
__global__ void bn_stats_kernel(const double* __restrict__ part, int nblocks, int F, int training, float* running_mean, float* running_var,
                                float* __restrict__ stats) {
  if (threadIdx.x != 0) return;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblocks; ++b) { s += part[2 * b]; q += part[2 * b + 1]; }
    const double mean = s / F;
    double var = q / F - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[0] = (float)mean;
    stats[1] = (float)(1.0 / sqrt(var + 1e-5));
    const double unbiased = (F > 1) ? var * ((double)F / (double)(F - 1)) : var;
    *running_mean = (float)(0.9 * (double)*running_mean + 0.1 * mean);  // momentum 0.1
    *running_var = (float)(0.9 * (double)*running_var + 0.1 * unbiased);
  } else {
    stats[0] = *running_mean;
    stats[1] = 1.0f / sqrtf(*running_var + 1e-5f);
  }
}

__global__ void bn_apply_kernel(const float* __restrict__ raw, int F, const float* __restrict__ stats, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  out[f] = (raw[f] - stats[0]) * stats[1] * (*w) + (*b);
}

// ------------------------------------------------------------------------------------------------ a11
__global__ void integrator_kernel(const float* __restrict__ dec, const float* __restrict__ unv, const float* __restrict__ fa,
                                  const int32_t* __restrict__ cfT, int C, float inv_rho, float rhs, float* __restrict__ du,
                                  float* __restrict__ cont) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float A0[3], A1[3], P0[3], P1[3], D0[3], D1[3], Ct[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int f = cfT[(size_t)k * C + c];
    const float u = dec[5 * (size_t)f], v = dec[5 * (size_t)f + 1], p = dec[5 * (size_t)f + 2];
    D0[k] = dec[5 * (size_t)f + 3];
    D1[k] = dec[5 * (size_t)f + 4];
    const float nx = unv[(size_t)c * 6 + 2 * k], ny = unv[(size_t)c * 6 + 2 * k + 1];
    const float a = fa[f];
    // uu_vu_face = [u*u, u*v, v*u, v*v]; chain_flux_dot_product: (uu*nx + uv*ny, vu*nx + vv*ny) * a   (:270-303)
    A0[k] = ((u * u) * nx + (u * v) * ny) * a;
    A1[k] = ((v * u) * nx + (v * v) * ny) * a;
    P0[k] = (p * nx) * a;                                     // (:313-318)
    P1[k] = (p * ny) * a;
    Ct[k] = (u * nx + v * ny) * a;                            // (:281-293)
  }
  const float A_x = (A0[0] + A0[1]) + A0[2], A_y = (A1[0] + A1[1]) + A1[2];
  const float P_x = (P0[0] + P0[1]) + P0[2], P_y = (P1[0] + P1[1]) + P1[2];
  const float D_x = (D0[0] + D0[1]) + D0[2], D_y = (D1[0] + D1[1]) + D1[2];
  du[2 * (size_t)c] = rhs * (-A_x - inv_rho * P_x) + D_x;      // (:320-331)
  du[2 * (size_t)c + 1] = rhs * (-A_y - inv_rho * P_y) + D_y;
  if (cont) cont[c] = (Ct[0] + Ct[1]) + Ct[2];
}

// tri / quad hybrid meshes: [4,C] face table (-1 = no 4th face), unv [C,4,2]; the first three terms are summed exactly as above, the 4th
// is added where it exists (oracle/fvgn_oracle.py::_integrator_poly)
__global__ void integrator_poly_kernel(const float* __restrict__ dec, const float* __restrict__ unv, const float* __restrict__ fa,
                                       const int32_t* __restrict__ cf4T, int C, float inv_rho, float rhs, float* __restrict__ du,
                                       float* __restrict__ cont) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float A0[4], A1[4], P0[4], P1[4], D0[4], D1[4], Ct[4];
  const bool quad = cf4T[3 * (size_t)C + c] >= 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = cf4T[(size_t)k * C + c];
    if (f < 0) { A0[k] = A1[k] = P0[k] = P1[k] = D0[k] = D1[k] = Ct[k] = 0.f; continue; }
    const float u = dec[5 * (size_t)f], v = dec[5 * (size_t)f + 1], p = dec[5 * (size_t)f + 2];
    D0[k] = dec[5 * (size_t)f + 3];
    D1[k] = dec[5 * (size_t)f + 4];
    const float nx = unv[(size_t)c * 8 + 2 * k], ny = unv[(size_t)c * 8 + 2 * k + 1];
    const float a = fa[f];
    A0[k] = ((u * u) * nx + (u * v) * ny) * a;
    A1[k] = ((v * u) * nx + (v * v) * ny) * a;
    P0[k] = (p * nx) * a;
    P1[k] = (p * ny) * a;
    Ct[k] = (u * nx + v * ny) * a;
  }
  float A_x = (A0[0] + A0[1]) + A0[2], A_y = (A1[0] + A1[1]) + A1[2];
  float P_x = (P0[0] + P0[1]) + P0[2], P_y = (P1[0] + P1[1]) + P1[2];
  float D_x = (D0[0] + D0[1]) + D0[2], D_y = (D1[0] + D1[1]) + D1[2];
  float Cs = (Ct[0] + Ct[1]) + Ct[2];
  if (quad) { A_x = A_x + A0[3]; A_y = A_y + A1[3]; P_x = P_x + P0[3]; P_y = P_y + P1[3]; D_x = D_x + D0[3]; D_y = D_y + D1[3]; Cs = Cs + Ct[3]; }
  du[2 * (size_t)c] = rhs * (-A_x - inv_rho * P_x) + D_x;
  du[2 * (size_t)c + 1] = rhs * (-A_y - inv_rho * P_y) + D_y;
  if (cont) cont[c] = Cs;
}

// ------------------------------------------------------------------------------------------------ a12
constexpr int LOSS_CHUNKS = 32;
__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
  return lo;
}

// grid (LOSS_CHUNKS, n_segments).  part[seg][chunk][6] = {mom_sum, n_cells, uv_sum, n_interior, p_sum, n_faces}
__global__ void __launch_bounds__(RED_THREADS) loss_partial_kernel(const float* __restrict__ pdu, const float* __restrict__ tdu,
                                                                  const float* __restrict__ pe, const float* __restrict__ tf,
                                                                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ cb,
                                                                  const int32_t* __restrict__ fb, int C, int F, int segmented,
                                                                  double* __restrict__ part) {
  __shared__ double sh[32];
  const int seg = blockIdx.y, chunk = blockIdx.x;
  int c0 = 0, c1 = C, f0 = 0, f1 = F;
  if (segmented) {  // batch vectors are sorted (PyG block-diagonal batching) => each graph is one contiguous range
    c0 = lower_bound_i32(cb, C, seg); c1 = lower_bound_i32(cb, C, seg + 1);
    f0 = lower_bound_i32(fb, F, seg); f1 = lower_bound_i32(fb, F, seg + 1);
  }
  double mom = 0.0, uv = 0.0, pp = 0.0, nint = 0.0;
  {
    const int per = (c1 - c0 + LOSS_CHUNKS - 1) / LOSS_CHUNKS;
    const int a = c0 + chunk * per, b = min(c1, a + per);
    for (int c = a + threadIdx.x; c < b; c += blockDim.x) {
      const float d0 = tdu[2 * (size_t)c] - pdu[2 * (size_t)c], d1 = tdu[2 * (size_t)c + 1] - pdu[2 * (size_t)c + 1];
      mom += (double)(d0 * d0) + (double)(d1 * d1);
    }
  }
  {
    const int per = (f1 - f0 + LOSS_CHUNKS - 1) / LOSS_CHUNKS;
    const int a = f0 + chunk * per, b = min(f1, a + per);
    for (int f = a + threadIdx.x; f < b; f += blockDim.x) {
      const float dp = tf[3 * (size_t)f + 2] - pe[5 * (size_t)f + 2];
      pp += (double)(dp * dp);
      if (mask[f]) {
        const float d0 = tf[3 * (size_t)f] - pe[5 * (size_t)f], d1 = tf[3 * (size_t)f + 1] - pe[5 * (size_t)f + 1];
        uv += (double)(d0 * d0) + (double)(d1 * d1);
        nint += 1.0;
      }
    }
  }
  mom = block_reduce_sum(mom, sh);
  uv = block_reduce_sum(uv, sh);
  pp = block_reduce_sum(pp, sh);
  nint = block_reduce_sum(nint, sh);
  if (threadIdx.x == 0) {
    double* o = part + ((size_t)seg * LOSS_CHUNKS + chunk) * 6;
    o[0] = mom; o[1] = (double)(c1 - c0); o[2] = uv; o[3] = nint; o[4] = pp; o[5] = (double)(f1 - f0);
  }
}

// one thread per segment, then thread 0 combines.  seg_terms[seg][4] = {mom, uv, p, total} kept for the backward.
__global__ void loss_final_kernel(const double* __restrict__ part, int n_seg, int mode, float w_mom, float w_uv, float w_p,
                                  float* __restrict__ out, float* __restrict__ seg_terms) {
  __shared__ float s_log[1024], s_m[1024], s_u[1024], s_p[1024];
  for (int seg = threadIdx.x; seg < n_seg; seg += blockDim.x) {
    double mom = 0, uv = 0, pp = 0, nint = 0;
    for (int k = 0; k < LOSS_CHUNKS; ++k) {
      const double* o = part + ((size_t)seg * LOSS_CHUNKS + k) * 6;
      mom += o[0]; uv += o[2]; nint += o[3]; pp += o[4];
    }
    const double nc = part[(size_t)seg * LOSS_CHUNKS * 6 + 1], nf = part[(size_t)seg * LOSS_CHUNKS * 6 + 5];
    float lm, lu, lp;
    if (mode == 0) {         // direct_mean_loss: MSELoss(mean) over all elements (loss_compute.py:92-109)
      lm = (float)(mom / fmax(nc * 2.0, 1.0)); lu = (float)(uv / fmax(nint * 2.0, 1.0)); lp = (float)(pp / fmax(nf, 1.0));
    } else if (mode == 1) {  // global_mean_sum: per-graph channel means summed over channels (:46-82)
      lm = (float)(mom / fmax(nc, 1.0)); lu = (float)(uv / fmax(nint, 1.0)); lp = (float)(pp / fmax(nf, 1.0));
    } else {                 // global_mean: face_uv is summed first (:52-53) then mean of 1 column; mom mean over 2 channels
      lm = (float)(mom / fmax(nc, 1.0)) / 2.0f; lu = (float)(uv / fmax(nint, 1.0)); lp = (float)(pp / fmax(nf, 1.0));
    }
    const float tot = w_mom * lm + w_uv * lu + w_p * lp;  // continuity term is 0.0 (App. D.2)
    s_log[seg] = logf(tot); s_m[seg] = lm; s_u[seg] = lu; s_p[seg] = lp;
    if (seg_terms) { seg_terms[4 * seg] = lm; seg_terms[4 * seg + 1] = lu; seg_terms[4 * seg + 2] = lp; seg_terms[4 * seg + 3] = tot; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0, m = 0, u = 0, p = 0;
    for (int s = 0; s < n_seg; ++s) { a += s_log[s]; m += s_m[s]; u += s_u[s]; p += s_p[s]; }
    out[0] = a / n_seg; out[1] = 0.0f; out[2] = m / n_seg; out[3] = u / n_seg; out[4] = p / n_seg;
  }
}


// ------------------------------------------------------------------------------------------------ backward of the head
// coef[seg][3] = d loss / d (sum of squared errors of momentum | face uv | face p) for that segment
__global__ void loss_coef_kernel(const double* __restrict__ part, const float* __restrict__ seg_terms, int n_seg, int mode, float w_mom,
                                 float w_uv, float w_p, const float* __restrict__ g_loss, float* __restrict__ coef) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= n_seg) return;
  double nint = 0;
  for (int k = 0; k < LOSS_CHUNKS; ++k) nint += part[((size_t)seg * LOSS_CHUNKS + k) * 6 + 3];
  const double nc = part[(size_t)seg * LOSS_CHUNKS * 6 + 1], nf = part[(size_t)seg * LOSS_CHUNKS * 6 + 5];
  const float g = (g_loss ? *g_loss : 1.0f) / ((float)n_seg * seg_terms[4 * seg + 3]);  // d mean(log tot) / d tot_seg
  const float dm = (mode == 1) ? 1.0f : 0.5f;                                            // mean over 2 channels vs sum
  const float du = (mode == 0) ? 0.5f : 1.0f;
  coef[3 * seg + 0] = g * w_mom * dm / (float)fmax(nc, 1.0);
  coef[3 * seg + 1] = g * w_uv * du / (float)fmax(nint, 1.0);
  coef[3 * seg + 2] = g * w_p / (float)fmax(nf, 1.0);
}

__global__ void loss_bwd_kernel(const float* __restrict__ pdu, const float* __restrict__ tdu, const float* __restrict__ pe,
                                const float* __restrict__ tf, const uint8_t* __restrict__ mask, const int32_t* __restrict__ cb,
                                const int32_t* __restrict__ fb, int C, int F, int segmented, const float* __restrict__ coef,
                                float* __restrict__ g_du, float* __restrict__ g_edge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float k = 2.0f * coef[3 * (segmented ? cb[i] : 0) + 0];
    g_du[2 * (size_t)i] = k * (pdu[2 * (size_t)i] - tdu[2 * (size_t)i]);
    g_du[2 * (size_t)i + 1] = k * (pdu[2 * (size_t)i + 1] - tdu[2 * (size_t)i + 1]);
  }
  if (i < F) {
    const int seg = segmented ? fb[i] : 0;
    const float ku = mask[i] ? 2.0f * coef[3 * seg + 1] : 0.0f, kp = 2.0f * coef[3 * seg + 2];
    float* g = g_edge + 5 * (size_t)i;
    g[0] = ku * (pe[5 * (size_t)i] - tf[3 * (size_t)i]);
    g[1] = ku * (pe[5 * (size_t)i + 1] - tf[3 * (size_t)i + 1]);
    g[2] = kp * (pe[5 * (size_t)i + 2] - tf[3 * (size_t)i + 2]);
    g[3] = 0.0f;
    g[4] = 0.0f;
  }
}

// transpose of the Intergrator's three gathers: one thread per face walks its incident (cell, k) pairs in CSR order
__global__ void integrator_bwd_kernel(const float* __restrict__ dec, const float* __restrict__ unv, const float* __restrict__ fa,
                                      const float* __restrict__ g_du, const int32_t* __restrict__ ptr, const int32_t* __restrict__ items,
                                      int C, int F, int kmax /* 3, or 4: [C,4,2] normals of a tri / quad mesh */, float inv_rho, float rhs,
                                      float* __restrict__ g_dec, float* __restrict__ g_fa) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float u = dec[5 * (size_t)f], v = dec[5 * (size_t)f + 1], p = dec[5 * (size_t)f + 2], a = fa[f];
  float gu = 0.f, gv = 0.f, gdx = 0.f, gdy = 0.f, ga = 0.f;
  for (int i = ptr[f]; i < ptr[f + 1]; ++i) {
    const int pos = items[i];
    const int k = pos / C, c = pos - k * C;
    const float gx = g_du[2 * (size_t)c], gy = g_du[2 * (size_t)c + 1];
    const float nx = unv[((size_t)c * kmax + k) * 2], ny = unv[((size_t)c * kmax + k) * 2 + 1];
    gu += -rhs * a * (gx * (2.0f * u * nx + v * ny) + gy * (v * nx));
    gv += -rhs * a * (gx * (u * ny) + gy * (u * nx + 2.0f * v * ny));
    gdx += gx;
    gdy += gy;
    ga += -rhs * (gx * ((u * u) * nx + (u * v) * ny) + gy * ((v * u) * nx + (v * v) * ny) + inv_rho * p * (gx * nx + gy * ny));
  }
  float* g = g_dec + 5 * (size_t)f;
  g[0] += gu;
  g[1] += gv;  // g[2] (pressure) gets no gradient from the integrator: p_face is detached (FVGN.py:231)
  g[3] += gdx;
  g[4] += gdy;
  g_fa[f] = ga;
}

// d BN weight = sum g * xhat ; d BN bias = sum g       (xhat from the saved raw feature and the (mean, invstd) used forward)
__global__ void __launch_bounds__(RED_THREADS) bn_bwd_partial_kernel(const float* __restrict__ g_fa, const float* __restrict__ raw,
                                                                    const float* __restrict__ stats, int F, double* __restrict__ part) {
  __shared__ double sh[32];
  const int per = (F + gridDim.x - 1) / gridDim.x;
  const int f0 = blockIdx.x * per, f1 = min(F, f0 + per);
  double sw = 0.0, sb = 0.0;
  for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
    const float xh = (raw[f] - stats[0]) * stats[1];
    sw += (double)(g_fa[f] * xh);
    sb += (double)g_fa[f];
  }
  sw = block_reduce_sum(sw, sh);
  sb = block_reduce_sum(sb, sh);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = sw; part[2 * blockIdx.x + 1] = sb; }
}
__global__ void bn_bwd_final_kernel(const double* __restrict__ part, int nb, float* __restrict__ g_wb) {
  if (threadIdx.x != 0) return;
  double sw = 0.0, sb = 0.0;
  for (int b = 0; b < nb; ++b) { sw += part[2 * b]; sb += part[2 * b + 1]; }
  g_wb[0] = (float)sw;
  g_wb[1] = (float)sb;
}

}  // namespace fvgn

using namespace fvgn;

extern "C" int fvgn_interp_nodes(const float* node_field, int ld, int col0, int width, const int32_t* cells_node_T,
                                 const float* cell_factor, const int32_t* face_node, int n_cells, int n_faces, float* on_cell,
                                 float* on_face, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(node_field && ld > 0 && col0 >= 0 && width > 0 && col0 + width <= ld, "bad field layout");
  FVGN_CHECK_ARG(!on_cell || (cells_node_T && cell_factor), "cell tables missing");
  FVGN_CHECK_ARG(!on_face || face_node, "face table missing");
  const int n = max(on_cell ? n_cells : 0, on_face ? n_faces : 0);
  if (n <= 0) return 0;
  interp_nodes_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(node_field, ld, col0, width, cells_node_T, cell_factor, face_node,
                                                                   n_cells, n_faces, on_cell, on_face);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_assemble_edge_attr(const float* cell_uv, int ld_uv, const float* centroid, const int32_t* neighbour_cell,
                                       const float* face_type_len, const float* bc_uv, int ld_bc, const uint8_t* flip_keep, int n_faces,
                                       float* edge_attr, uint8_t* mask_interior, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(cell_uv && centroid && neighbour_cell && face_type_len && bc_uv && edge_attr, "NULL pointer");
  if (n_faces <= 0) return 0;
  assemble_edge_attr_kernel<<<cdiv(n_faces, 256), 256, 0, as_stream(stream)>>>(cell_uv, ld_uv, centroid, neighbour_cell, face_type_len,
                                                                               bc_uv, ld_bc, flip_keep, n_faces, edge_attr, mask_interior);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_create_next_graph(const float* next_uv, const int32_t* neighbour_cell, const float* face_type_len, const float* bc_uv,
                                      int ld_bc, int n_cells, int n_faces, float* x, float* edge_attr, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(next_uv && neighbour_cell && face_type_len && bc_uv && x && edge_attr, "NULL pointer");
  const int n = max(n_cells, n_faces);
  if (n <= 0) return 0;
  next_graph_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(next_uv, neighbour_cell, face_type_len, bc_uv, ld_bc, n_cells, n_faces, x,
                                                                 edge_attr);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t fvgn_norm_workspace_floats(int width) { return (size_t)2 * RED_BLOCKS * 2 * (size_t)(width > 0 ? width : 1) + 8; }

extern "C" int fvgn_normalizer_accumulate(const float* x, int ld_x, int width_raw, const float* type_col, int ld_type, int one_hot, int rows,
                                          float max_accumulations, float* acc_count, float* num_accumulations, float* acc_sum,
                                          float* acc_sum_squared, float* workspace, fvgn_stream_t stream) {
  const int width = width_raw + one_hot;
  FVGN_CHECK_ARG(x && acc_count && num_accumulations && acc_sum && acc_sum_squared && workspace, "NULL pointer");
  FVGN_CHECK_ARG(width >= 1 && width <= 32 && rows > 0, "width must be in [1,32], rows > 0");
  FVGN_CHECK_ARG(one_hot == 0 || type_col, "type column missing");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "workspace must be 8B aligned");
  double* part = reinterpret_cast<double*>(workspace);
  const int nb = min(RED_BLOCKS, cdiv(rows, RED_THREADS));
  norm_partial_kernel<<<nb, RED_THREADS, 0, as_stream(stream)>>>(x, ld_x, width_raw, type_col, ld_type, width, rows, part);
  FVGN_LAUNCH_CHECK();
  norm_final_kernel<<<1, 32, 0, as_stream(stream)>>>(part, nb, width, rows, max_accumulations, acc_count, num_accumulations, acc_sum,
                                                     acc_sum_squared);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_normalizer_apply(const float* x, int ld_x, int width_raw, const float* type_col, int ld_type, int one_hot, int rows,
                                     const float* acc_count, const float* acc_sum, const float* acc_sum_squared, int inverse,
                                     const float* addend, int ld_addend, float* out, int ld_out, fvgn_stream_t stream) {
  const int width = width_raw + one_hot;
  FVGN_CHECK_ARG(x && acc_count && acc_sum && acc_sum_squared && out, "NULL pointer");
  FVGN_CHECK_ARG(width >= 1 && width <= 32 && ld_out >= width, "width must be in [1,32]");
  FVGN_CHECK_ARG(one_hot == 0 || type_col, "type column missing");
  if (rows <= 0) return 0;
  const long long total = (long long)rows * width;
  long long nbl = (total + 255) / 256; const int nb = (int)(nbl < 148LL * 16 ? nbl : 148LL * 16);
  norm_apply_kernel<<<nb, 256, 0, as_stream(stream)>>>(x, ld_x, width_raw, type_col, ld_type, width, rows, acc_count, acc_sum, acc_sum_squared,
                                                       inverse, addend, ld_addend, out, ld_out);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_face_area_bn_fwd(const float* face_type_len, const float* cell_area, const int32_t* neighbour_cell, int n_faces, float dt,
                                     int training, const float* bn_weight, const float* bn_bias, float* running_mean, float* running_var,
                                     float* face_area, float* raw, float* batch_stats, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(face_type_len && cell_area && neighbour_cell && bn_weight && bn_bias && running_mean && running_var && face_area && workspace,
                 "NULL pointer");
  FVGN_CHECK_ARG(n_faces > 0, "no faces");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "workspace must be 8B aligned");
  cudaStream_t st = as_stream(stream);
  double* part = reinterpret_cast<double*>(workspace);
  float* stats = batch_stats ? batch_stats : reinterpret_cast<float*>(part + 2 * RED_BLOCKS);
  float* rawbuf = raw ? raw : face_area;  // in-place when the caller does not keep the raw feature
  const int nb = min(RED_BLOCKS, cdiv(n_faces, RED_THREADS));
  FVGN_CHECK_ARG(training >= 0 && training <= 2, "training must be 0 (running stats), 1 (batch stats) or 2 (caller-provided stats)");
  FVGN_CHECK_ARG(training != 2 || batch_stats, "training == 2 needs batch_stats = [mean, invstd]");
  face_area_raw_kernel<<<nb, RED_THREADS, 0, st>>>(face_type_len, cell_area, neighbour_cell, n_faces, dt, rawbuf, training == 1 ? part : nullptr);
  FVGN_LAUNCH_CHECK();
  if (training != 2) {
    bn_stats_kernel<<<1, 32, 0, st>>>(part, nb, n_faces, training, running_mean, running_var, stats);
    FVGN_LAUNCH_CHECK();
  }
  bn_apply_kernel<<<cdiv(n_faces, 256), 256, 0, st>>>(rawbuf, n_faces, stats, bn_weight, bn_bias, face_area);
  FVGN_LAUNCH_CHECK();
  return 0;
}

namespace fvgn {
__global__ void sum_partials2_kernel(const double* __restrict__ part, int nb, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nb; ++b) { s += part[2 * b]; q += part[2 * b + 1]; }
  out[0] = s; out[1] = q;
}
}  // namespace fvgn

extern "C" int fvgn_face_area_sums(const float* face_type_len, const float* cell_area, const int32_t* neighbour_cell, int n_faces, float dt,
                                   float* raw, double* sums, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(face_type_len && cell_area && neighbour_cell && raw && sums && workspace && n_faces > 0, "NULL pointer");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0 && (reinterpret_cast<uintptr_t>(sums) & 7) == 0, "8B alignment");
  double* part = reinterpret_cast<double*>(workspace);
  const int nb = min(RED_BLOCKS, cdiv(n_faces, RED_THREADS));
  face_area_raw_kernel<<<nb, RED_THREADS, 0, as_stream(stream)>>>(face_type_len, cell_area, neighbour_cell, n_faces, dt, raw, part);
  FVGN_LAUNCH_CHECK();
  sum_partials2_kernel<<<1, 32, 0, as_stream(stream)>>>(part, nb, sums);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_integrator_fwd(const float* decoded, const float* unv, const float* face_area, const int32_t* cells_face_T, int n_cells,
                                   float rho, float rhs_coef, float* delta_u, float* continuity, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(decoded && unv && face_area && cells_face_T && delta_u, "NULL pointer");
  FVGN_CHECK_ARG(rho != 0.f, "rho must be non-zero");
  if (n_cells <= 0) return 0;
  integrator_kernel<<<cdiv(n_cells, 256), 256, 0, as_stream(stream)>>>(decoded, unv, face_area, cells_face_T, n_cells, 1.0f / rho, rhs_coef,
                                                                       delta_u, continuity);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_integrator_fwd_poly(const float* decoded, const float* unv4, const float* face_area, const int32_t* cells_face4_T, int n_cells,
                                        float rho, float rhs_coef, float* delta_u, float* continuity, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(decoded && unv4 && face_area && cells_face4_T && delta_u, "NULL pointer");
  FVGN_CHECK_ARG(rho != 0.f, "rho must be non-zero");
  if (n_cells <= 0) return 0;
  integrator_poly_kernel<<<cdiv(n_cells, 256), 256, 0, as_stream(stream)>>>(decoded, unv4, face_area, cells_face4_T, n_cells, 1.0f / rho, rhs_coef,
                                                                            delta_u, continuity);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_interp_cells_poly(const float* node_field, int ld, int col0, int width, const int32_t* cells_node4_T, const float* cell_factor4,
                                      int n_cells, float* on_cell, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(node_field && cells_node4_T && cell_factor4 && on_cell, "NULL pointer");
  FVGN_CHECK_ARG(width >= 1 && col0 >= 0 && ld >= col0 + width, "bad column range");
  if (n_cells <= 0) return 0;
  interp_cells_poly_kernel<<<cdiv(n_cells, 256), 256, 0, as_stream(stream)>>>(node_field, ld, col0, width, cells_node4_T, cell_factor4, n_cells,
                                                                              on_cell);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t fvgn_loss_workspace_floats(int n_graphs) {
  const size_t g = n_graphs > 0 ? n_graphs : 1;
  return 2 * (g * LOSS_CHUNKS * 6) + 4 * g + 8;
}

extern "C" int fvgn_loss_partials(const float* pred_delta_u, const float* target_delta_u, const float* pred_edge, const float* target_face,
                                  const uint8_t* mask_interior, const int32_t* cell_batch, const int32_t* face_batch, int n_cells, int n_faces,
                                  int n_graphs, int mode, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(pred_delta_u && target_delta_u && pred_edge && target_face && mask_interior && workspace, "NULL pointer");
  FVGN_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0,1,2");
  FVGN_CHECK_ARG(mode == 0 || (cell_batch && face_batch && n_graphs >= 1 && n_graphs <= 1024), "global_mean needs batch vectors, 1..1024 graphs");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "workspace must be 8B aligned");
  const int n_seg = (mode == 0) ? 1 : n_graphs;
  double* part = reinterpret_cast<double*>(workspace);
  loss_partial_kernel<<<dim3(LOSS_CHUNKS, n_seg), RED_THREADS, 0, as_stream(stream)>>>(pred_delta_u, target_delta_u, pred_edge, target_face,
                                                                                      mask_interior, cell_batch, face_batch, n_cells, n_faces,
                                                                                      mode != 0, part);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_loss_finish(int n_graphs, int mode, float w_mom, float w_uv, float w_p, float* out, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(out && workspace, "NULL pointer");
  FVGN_CHECK_ARG(mode >= 0 && mode <= 2 && (mode == 0 || (n_graphs >= 1 && n_graphs <= 1024)), "mode must be 0,1,2; 1..1024 graphs");
  const int n_seg = (mode == 0) ? 1 : n_graphs;
  double* part = reinterpret_cast<double*>(workspace);
  float* seg_terms = reinterpret_cast<float*>(part + (size_t)n_seg * LOSS_CHUNKS * 6);
  loss_final_kernel<<<1, 256, 0, as_stream(stream)>>>(part, n_seg, mode, w_mom, w_uv, w_p, out, seg_terms);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_loss_fwd(const float* pred_delta_u, const float* target_delta_u, const float* pred_edge, const float* target_face,
                             const uint8_t* mask_interior, const int32_t* cell_batch, const int32_t* face_batch, int n_cells, int n_faces,
                             int n_graphs, int mode, float w_mom, float w_uv, float w_p, float* out, float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(out != nullptr, "NULL pointer");
  if (int e = fvgn_loss_partials(pred_delta_u, target_delta_u, pred_edge, target_face, mask_interior, cell_batch, face_batch, n_cells, n_faces,
                                 n_graphs, mode, workspace, stream)) return e;
  return fvgn_loss_finish(n_graphs, mode, w_mom, w_uv, w_p, out, workspace, stream);
}

extern "C" int fvgn_loss_bwd(const float* pred_delta_u, const float* target_delta_u, const float* pred_edge, const float* target_face,
                             const uint8_t* mask_interior, const int32_t* cell_batch, const int32_t* face_batch, int n_cells, int n_faces,
                             int n_graphs, int mode, float w_mom, float w_uv, float w_p, const float* g_loss, const float* fwd_workspace,
                             float* g_delta_u, float* g_edge, float* coef_scratch, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(pred_delta_u && target_delta_u && pred_edge && target_face && mask_interior && fwd_workspace && g_delta_u && g_edge && coef_scratch,
                 "NULL pointer");
  FVGN_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0,1,2");
  FVGN_CHECK_ARG(mode == 0 || (cell_batch && face_batch && n_graphs >= 1 && n_graphs <= 1024), "global_mean needs batch vectors");
  const int n_seg = (mode == 0) ? 1 : n_graphs;
  const double* part = reinterpret_cast<const double*>(fwd_workspace);
  const float* seg_terms = reinterpret_cast<const float*>(part + (size_t)n_seg * LOSS_CHUNKS * 6);
  cudaStream_t st = as_stream(stream);
  loss_coef_kernel<<<cdiv(n_seg, 128), 128, 0, st>>>(part, seg_terms, n_seg, mode, w_mom, w_uv, w_p, g_loss, coef_scratch);
  FVGN_LAUNCH_CHECK();
  const int n = max(n_cells, n_faces);
  loss_bwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(pred_delta_u, target_delta_u, pred_edge, target_face, mask_interior, cell_batch, face_batch,
                                                 n_cells, n_faces, mode != 0, coef_scratch, g_delta_u, g_edge);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_integrator_bwd(const float* decoded, const float* unv, const float* face_area, const float* g_delta_u,
                                   const int32_t* face_ptr, const int32_t* face_items, int n_cells, int n_faces, float rho, float rhs_coef,
                                   float* g_decoded, float* g_face_area, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(decoded && unv && face_area && g_delta_u && face_ptr && face_items && g_decoded && g_face_area, "NULL pointer");
  FVGN_CHECK_ARG(rho != 0.f, "rho must be non-zero");
  if (n_faces <= 0) return 0;
  integrator_bwd_kernel<<<cdiv(n_faces, 256), 256, 0, as_stream(stream)>>>(decoded, unv, face_area, g_delta_u, face_ptr, face_items, n_cells,
                                                                           n_faces, 3, 1.0f / rho, rhs_coef, g_decoded, g_face_area);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_integrator_bwd_poly(const float* decoded, const float* unv4, const float* face_area, const float* g_delta_u,
                                        const int32_t* face_ptr, const int32_t* face_items, int n_cells, int n_faces, float rho, float rhs_coef,
                                        float* g_decoded, float* g_face_area, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(decoded && unv4 && face_area && g_delta_u && face_ptr && face_items && g_decoded && g_face_area, "NULL pointer");
  FVGN_CHECK_ARG(rho != 0.f, "rho must be non-zero");
  if (n_faces <= 0) return 0;
  integrator_bwd_kernel<<<cdiv(n_faces, 256), 256, 0, as_stream(stream)>>>(decoded, unv4, face_area, g_delta_u, face_ptr, face_items, n_cells,
                                                                           n_faces, 4, 1.0f / rho, rhs_coef, g_decoded, g_face_area);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_face_area_bn_bwd(const float* g_face_area, const float* raw, const float* batch_stats, int n_faces, float* g_weight_bias,
                                     float* workspace, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(g_face_area && raw && batch_stats && g_weight_bias && workspace && n_faces > 0, "NULL pointer");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "workspace must be 8B aligned");
  double* part = reinterpret_cast<double*>(workspace);
  const int nb = min(RED_BLOCKS, cdiv(n_faces, RED_THREADS));
  bn_bwd_partial_kernel<<<nb, RED_THREADS, 0, as_stream(stream)>>>(g_face_area, raw, batch_stats, n_faces, part);
  FVGN_LAUNCH_CHECK();
  bn_bwd_final_kernel<<<1, 32, 0, as_stream(stream)>>>(part, nb, g_weight_bias);
  FVGN_LAUNCH_CHECK();
  return 0;
}
