// mlp_simt.cu — strict-fp32 (FFMA) fused gather -> 3-layer MLP -> LayerNorm -> residual kernels.
//
// One CTA = 64 rows (cells or faces) x the full 128-wide MLP of build_mlp (EncoderProcesserDecoder.py:12-48):
//   A tile   [64][K+4]  fp32 in shared memory, filled by the mode's gather (never materialised in HBM; the reference
//                       materialises cat([x, cell_agg]) [C,192] and cat([x'_s, x'_r, e]) [F,384] every layer)
//   weights  streamed from L2 in 16-wide K chunks, transposed on the fly into Wt[k][n] (double buffered)
//   thread tile 4 rows x 8 cols (cols {4tc..4tc+3} U {64+4tc..}), 32 FFMA per 3 LDS.128
//   hidden activations ping-pong between two [64][132] shared buffers; LayerNorm = 16-lane shuffle reduction;
//   epilogue adds the GnBlock residual (EncoderProcesserDecoder.py:121-122) from the fp32 copy in the A tile.
// Modes: ROWS (Encoder / Decoder / generic build_mlp), CELL (CellBlock, GN_blocks.py:125-143), EDGE (EdgeBlock, :43-52).
#include "simt_tile.cuh"

namespace fvgn {

enum { MODE_ROWS = 0, MODE_CELL = 1, MODE_EDGE = 2 };

struct SimtParams {
  fvgn_mlp_t mlp;
  int rows;
  // ROWS
  const float* in; int ld_in;
  // CELL: x [C,128], node_agg [N,64], cells_node_T [3,C]
  // EDGE: x_prime [C,128] (gathered), e [F,128], neighbour_cell [2,F]
  const float* src0; const float* src1; const int32_t* idx;
  float* out0; int ld_out;  // MLP(+LN) output (x_prime / e_prime / rows out)
  float* out1;              // residual-added stream (x_out / e_out)
};

template <int MODE, int KPAD>
__global__ void __launch_bounds__(NT, 1) fused_mlp_simt_kernel(const SimtParams p) {
  constexpr int LDA = KPAD + 4;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                 // [BM][LDA]
  float* H1 = As + BM * LDA;        // [BM][HLD]
  float* H2 = H1 + BM * HLD;        // [BM][HLD]
  float* Wt = H2 + BM * HLD;        // [2][KC][WLD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * BM;
  const int rows = p.rows;

  // ------------------------------------------------------------------ gather: build the A tile
  if (MODE == MODE_ROWS) {
    const int k_in = p.mlp.k_in;
    for (int idx = tid; idx < BM * KPAD; idx += NT) {
      const int r = idx / KPAD, k = idx - r * KPAD;
      const int g = row0 + r;
      float v = 0.f;
      if (g < rows && k < k_in) v = p.in[(size_t)g * p.ld_in + k];
      As[r * LDA + k] = v;
    }
  } else if (MODE == MODE_CELL) {
    const float* x = p.src0;
    const float* na = p.src1;
    const int C = rows;
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = row0 + warp * 8 + i;
      xv[i] = (c < C) ? *reinterpret_cast<const float4*>(x + (size_t)c * HID + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int hl = lane & 15, hw = lane >> 4;
    float4 ag[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = row0 + warp * 8 + i * 2 + hw;
      const float4 o = (c < C) ? cell_agg4(na, p.idx, (size_t)C, c, hl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      ag[i] = o;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(As + (warp * 8 + i) * LDA + lane * 4) = xv[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(As + (warp * 8 + i * 2 + hw) * LDA + 128 + hl * 4) = ag[i];
  } else {  // MODE_EDGE
    const float* xp = p.src0;
    const float* e = p.src1;
    const int F = rows;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 vs[4], vr[4], ve[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int f = row0 + warp * 8 + half * 4 + i;
        if (f < F) {
          const int s = __ldg(p.idx + f), r = __ldg(p.idx + (size_t)F + f);
          vs[i] = ldg4(xp + (size_t)s * HID + lane * 4);
          vr[i] = ldg4(xp + (size_t)r * HID + lane * 4);
          ve[i] = *reinterpret_cast<const float4*>(e + (size_t)f * HID + lane * 4);
        } else {
          vs[i] = vr[i] = ve[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* a = As + (warp * 8 + half * 4 + i) * LDA + lane * 4;
        *reinterpret_cast<float4*>(a) = vs[i];
        *reinterpret_cast<float4*>(a + 128) = vr[i];
        *reinterpret_cast<float4*>(a + 256) = ve[i];
      }
    }
  }
  // (visibility of the A tile is provided by the first __syncthreads inside gemm_layer)

  float acc[4][8];
  // ------------------------------------------------------------------ layer 1, 2
  gemm_layer<KPAD, LDA>(As, p.mlp.w1, p.mlp.k_in, p.mlp.k_in, HID, Wt, acc, tid);
  store_hidden(acc, p.mlp.b1, H1, tid);
  gemm_layer<HID, HLD>(H1, p.mlp.w2, HID, HID, HID, Wt, acc, tid);
  store_hidden(acc, p.mlp.b2, H2, tid);
  // ------------------------------------------------------------------ layer 3 + LayerNorm + residual
  const int n_out = p.mlp.n_out;
  gemm_layer<HID, HLD>(H2, p.mlp.w3, HID, HID, n_out, Wt, acc, tid);

  const int tr = tid >> 4, tc = tid & 15;
  int col[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { col[j] = tc * 4 + j; col[4 + j] = 64 + tc * 4 + j; }
  float bias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias[j] = (col[j] < n_out) ? __ldg(p.mlp.b3 + col[j]) : 0.f;
  const bool has_ln = (p.mlp.ln_w != nullptr);
  float g[8], bt[8];
  if (has_ln) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { g[j] = __ldg(p.mlp.ln_w + col[j]); bt[j] = __ldg(p.mlp.ln_b + col[j]); }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr * 4 + i;
    const int grow = row0 + r;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j] + bias[j];
    if (has_ln) {  // LayerNorm over the 128 columns of this row: 8 per thread x 16 lanes
      float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.0f / 128.0f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j] + bt[j];
    }
    if (grow < rows) {
      if (MODE == MODE_ROWS) {
        if (n_out == HID) {
          *reinterpret_cast<float4*>(p.out0 + (size_t)grow * p.ld_out + tc * 4) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(p.out0 + (size_t)grow * p.ld_out + 64 + tc * 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (col[j] < n_out) p.out0[(size_t)grow * p.ld_out + col[j]] = v[j];
        }
      } else {
        constexpr int RES = (MODE == MODE_CELL) ? 0 : 256;  // where the residual operand sits in the A tile
        if (p.out0) {
          *reinterpret_cast<float4*>(p.out0 + (size_t)grow * HID + tc * 4) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(p.out0 + (size_t)grow * HID + 64 + tc * 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (p.out1) {
          const float4 r0 = *reinterpret_cast<const float4*>(As + r * LDA + RES + tc * 4);
          const float4 r1 = *reinterpret_cast<const float4*>(As + r * LDA + RES + 64 + tc * 4);
          *reinterpret_cast<float4*>(p.out1 + (size_t)grow * HID + tc * 4) = make_float4(r0.x + v[0], r0.y + v[1], r0.z + v[2], r0.w + v[3]);
          *reinterpret_cast<float4*>(p.out1 + (size_t)grow * HID + 64 + tc * 4) = make_float4(r1.x + v[4], r1.y + v[5], r1.z + v[6], r1.w + v[7]);
        }
      }
    }
  }
}

template <int KPAD>
static constexpr size_t simt_smem_bytes() {
  return sizeof(float) * (size_t)(BM * (KPAD + 4) + 2 * BM * HLD + 2 * KC * WLD);
}

template <int MODE, int KPAD>
static int launch_simt(const SimtParams& p, cudaStream_t st) {
  if (p.rows <= 0) return 0;
  constexpr size_t smem = simt_smem_bytes<KPAD>();
  static ::fvgn::PerDeviceOnce configured;  // per (MODE,KPAD) instantiation
  if (configured.first()) {
    FVGN_CUDA(cudaFuncSetAttribute(fused_mlp_simt_kernel<MODE, KPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  fused_mlp_simt_kernel<MODE, KPAD><<<cdiv(p.rows, BM), NT, smem, st>>>(p);
  FVGN_LAUNCH_CHECK();
  return 0;
}

static int check_mlp(const fvgn_mlp_t* m, bool need_ln) {
  FVGN_CHECK_ARG(m != nullptr, "mlp is NULL");
  FVGN_CHECK_ARG(m->w1 && m->b1 && m->w2 && m->b2 && m->w3 && m->b3, "mlp weight pointer is NULL");
  FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384, "k_in must be in [1,384]");
  FVGN_CHECK_ARG(m->n_out >= 1 && m->n_out <= HID, "n_out must be in [1,128]");
  if (m->ln_w || need_ln) {
    FVGN_CHECK_ARG(m->ln_w && m->ln_b, "LayerNorm weight/bias missing");
    FVGN_CHECK_ARG(m->n_out == HID, "LayerNorm requires n_out == 128");
  }
  return 0;
}

int simt_mlp_rows(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, float* out, int ld_out, cudaStream_t st) {
  if (int e = check_mlp(mlp, false)) return e;
  FVGN_CHECK_ARG(rows >= 0 && (rows == 0 || (in && out)), "in/out NULL");
  FVGN_CHECK_ARG(ld_in >= mlp->k_in && ld_out >= mlp->n_out, "leading dimension too small");
  if (mlp->n_out == HID) FVGN_CHECK_ARG((ld_out % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "128-wide output must be 16B aligned");
  SimtParams p{};
  p.mlp = *mlp; p.rows = rows; p.in = in; p.ld_in = ld_in; p.out0 = out; p.ld_out = ld_out;
  if (mlp->k_in <= 16) return launch_simt<MODE_ROWS, 16>(p, st);
  if (mlp->k_in <= 128) return launch_simt<MODE_ROWS, 128>(p, st);
  if (mlp->k_in <= 192) return launch_simt<MODE_ROWS, 192>(p, st);
  return launch_simt<MODE_ROWS, 384>(p, st);
}

int simt_cell_block(const fvgn_mlp_t* mlp, const float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells,
                    float* x_prime, float* x_out, cudaStream_t st) {
  if (int e = check_mlp(mlp, true)) return e;
  FVGN_CHECK_ARG(mlp->k_in == 192, "CellBlock MLP must have k_in == 192 (128 + 64)");
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (x && node_agg)), "NULL input");  // cells_node_T == NULL: mp_times = 1
  SimtParams p{};
  p.mlp = *mlp; p.rows = n_cells; p.src0 = x; p.src1 = node_agg; p.idx = cells_node_T; p.out0 = x_prime; p.out1 = x_out; p.ld_out = HID;
  return launch_simt<MODE_CELL, 192>(p, st);
}

int simt_edge_block(const fvgn_mlp_t* mlp, const float* x_prime, const float* e, const int32_t* neighbour_cell, int n_faces,
                    float* e_prime, float* e_out, cudaStream_t st) {
  if (int er = check_mlp(mlp, true)) return er;
  FVGN_CHECK_ARG(mlp->k_in == 384, "EdgeBlock MLP must have k_in == 384 (3 x 128)");
  FVGN_CHECK_ARG(n_faces >= 0 && (n_faces == 0 || (x_prime && e && neighbour_cell)), "NULL input");
  SimtParams p{};
  p.mlp = *mlp; p.rows = n_faces; p.src0 = x_prime; p.src1 = e; p.idx = neighbour_cell; p.out0 = e_prime; p.out1 = e_out; p.ld_out = HID;
  return launch_simt<MODE_EDGE, 384>(p, st);
}

// ------------------------------------------------------------------------------------------------------------
// first message aggregation (GN_blocks.py:91,111-122): deterministic CSR segment sum, half-warp per node,
// 128-bit loads of the 64-float half rows, sequential fp32 adds in the reference's scatter order.
// items[i] = position in cat(face[0], face[1]): pos < F -> (f=pos, half 0) ; else (f=pos-F, half 1).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) node_aggregate_kernel(const float* __restrict__ e, const int32_t* __restrict__ ptr,
                                                             const int32_t* __restrict__ items, int n_nodes, int n_faces,
                                                             float* __restrict__ node_agg) {
  const int hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;  // half-warp id = node
  const int hl = threadIdx.x & 15;
  if (hw >= n_nodes) return;
  const int beg = __ldg(ptr + hw), end = __ldg(ptr + hw + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // batches of 8 incident face halves (a node of a triangulation has ~6): all index loads, then all row loads, in flight together —
  // three dependent memory round trips per node instead of two per face; the adds stay sequential in CSR order
  for (int i = beg; i < end; i += 8) {
    int pos[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) pos[u] = (i + u < end) ? __ldg(items + i + u) : -1;
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t off = (pos[u] < n_faces) ? (size_t)(pos[u] < 0 ? 0 : pos[u]) * HID : (size_t)(pos[u] - n_faces) * HID + 64;
      v[u] = (pos[u] >= 0) ? ldg4(e + off + hl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (pos[u] >= 0) {
        acc.x = __fadd_rn(acc.x, v[u].x); acc.y = __fadd_rn(acc.y, v[u].y);
        acc.z = __fadd_rn(acc.z, v[u].z); acc.w = __fadd_rn(acc.w, v[u].w);
      }
  }
  *reinterpret_cast<float4*>(node_agg + (size_t)hw * 64 + hl * 4) = acc;
}

int node_aggregate(const float* e, const int32_t* ptr, const int32_t* items, int n_nodes, int n_faces, float* node_agg, cudaStream_t st) {
  FVGN_CHECK_ARG(n_nodes >= 0 && n_faces >= 0, "negative size");
  if (n_nodes == 0) return 0;
  FVGN_CHECK_ARG(e && ptr && items && node_agg, "NULL pointer");
  const long long threads = (long long)n_nodes * 16;
  node_aggregate_kernel<<<cdiv(threads, 256), 256, 0, st>>>(e, ptr, items, n_nodes, n_faces, node_agg);
  FVGN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fvgn
