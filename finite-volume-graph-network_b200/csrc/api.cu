// api.cu — C-ABI entry points of libfvgn_b200.so that dispatch to the MLP kernels, plus error/device plumbing.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace fvgn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = cached[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// mlp_simt.cu
int simt_mlp_rows(const fvgn_mlp_t*, const float*, int, int, float*, int, cudaStream_t);
int simt_cell_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, cudaStream_t);
int simt_edge_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, cudaStream_t);
int node_aggregate(const float*, const int32_t*, const int32_t*, int, int, float*, cudaStream_t);
// mlp_tc.cu (tcgen05)
int tc_mlp_rows(const fvgn_mlp_t*, const float*, int, int, float*, int, int, cudaStream_t);
int tc_cell_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, int, cudaStream_t, uint16_t* = nullptr);
int tc_edge_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, int, cudaStream_t, const uint16_t* = nullptr);

// mlp_tc3.cu (tcgen05, 3-way split bf16 ~ fp32)
int tc3_mlp_rows(const fvgn_mlp_t*, const float*, int, int, float*, int, cudaStream_t);
int tc3_cell_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, cudaStream_t);
int tc3_edge_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, cudaStream_t);
int tc3_cell_mean_split(const float*, const int32_t*, int, void*, cudaStream_t);
int tc3_cell_mean(const float*, const int32_t*, int, float*, cudaStream_t);
int tc3_cell_mean_poly(const float*, const int32_t*, int, float*, void*, cudaStream_t);
int tc3_cell_block_sh(const fvgn_mlp_t*, float*, const void*, int, void*, cudaStream_t);
int tc3_edge_block_sh(const fvgn_mlp_t*, const void*, float*, const int32_t*, int, cudaStream_t);
int tc3_status_read(unsigned int*, int);
int tc3_pair_enable(int);
size_t tc3_packed_bytes(int k_in);
int tc3_pack(const fvgn_mlp_t*, void*, size_t, cudaStream_t);
int tc3_pack_many(const fvgn_mlp_t*, int, cudaStream_t);
size_t tc3_packed_bwd_bytes(int k_in);
int tc3_pack_bwd(const fvgn_mlp_t*, void*, size_t, cudaStream_t);
int tc3_pack_bwd_many(const fvgn_mlp_t*, int, cudaStream_t);
int tc3_ln_bwd(const float*, int, const float*, const float*, int, float*, float*, cudaStream_t);
size_t tc3_ln_bwd_sums_floats(int rows);
size_t tc3_wgrad_workspace_floats(int k_in);
int tc3_wgrad(const fvgn_mlp_t*, int, const float*, int, const float*, const float*, const int32_t*, int, const float*, int, const float*,
              const float*, float*, float*, cudaStream_t);
int tc3_dgrad(const fvgn_mlp_t*, const float*, int, int, float*, float*, int, const float*, int, cudaStream_t);

size_t tc_packed_bytes(int k_in);
int tc_pack(const fvgn_mlp_t*, void*, size_t, cudaStream_t);

static int check_device() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      set_error("no CUDA device (libfvgn_b200 has no CPU fallback)");
      return FVGN_E_NO_DEVICE;
    }
    if (major != 10) {
      set_error("device compute capability %d.x is not sm_100 (this library is built for sm_100a only)", major);
      return FVGN_E_NO_DEVICE;
    }
    ok = 1;
  }
  return 0;
}

}  // namespace fvgn

using namespace fvgn;

extern "C" int fvgn_abi_version(void) { return FVGN_ABI_VERSION; }
extern "C" const char* fvgn_last_error(void) { return g_err; }

extern "C" int fvgn_x3_pair_enable(int on) { return tc3_pair_enable(on); }

extern "C" int fvgn_status_read(unsigned int* status_host, int clear) { return tc3_status_read(status_host, clear); }

extern "C" int fvgn_device_check(int* sm, int* major, int* minor) {
  int dev = 0, a = 0, b = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return FVGN_E_NO_DEVICE; }
  cudaDeviceGetAttribute(&a, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (sm) *sm = n;
  if (major) *major = a;
  if (minor) *minor = b;
  return check_device();
}

extern "C" size_t fvgn_mlp_packed_bytes(int k_in) { return tc_packed_bytes(k_in); }

extern "C" int fvgn_mlp_pack_bf16(const fvgn_mlp_t* mlp, void* packed, size_t packed_bytes, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc_pack(mlp, packed, packed_bytes, as_stream(stream));
}

extern "C" size_t fvgn_mlp_packed_x3_bytes(int k_in) { return tc3_packed_bytes(k_in); }

extern "C" int fvgn_mlp_pack_bf16x3(const fvgn_mlp_t* mlp, void* packed, size_t packed_bytes, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_pack(mlp, packed, packed_bytes, as_stream(stream));
}

extern "C" int fvgn_mlp_pack_bf16x3_many(const fvgn_mlp_t* mlps, int n_mlps, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_pack_many(mlps, n_mlps, as_stream(stream));
}

extern "C" int fvgn_mlp_rows_fwd(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, float* out, int ld_out, int math_mode,
                                 fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  if (math_mode == FVGN_MATH_FP32_SIMT) return simt_mlp_rows(mlp, in, ld_in, rows, out, ld_out, as_stream(stream));
  if (math_mode == FVGN_MATH_BF16_TC) return tc_mlp_rows(mlp, in, ld_in, rows, out, ld_out, math_mode, as_stream(stream));
  if (math_mode == FVGN_MATH_BF16X3_TC) return tc3_mlp_rows(mlp, in, ld_in, rows, out, ld_out, as_stream(stream));
  set_error("unknown math_mode %d", math_mode);
  return FVGN_E_BADARG;
}

extern "C" int fvgn_node_aggregate_fwd(const float* e, const int32_t* node_ptr, const int32_t* node_items, int n_nodes, int n_faces,
                                       float* node_agg, fvgn_stream_t stream) {
  if (int er = check_device()) return er;
  return node_aggregate(e, node_ptr, node_items, n_nodes, n_faces, node_agg, as_stream(stream));
}

extern "C" int fvgn_cell_block_fwd(const fvgn_mlp_t* mlp, const float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells,
                                   float* x_prime, float* x_out, int math_mode, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  if (math_mode == FVGN_MATH_FP32_SIMT) return simt_cell_block(mlp, x, node_agg, cells_node_T, n_cells, x_prime, x_out, as_stream(stream));
  if (math_mode == FVGN_MATH_BF16_TC) return tc_cell_block(mlp, x, node_agg, cells_node_T, n_cells, x_prime, x_out, math_mode, as_stream(stream));
  if (math_mode == FVGN_MATH_BF16X3_TC) return tc3_cell_block(mlp, x, node_agg, cells_node_T, n_cells, x_prime, x_out, as_stream(stream));
  set_error("unknown math_mode %d", math_mode);
  return FVGN_E_BADARG;
}

extern "C" int fvgn_edge_block_fwd(const fvgn_mlp_t* mlp, const float* x_prime, const float* e, const int32_t* neighbour_cell, int n_faces,
                                   float* e_prime, float* e_out, int math_mode, fvgn_stream_t stream) {
  if (int er = check_device()) return er;
  if (math_mode == FVGN_MATH_FP32_SIMT) return simt_edge_block(mlp, x_prime, e, neighbour_cell, n_faces, e_prime, e_out, as_stream(stream));
  if (math_mode == FVGN_MATH_BF16_TC) return tc_edge_block(mlp, x_prime, e, neighbour_cell, n_faces, e_prime, e_out, math_mode, as_stream(stream));
  if (math_mode == FVGN_MATH_BF16X3_TC) return tc3_edge_block(mlp, x_prime, e, neighbour_cell, n_faces, e_prime, e_out, as_stream(stream));
  set_error("unknown math_mode %d", math_mode);
  return FVGN_E_BADARG;
}

extern "C" int fvgn_cell_block_fwd_xbf16(const fvgn_mlp_t* mlp, float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells,
                                         uint16_t* x_prime_bf16, int math_mode, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  if (math_mode != FVGN_MATH_BF16_TC) { set_error("fvgn_cell_block_fwd_xbf16: tensor-core math mode only"); return FVGN_E_BADARG; }
  if (!x_prime_bf16) { set_error("fvgn_cell_block_fwd_xbf16: x_prime_bf16 is NULL"); return FVGN_E_BADARG; }
  return tc_cell_block(mlp, x, node_agg, cells_node_T, n_cells, nullptr, x, math_mode, as_stream(stream), x_prime_bf16);
}

extern "C" int fvgn_edge_block_fwd_xbf16(const fvgn_mlp_t* mlp, const uint16_t* x_prime_bf16, float* e, const int32_t* neighbour_cell,
                                         int n_faces, int math_mode, fvgn_stream_t stream) {
  if (int er = check_device()) return er;
  if (math_mode != FVGN_MATH_BF16_TC) { set_error("fvgn_edge_block_fwd_xbf16: tensor-core math mode only"); return FVGN_E_BADARG; }
  if (!x_prime_bf16) { set_error("fvgn_edge_block_fwd_xbf16: x_prime_bf16 is NULL"); return FVGN_E_BADARG; }
  return tc_edge_block(mlp, nullptr, e, neighbour_cell, n_faces, nullptr, e, math_mode, as_stream(stream), x_prime_bf16);
}

extern "C" int fvgn_cell_mean_split(const float* node_agg, const int32_t* cells_node_T, int n_cells, void* cell_mean_split, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_cell_mean_split(node_agg, cells_node_T, n_cells, cell_mean_split, as_stream(stream));
}

extern "C" int fvgn_cell_mean(const float* node_agg, const int32_t* cells_node_T, int n_cells, float* cell_mean, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_cell_mean(node_agg, cells_node_T, n_cells, cell_mean, as_stream(stream));
}

extern "C" int fvgn_cell_mean_poly(const float* node_agg, const int32_t* cells_node4_T, int n_cells, float* cell_mean, void* cell_mean_split,
                                   fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_cell_mean_poly(node_agg, cells_node4_T, n_cells, cell_mean, cell_mean_split, as_stream(stream));
}

extern "C" int fvgn_cell_block_fwd_split(const fvgn_mlp_t* mlp, float* x, const void* cell_mean_split, int n_cells, void* x_prime_split,
                                         fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_cell_block_sh(mlp, x, cell_mean_split, n_cells, x_prime_split, as_stream(stream));
}

extern "C" int fvgn_edge_block_fwd_split(const fvgn_mlp_t* mlp, const void* x_prime_split, float* e, const int32_t* neighbour_cell,
                                         int n_faces, fvgn_stream_t stream) {
  if (int er = check_device()) return er;
  return tc3_edge_block_sh(mlp, x_prime_split, e, neighbour_cell, n_faces, as_stream(stream));
}

extern "C" size_t fvgn_mlp_packed_x3_bwd_bytes(int k_in) { return tc3_packed_bwd_bytes(k_in); }

extern "C" int fvgn_mlp_pack_bf16x3_bwd_many(const fvgn_mlp_t* mlps, int n_mlps, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_pack_bwd_many(mlps, n_mlps, as_stream(stream));
}

extern "C" int fvgn_mlp_pack_bf16x3_bwd(const fvgn_mlp_t* mlp, void* packed, size_t packed_bytes, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_pack_bwd(mlp, packed, packed_bytes, as_stream(stream));
}

extern "C" size_t fvgn_ln_bwd_sums_floats(int rows) { return tc3_ln_bwd_sums_floats(rows); }

extern "C" int fvgn_ln_bwd(const float* d_out, int ld_dout, const float* z_save, const float* ln_w, int rows, float* dz, float* gxhat,
                           fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_ln_bwd(d_out, ld_dout, z_save, ln_w, rows, dz, gxhat, as_stream(stream));
}

extern "C" size_t fvgn_mlp_wgrad_x3_workspace_floats(int k_in) { return tc3_wgrad_workspace_floats(k_in); }

extern "C" int fvgn_mlp_wgrad_x3(const fvgn_mlp_t* mlp, int mode, const float* in, int ld_in, const float* src0, const float* src1,
                                 const int32_t* idx, int rows, const float* d_out, int ld_dout, const float* dz, const float* gxhat,
                                 float* grads, float* workspace, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_wgrad(mlp, mode, in, ld_in, src0, src1, idx, rows, d_out, ld_dout, dz, gxhat, grads, workspace, as_stream(stream));
}

extern "C" int fvgn_mlp_dgrad_x3(const fvgn_mlp_t* mlp, const float* dz, int ld_dz, int rows, float* da, float* dA, int ld_dA,
                                 const float* dA_base, int ld_base, fvgn_stream_t stream) {
  if (int e = check_device()) return e;
  return tc3_dgrad(mlp, dz, ld_dz, rows, da, dA, ld_dA, dA_base, ld_base, as_stream(stream));
}
