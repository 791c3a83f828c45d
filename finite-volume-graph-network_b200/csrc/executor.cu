// executor.cu — the processor's layer loops of ONE training step as two C calls (tensor-core training path).
//
// train.py:119-160 draws a new random batch every step, so the batch TOPOLOGY changes every step and a captured CUDA graph of the whole
// step cannot be replayed; the eager step then is bound by its ~1200 Python-driven launches (autograd.NetFunction: ~9 ctypes calls per
// GnBlock in the backward, each with its validators and output allocations).  These two entry points issue the same kernels, in the same
// order, on the same two streams, from C: per step the host side drops from ~200 ctypes calls to 2 (+ one per data-parallel reduce group),
// and all per-layer outputs live in a handful of [L, rows, cols] arrays / one scratch block the caller allocates once per step.
//   forward   fvgn_processor_train_fwd   = L x (node aggregate, per-cell mean, fused CellBlock, fused EdgeBlock), saving x', the per-cell
//                                          aggregate and every MLP's pre-activations (GN_blocks.py:16-145, EncoderProcesserDecoder.py:112-124)
//   backward  fvgn_processor_train_bwd   = for the layers [lo, hi) in reverse: LayerNorm backward, data-gradient chain, weight gradients
//                                          (side stream), the three CSR transposes of the gathers, gradient combination
// Same arithmetic as the per-call path: gradients are bit-identical (tests/test_gpu_executor.py).
#include "common.cuh"

namespace fvgn {
// mlp_simt.cu / mlp_tc3.cu / mlp_tc3_bwd.cu
int node_aggregate(const float*, const int32_t*, const int32_t*, int, int, float*, cudaStream_t);
int tc3_cell_mean(const float*, const int32_t*, int, float*, cudaStream_t);
int tc3_cell_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, cudaStream_t);
int tc3_edge_block(const fvgn_mlp_t*, const float*, const float*, const int32_t*, int, float*, float*, cudaStream_t);
int tc3_ln_bwd(const float*, int, const float*, const float*, int, float*, float*, cudaStream_t);
size_t tc3_ln_bwd_sums_floats(int rows);
size_t tc3_wgrad_workspace_floats(int k_in);
int tc3_wgrad(const fvgn_mlp_t*, int, const float*, int, const float*, const float*, const int32_t*, int, const float*, int, const float*,
              const float*, float*, float*, cudaStream_t);
int tc3_dgrad(const fvgn_mlp_t*, const float*, int, int, float*, float*, int, const float*, int, cudaStream_t);

static inline size_t up64(size_t floats) { return (floats + 63) & ~(size_t)63; }  // 256-byte granules

struct BwdRingSet {
  float *dz_e, *gxh_e, *da_e, *dA_e, *g_xp, *dz_c, *gxh_c, *da_c, *dA_c, *g_na, *ws_e, *ws_c;
};
struct BwdScratch {
  BwdRingSet set[3];
  float* g_e[3];
  size_t total;
};

static BwdScratch carve_bwd(float* base, int C, int F, int N) {
  BwdScratch s{};
  size_t o = 0;
  auto take = [&](size_t n) { float* p = base ? base + o : nullptr; o += up64(n); return p; };
  for (int k = 0; k < 3; ++k) {
    BwdRingSet& r = s.set[k];
    r.dz_e = take((size_t)F * 128); r.gxh_e = take(tc3_ln_bwd_sums_floats(F)); r.da_e = take((size_t)F * 256); r.dA_e = take((size_t)F * 384);
    r.g_xp = take((size_t)C * 128); r.dz_c = take((size_t)C * 128); r.gxh_c = take(tc3_ln_bwd_sums_floats(C)); r.da_c = take((size_t)C * 256);
    r.dA_c = take((size_t)C * 192); r.g_na = take((size_t)(N > C ? N : C) * 64);
    r.ws_e = take(tc3_wgrad_workspace_floats(384)); r.ws_c = take(tc3_wgrad_workspace_floats(192));
    s.g_e[k] = take((size_t)F * 128);
  }
  s.total = o;
  return s;
}
}  // namespace fvgn

using namespace fvgn;

static int check_tables(const fvgn_train_tables_t* t, int mp_times) {
  FVGN_CHECK_ARG(t != nullptr && t->n_cells > 0 && t->n_faces > 0 && t->n_nodes > 0, "empty mesh");
  FVGN_CHECK_ARG(t->neighbour_cell && t->cell_from_face_ptr && t->cell_from_face_items, "neighbour_cell / cell<-face CSR missing");
  if (mp_times >= 2) FVGN_CHECK_ARG(t->node_ptr && t->node_items && t->cells_node_T && t->face_node, "node tables missing (mp_times >= 2)");
  return 0;
}

extern "C" int fvgn_processor_train_fwd(const fvgn_mlp_t* cell_mlps, const fvgn_mlp_t* edge_mlps, int n_layers, int mp_times,
                                        const fvgn_train_tables_t* t, float* X, float* E, float* XP, float* NA, float* ZC, float* ZE,
                                        float* node_scratch, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(cell_mlps && edge_mlps && n_layers >= 0 && X && E && XP && NA && ZC && ZE, "NULL pointer");
  if (int e = check_tables(t, mp_times)) return e;
  FVGN_CHECK_ARG(mp_times < 2 || node_scratch, "node_scratch [N,64] missing");
  cudaStream_t st = as_stream(stream);
  const size_t C = t->n_cells, F = t->n_faces;
  for (int l = 0; l < n_layers; ++l) {
    const float* x = X + (size_t)l * C * 128;
    const float* e = E + (size_t)l * F * 128;
    float* x_new = X + (size_t)(l + 1) * C * 128;
    float* e_new = E + (size_t)(l + 1) * F * 128;
    float* xp = XP + (size_t)l * C * 128;
    float* na = NA + (size_t)l * C * 64;
    if (mp_times >= 2) {  // faces -> nodes -> cells (GN_blocks.py:110-129)
      if (int r = node_aggregate(e, t->node_ptr, t->node_items, t->n_nodes, t->n_faces, node_scratch, st)) return r;
      if (int r = tc3_cell_mean(node_scratch, t->cells_node_T, t->n_cells, na, st)) return r;
    } else {              // faces -> their two cells (:130-138)
      if (int r = node_aggregate(e, t->cell_from_face_ptr, t->cell_from_face_items, t->n_cells, t->n_faces, na, st)) return r;
    }
    fvgn_mlp_t cm = cell_mlps[l], em = edge_mlps[l];
    cm.z_save = ZC + (size_t)l * C * 384;
    em.z_save = ZE + (size_t)l * F * 384;
    if (int r = tc3_cell_block(&cm, x, na, nullptr, t->n_cells, xp, x_new, st)) return r;
    if (int r = tc3_edge_block(&em, xp, e, t->neighbour_cell, t->n_faces, nullptr, e_new, st)) return r;
  }
  return 0;
}

extern "C" size_t fvgn_processor_train_bwd_scratch_floats(int n_cells, int n_faces, int n_nodes) {
  return carve_bwd(nullptr, n_cells, n_faces, n_nodes).total;
}

// Layers hi-1 ... lo.  g_e_in [F,128] = d loss / d e_out of layer hi-1 (the latent after that layer); g_e_out [F,128] (a DIFFERENT
// buffer: the weight-gradient kernels still read g_e_in from the side stream) = d / d e_in of layer lo.  g_x_in [C,128] = d / d x_out of
// layer hi-1, or NULL (the last layer's cell latent feeds nothing else); g_x_out [C,128] = d / d x_in of layer lo.  grads[2 l] / grads[2 l + 1]: flat gradient buffers (fvgn_mlp_grad_floats) of the EdgeBlock / CellBlock
// MLP of layer l.  The weight-gradient kernels run on `side_stream` (forked from / joined into `stream` with events); when the call
// returns, `stream` has been made to wait for all of them.
extern "C" int fvgn_processor_train_bwd(const fvgn_mlp_t* cell_mlps, const fvgn_mlp_t* edge_mlps, int lo, int hi, int mp_times, int n_products,
                                        const fvgn_train_tables_t* t, const float* X, const float* E, const float* XP, const float* NA,
                                        const float* ZC, const float* ZE, const float* g_e_in, float* g_e_out, const float* g_x_in, float* g_x_out,
                                        float* const* grads,
                                        float* scratch, size_t scratch_floats, fvgn_stream_t side_stream, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(cell_mlps && edge_mlps && 0 <= lo && lo <= hi && X && E && XP && NA && ZC && ZE && g_e_in && g_e_out && g_x_out && grads && scratch,
                 "NULL pointer / bad layer range");
  FVGN_CHECK_ARG(g_e_in != g_e_out && g_x_in != g_x_out, "the gradient inputs and outputs must be different buffers");
  if (int e = check_tables(t, mp_times)) return e;
  FVGN_CHECK_ARG(mp_times < 2 || (t->node_from_cell_ptr && t->node_from_cell_items), "node<-cell CSR missing (mp_times >= 2)");
  const int Cn = t->n_cells, Fn = t->n_faces, Nn = t->n_nodes;
  const size_t C = Cn, F = Fn;
  BwdScratch s = carve_bwd(scratch, Cn, Fn, Nn);
  if (scratch_floats < s.total) { set_error("fvgn_processor_train_bwd: scratch too small: %zu < %zu floats", scratch_floats, s.total); return FVGN_E_WORKSPACE; }
  if (lo == hi) return 0;
  cudaStream_t st = as_stream(stream), side = as_stream(side_stream);
  const bool use_side = side != nullptr && side != st;
  cudaStream_t wst = use_side ? side : st;
  // events (created once per device, never destroyed: they may be referenced by a stream capture in progress):
  // fork = "main has produced the inputs of these weight gradients", done[k] = "side has finished the weight gradients of ring set k"
  static cudaEvent_t ev_pool[64][5];
  static bool ev_ready[64] = {};
  int devi = 0;
  cudaGetDevice(&devi);
  devi &= 63;
  if (use_side && !ev_ready[devi]) {
    for (auto& ev : ev_pool[devi]) FVGN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ev_ready[devi] = true;
  }
  cudaEvent_t* fork_e = ev_pool[devi];
  cudaEvent_t* done = ev_pool[devi] + 2;
  bool done_valid[3] = {false, false, false};
  int rc = 0;
#define EX(call) do { rc = (call); if (rc) goto cleanup; } while (0)
#define EXC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s -> %s", #call, cudaGetErrorString(e__)); rc = (int)e__; goto cleanup; } } while (0)
  {
    const float* g_e = g_e_in;
    const float* g_x = g_x_in;  // [C, ld_gx] or NULL
    int ld_gx = 128;
    int it = 0;
    for (int l = hi - 1; l >= lo; --l, ++it) {
      BwdRingSet& r = s.set[it % 3];
      // the ring set of iteration it - 3 is about to be overwritten: its weight gradients (side stream) were waited for at the end of
      // iteration it - 2 (one-stage delay, below)
      const float* x_in = X + (size_t)l * C * 128;
      const float* e_in = E + (size_t)l * F * 128;
      const float* xp = XP + (size_t)l * C * 128;
      const float* na = NA + (size_t)l * C * 64;
      fvgn_mlp_t em = edge_mlps[l], cm = cell_mlps[l];
      em.z_save = const_cast<float*>(ZE + (size_t)l * F * 384);
      cm.z_save = const_cast<float*>(ZC + (size_t)l * C * 384);
      if (em.n_products != 1) em.n_products = n_products;
      if (cm.n_products != 1) cm.n_products = n_products;
      // ---- EdgeBlock: LayerNorm', data gradients, weight gradients
      EX(tc3_ln_bwd(g_e, 128, em.z_save, em.ln_w, Fn, r.dz_e, r.gxh_e, st));
      EX(tc3_dgrad(&em, r.dz_e, 128, Fn, r.da_e, r.dA_e, 384, nullptr, 0, st));
      em.da_in = r.da_e;
      if (use_side) { EXC(cudaEventRecord(fork_e[0], st)); EXC(cudaStreamWaitEvent(side, fork_e[0], 0)); }
      EX(tc3_wgrad(&em, 2, nullptr, 0, xp, e_in, t->neighbour_cell, Fn, g_e, 128, r.dz_e, r.gxh_e, grads[2 * l], r.ws_e, wst));
      // ---- d x' = residual path + transpose of the x'[s], x'[r] gathers
      EX(fvgn_segment_sum_rows(r.dA_e, 384, Fn, 0, 128, 128, t->cell_from_face_ptr, t->cell_from_face_items, Cn, g_x, ld_gx, 1.0f, r.g_xp, 128, st));
      // ---- CellBlock
      EX(tc3_ln_bwd(r.g_xp, 128, cm.z_save, cm.ln_w, Cn, r.dz_c, r.gxh_c, st));
      EX(tc3_dgrad(&cm, r.dz_c, 128, Cn, r.da_c, r.dA_c, 192, g_x, ld_gx, st));
      cm.da_in = r.da_c;
      if (use_side) { EXC(cudaEventRecord(fork_e[1], st)); EXC(cudaStreamWaitEvent(side, fork_e[1], 0)); }
      EX(tc3_wgrad(&cm, 1, nullptr, 0, x_in, na, nullptr, Cn, r.g_xp, 128, r.dz_c, r.gxh_c, grads[2 * l + 1], r.ws_c, wst));
      // ---- transposes of the twice-message aggregation and the combination into d e_in
      const int32_t* face_tab = t->neighbour_cell;
      if (mp_times >= 2) {
        EX(fvgn_segment_sum_rows(r.dA_c, 192, Cn, 128, 0, 64, t->node_from_cell_ptr, t->node_from_cell_items, Nn, nullptr, 0, 1.0f / 3.0f, r.g_na, 64, st));
        face_tab = t->face_node;
      } else {
        EXC(cudaMemcpy2DAsync(r.g_na, 64 * sizeof(float), r.dA_c + 128, 192 * sizeof(float), 64 * sizeof(float), C, cudaMemcpyDeviceToDevice, st));
      }
      float* g_e_next = (l == lo) ? g_e_out : s.g_e[it % 3];
      EX(fvgn_edge_grad_combine(g_e, r.dA_e, r.g_na, face_tab, Fn, g_e_next, st));
      g_e = g_e_next;
      g_x = r.dA_c;
      ld_gx = 192;
      // ---- one-stage-delayed join: main waits for the weight gradients of the PREVIOUS iteration, then this iteration's are marked
      if (use_side) {
        const int prev = (it + 2) % 3;
        if (done_valid[prev]) { EXC(cudaStreamWaitEvent(st, done[prev], 0)); done_valid[prev] = false; }
        EXC(cudaEventRecord(done[it % 3], side));
        done_valid[it % 3] = true;
      }
    }
    // d / d x_in of layer lo: the first 128 columns of the last dA_c (row stride 192) -> contiguous output
    EXC(cudaMemcpy2DAsync(g_x_out, 128 * sizeof(float), g_x, 192 * sizeof(float), 128 * sizeof(float), C, cudaMemcpyDeviceToDevice, st));
    if (use_side)
      for (int k = 0; k < 3; ++k)
        if (done_valid[k]) EXC(cudaStreamWaitEvent(st, done[k], 0));
  }
cleanup:
#undef EX
#undef EXC
  return rc;
}
