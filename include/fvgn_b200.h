/* fvgn_b200.h — C-ABI of the B200-native FVGN hot path (libfvgn_b200.so).
 *
 * The reference (Litianyu141/Finite-Volume-Graph-Network) is pure Python: it has no FFI layer of its own, so
 * each entry point below replaces the *Python function* cited beside it (paths relative to
 * /root/reference/FVGN-src/main).  INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.  Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all floating point is fp32, all indices int32 (the reference's int64 tables are narrowed once by the host);
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (CUDA-graph capturable)
 *     unless documented "synchronous";
 *   - return 0 on success, a non-zero cudaError_t / negative FVGN_E* code otherwise; fvgn_last_error() gives text;
 *   - no CPU fallback: a call on a machine without an sm_100 device fails with FVGN_E_NO_DEVICE.
 *   - index tables use the reference's own layouts: [2,F] / [3,C] row-major ("edge_index", "face" of PyG Data).
 */
#ifndef FVGN_B200_H
#define FVGN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVGN_ABI_VERSION 11 /* v11: fvgn_exchange_* (domain-decomposition exchanges over peer memory); v10: fvgn_integrator_bwd_poly (training on tri / quad tables); v9: fvgn_status_read, fvgn_loss_partials / fvgn_loss_finish (data-parallel reduce-before-log); [4,C] tri/quad tables;  v7: cells_node_T == NULL selects the mp_times = 1 CellBlock in every cell-block entry point; v8: fvgn_mlp_pack_bf16x3_many, fvgn_mlp_pack_bf16x3_bwd_many, fvgn_cell_mean, fvgn_ln_bwd column-sum partials, n_products = 3 */
#define FVGN_HIDDEN 128 /* latent width; the reference ignores --hidden_size */

#define FVGN_E_BADARG (-1)
#define FVGN_E_NO_DEVICE (-2)
#define FVGN_E_WORKSPACE (-3)
#define FVGN_E_UNSUPPORTED (-4)

/* arithmetic mode of the 128-wide MLP tiles */
#define FVGN_MATH_FP32_SIMT 0 /* FFMA, strict fp32 (rel 1e-5/step vs the reference)            */
#define FVGN_MATH_BF16_TC 1   /* tcgen05 kind::f16, bf16 operands, fp32 accumulate in TMEM       */
#define FVGN_MATH_BF16X3_TC 2 /* tcgen05, exact 3-way bf16 split of every operand, 6 MMAs per product, fp32 epilogue:
                                 fp32-equivalent (rel 1e-5/step vs the reference) at ~1/10 of the FFMA time     */

#if defined(__GNUC__)
#define FVGN_API __attribute__((visibility("default")))
#else
#define FVGN_API
#endif

typedef void* fvgn_stream_t;

FVGN_API int fvgn_abi_version(void);
FVGN_API const char* fvgn_last_error(void);
/* 0 if the current device is sm_100; fills sm count. */
FVGN_API int fvgn_device_check(int* sm_count, int* cc_major, int* cc_minor);
/* Device status word of the current device (SYNCHRONOUS: waits for the device; not for use inside a stream capture).
 * FVGN_STATUS_NONFINITE_X3: some row of a FVGN_MATH_BF16X3_TC forward tile came out non-finite since the last clear — the mode's fp16
 * operand terms overflow for |activation| >= ~8190 (the reference's fp32 torch path, FVGN_model/EncoderProcesserDecoder.py:12-48,
 * would stay finite on such an input), or the input already held inf / NaN.  Re-run the step with FVGN_MATH_FP32_SIMT. */
#define FVGN_STATUS_NONFINITE_X3 1u
FVGN_API int fvgn_status_read(unsigned int* status_host, int clear);
/* Tuning knob (A/B measurements, parity tests): 1 = the split-shadow inference kernels of FVGN_MATH_BF16X3_TC (fvgn_cell_block_fwd_split /
 * fvgn_edge_block_fwd_split) run on CTA pairs with tcgen05.mma.cta_group::2 (each CTA holds half of every weight K-tile), 0 (default) = on
 * single CTAs; same results bit for bit.  Returns the previous setting.  FVGN_X3_PAIR=1 in the environment sets the initial value to 1. */
FVGN_API int fvgn_x3_pair_enable(int on);

/* One MLP of build_mlp (FVGN_model/EncoderProcesserDecoder.py:12-48): Linear(k_in,128)-SiLU-Linear(128,128)-SiLU-
 * Linear(128,n_out) [+ LayerNorm(n_out), eps 1e-5].  Pointers are the nn.Linear tensors as they sit in the
 * state_dict: weight [out,in] row-major, bias [out].  ln_w/ln_b may be NULL (decoder). */
typedef struct {
  const float* w1; const float* b1; /* [128,k_in], [128] */
  const float* w2; const float* b2; /* [128,128],  [128] */
  const float* w3; const float* b3; /* [n_out,128],[n_out] */
  const float* ln_w; const float* ln_b; /* [n_out] or NULL */
  int k_in;
  int n_out;
  const void* packed;    /* FVGN_MATH_BF16_TC only: buffer filled by fvgn_mlp_pack_bf16 for the CURRENT weights; NULL otherwise */
  const void* packed_x3; /* FVGN_MATH_BF16X3_TC only: buffer filled by fvgn_mlp_pack_bf16x3 for the CURRENT weights; NULL otherwise */
  float* z_save;         /* optional [rows, 384] fp32 = (z1 | z2 | z3): the pre-activations of the three layers (bias added, before SiLU /
                            LayerNorm).  Forward in FVGN_MATH_BF16X3_TC mode: written when non-NULL.  Backward (fvgn_*_bwd): read when
                            non-NULL instead of recomputing the forward inside the tile (a third of the backward's FLOPs). */
  const void* packed_x3_bwd; /* fvgn_mlp_dgrad_x3 only: buffer filled by fvgn_mlp_pack_bf16x3_bwd for the CURRENT weights */
  const float* da_in;    /* optional [rows, 256] = (da2 | da1) from fvgn_mlp_dgrad_x3.  Backward (fvgn_*_bwd): when non-NULL the FFMA
                            kernel skips its data-gradient GEMMs (and writes no d_in): only weight / bias / LayerNorm gradients remain. */
  int n_products;        /* FVGN_MATH_BF16X3_TC kernels (forward, dgrad, wgrad): 0 or 6 = all six split products (fp32-equivalent);
                            1 = the leading product only, i.e. plain bf16 tiles with fp32 accumulation ("bf16" training, BASELINE configs[2]);
                            3 (dgrad / wgrad only) = the three leading products a0 w0 + a0 w1 + a1 w0 of the 3-way bf16 splits: the dropped
                            terms are <= 3 * 2^-17 of a product, well inside the 1e-4 gradient tolerance, for half the MMAs */
} fvgn_mlp_t;

/* Tensor-core modes read the weights as pre-swizzled bf16 tiles.  Re-pack whenever the fp32 weights change. */
FVGN_API size_t fvgn_mlp_packed_bytes(int k_in);
FVGN_API int fvgn_mlp_pack_bf16(const fvgn_mlp_t* mlp, void* packed, size_t packed_bytes, fvgn_stream_t stream);
/* FVGN_MATH_BF16X3_TC: every weight split exactly into three bf16 terms (w = w0 + w1 + w2), three swizzled tiles per K-tile. */
FVGN_API size_t fvgn_mlp_packed_x3_bytes(int k_in);
FVGN_API int fvgn_mlp_pack_bf16x3(const fvgn_mlp_t* mlp, void* packed, size_t packed_bytes, fvgn_stream_t stream);
/* The same for an array of MLPs in ONE launch (a training step re-packs all 33 after the optimizer step): mlps[i].packed_x3 is the
 * destination of MLP i (>= fvgn_mlp_packed_x3_bytes(mlps[i].k_in) bytes, 16-byte aligned). */
FVGN_API int fvgn_mlp_pack_bf16x3_many(const fvgn_mlp_t* mlps, int n_mlps, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a1. connectivity + geometry tables  — replaces extract_mesh_state + triangles_to_faces + reorder_face
 *     (Extract_mesh/parse_tfrecord_refactor.py:419-971, 1412-1457, 1343-1375).  Integer tables are bit-exact.
 * Two phases because F (number of unique faces) is data dependent.
 * ---------------------------------------------------------------------------------------------------------- */
FVGN_API size_t fvgn_connectivity_workspace_bytes(int n_cells);
/* Phase 1 (synchronous on `stream`): sorts each cell's nodes, packs the 3C (max,min) edges, radix-sorts them,
 * and counts unique faces.  Writes cells_node[C,3] and *n_faces_host. */
FVGN_API int fvgn_connectivity_faces(const int32_t* cells /*[C,3]*/, int n_cells, int n_nodes, int32_t* cells_node /*[C,3]*/,
                            int* n_faces_host, void* workspace, size_t workspace_bytes, fvgn_stream_t stream);
/* Phase 2: fills every table for the F found by phase 1 (same workspace, untouched in between).
 * face_rule: 0 = "A" (the branch the shipped converter takes, parse_tfrecord_refactor.py:491-509),
 *            1 = "B" (the wall/inflow/outflow branch, :511-620). */
FVGN_API int fvgn_connectivity_tables(const float* mesh_pos /*[N,2]*/, const int32_t* node_type /*[N]*/, const int32_t* cells_node,
                             int n_cells, int n_nodes, int n_faces, int face_rule,
                             int32_t* face /*[2,F]*/, int32_t* cells_face /*[C,3]*/, int32_t* neighbour_cell /*[2,F]*/,
                             int32_t* face_type /*[F]*/, int32_t* cells_type /*[C]*/, float* face_length /*[F]*/,
                             float* unit_norm_v /*[C,3,2]*/, float* centroid /*[C,2]*/, float* cells_area /*[C]*/,
                             float* cell_factor /*[C,3]*/, void* workspace, size_t workspace_bytes, fvgn_stream_t stream);

/* Destination-sorted CSR of an incidence list (stable: items of one destination keep list order, which is what makes
 * the segment sums reproduce the reference's scatter_add order, GN_blocks.py:111-122).
 *   dst[n_items] in [0,n_dst)  ->  ptr[n_dst+1], items[n_items] (positions into dst). */
FVGN_API size_t fvgn_csr_workspace_bytes(int n_items);
FVGN_API int fvgn_build_csr(const int32_t* dst, int n_items, int n_dst, int32_t* ptr, int32_t* items, void* workspace,
                   size_t workspace_bytes, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a2/a3. pre/post step — replaces Data_Pool.valid_datapreprocessing / train_datapreprocessing / create_next_graph
 *        (dataset/Load_mesh.py:904-981, 983-1074, 1076-1115), dual_edge=False.
 * node_field [N, ld] rows hold (u,v,p) at column col0 (current) and, for training, col1 (next step).
 * ---------------------------------------------------------------------------------------------------------- */
/* node->cell (cell_factor-weighted 3-gather) and node->face (2-gather mean) interpolation of `width` columns. */
FVGN_API int fvgn_interp_nodes(const float* node_field, int ld, int col0, int width, const int32_t* cells_node_T /*[3,C]*/,
                      const float* cell_factor /*[C,3]*/, const int32_t* face_node /*[2,F]*/, int n_cells, int n_faces,
                      float* on_cell /*[C,width] or NULL*/, float* on_face /*[F,width] or NULL*/, fvgn_stream_t stream);
/* edge_attr[F,5] = [u_a-u_b, v_a-v_b, x_a-x_b, y_a-y_b, face_len] with (a,b)=(s,r), or (r,s) where flip_keep[f]==0
 * (training's random direction, Load_mesh.py:1046-1061; NULL = never flipped); boundary faces (type not NORMAL/OUTFLOW)
 * get their first two entries overwritten by bc_uv[f,0:2] (row stride ld_bc).  Also writes mask_interior[F] (uint8). */
FVGN_API int fvgn_assemble_edge_attr(const float* cell_uv, int ld_uv, const float* centroid /*[C,2]*/, const int32_t* neighbour_cell /*[2,F]*/,
                            const float* face_type_len /*[F,2]*/, const float* bc_uv, int ld_bc, const uint8_t* flip_keep,
                            int n_faces, float* edge_attr /*[F,5]*/, uint8_t* mask_interior /*[F] or NULL*/, fvgn_stream_t stream);
/* create_next_graph: x[C,4] <- [cell_type, Re, next_uv];  edge_attr[:,0:2] <- uv[s]-uv[r] (boundary <- bc_uv), cols 2:5 kept. */
FVGN_API int fvgn_create_next_graph(const float* next_uv /*[C,2]*/, const int32_t* neighbour_cell, const float* face_type_len,
                           const float* bc_uv, int ld_bc, int n_cells, int n_faces, float* x /*[C,4] in/out*/,
                           float* edge_attr /*[F,5] in/out*/, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a4/a13. online Normalizer (utils/normalization.py:5-86) and the feature assembly of FVGN.update_cell_attr /
 *         update_edge_attr (FVGN_model/FVGN.py:83-138).  stats = [acc_count, num_accumulations] scalars and
 *         acc_sum[w], acc_sum_squared[w] device buffers (the module's registered buffers, updated in place).
 * ---------------------------------------------------------------------------------------------------------- */
/* acc_sum += sum_rows(x), acc_sum_squared += sum_rows(x^2), acc_count += rows, num_accumulations += 1 — skipped on
 * device when num_accumulations >= max_accumulations.  one_hot>0: x is [rows,width_raw] and the accumulated features
 * are [x, one_hot(type,one_hot)] (type read from type_col[rows*ld_type]).  Deterministic two-level reduction. */
FVGN_API int fvgn_normalizer_accumulate(const float* x, int ld_x, int width_raw, const float* type_col, int ld_type, int one_hot,
                               int rows, float max_accumulations, float* acc_count, float* num_accumulations, float* acc_sum,
                               float* acc_sum_squared, float* workspace /*>= fvgn_norm_workspace_floats()*/, fvgn_stream_t stream);
FVGN_API size_t fvgn_norm_workspace_floats(int width);
/* out[rows, width] = ([x, one_hot(type)] - mean) / std   (inverse=0)   or   x*std + mean (+ addend)   (inverse=1). */
FVGN_API int fvgn_normalizer_apply(const float* x, int ld_x, int width_raw, const float* type_col, int ld_type, int one_hot, int rows,
                          const float* acc_count, const float* acc_sum, const float* acc_sum_squared, int inverse,
                          const float* addend, int ld_addend, float* out, int ld_out, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a5/a9. Encoder / Decoder MLPs over rows (EncoderProcesserDecoder.py:51-76, 127-169) — also build_mlp's forward.
 * ---------------------------------------------------------------------------------------------------------- */
FVGN_API int fvgn_mlp_rows_fwd(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, float* out, int ld_out, int math_mode,
                      fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a6. CellBlock.forward (FVGN_model/GN_blocks.py:78-145), mp_times>=2, dual_edge=False.
 * ---------------------------------------------------------------------------------------------------------- */
/* first aggregation: node_agg[n,0:64] = sum_{f: face_node[0,f]=n} e[f,0:64] + sum_{f: face_node[1,f]=n} e[f,64:128],
 * summed in the reference's scatter order through the node CSR (ptr[N+1], items[2F] = positions in cat(face[0],face[1])). */
FVGN_API int fvgn_node_aggregate_fwd(const float* e /*[F,128]*/, const int32_t* node_ptr, const int32_t* node_items, int n_nodes, int n_faces,
                            float* node_agg /*[N,64]*/, fvgn_stream_t stream);
/* second aggregation + cell MLP + (fused GnBlock residual, EncoderProcesserDecoder.py:121-122):
 *   x_prime = LN(MLP([x[c] | (na[n0]+na[n1]+na[n2])/3]))      -> x_prime [C,128] (what EdgeBlock gathers)
 *   x_out   = x + x_prime                                     -> x_out  [C,128] (may alias x; NULL to skip)
 * mp_times = 1 (GN_blocks.py:130-138: the edge halves are scatter-added straight onto the face's two cells): pass cells_node_T = NULL
 * and node_agg = the per-cell aggregate [C,64] (fvgn_node_aggregate_fwd over the CSR of cat(neighbour_cell[0], neighbour_cell[1]));
 * it is concatenated as it is, no mean.  The same convention holds for fvgn_cell_block_fwd_xbf16, fvgn_cell_mean_split,
 * fvgn_cell_block_bwd and fvgn_mlp_wgrad_x3 (mode 1). */
FVGN_API int fvgn_cell_block_fwd(const fvgn_mlp_t* mlp, const float* x /*[C,128]*/, const float* node_agg /*[N,64]*/,
                        const int32_t* cells_node_T /*[3,C]*/, int n_cells, float* x_prime, float* x_out, int math_mode,
                        fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a7/a8. EdgeBlock.forward (GN_blocks.py:16-54) + residual:  e_prime = LN(MLP([x'[s] | x'[r] | e]));  e_out = e + e_prime.
 *        e_prime / e_out may be NULL; e_out may alias e. */
FVGN_API int fvgn_edge_block_fwd(const fvgn_mlp_t* mlp, const float* x_prime /*[C,128]*/, const float* e /*[F,128]*/,
                        const int32_t* neighbour_cell /*[2,F]*/, int n_faces, float* e_prime, float* e_out, int math_mode,
                        fvgn_stream_t stream);

/* Tensor-core inference fast path of one GnBlock (EncoderProcesserDecoder.py:112-124): x' only ever feeds the EdgeBlock's bf16
 * MMA operand, so it travels between the two kernels as a bf16 shadow [C,128] (half the gather bytes, cp.async straight into the
 * operand tiles) and both latents are updated IN PLACE (x += x', e += e').  Results are bit-identical to
 * fvgn_cell_block_fwd + fvgn_edge_block_fwd in FVGN_MATH_BF16_TC mode. */
FVGN_API int fvgn_cell_block_fwd_xbf16(const fvgn_mlp_t* mlp, float* x /*[C,128] in/out*/, const float* node_agg, const int32_t* cells_node_T,
                                       int n_cells, uint16_t* x_prime_bf16 /*[C,128] out*/, int math_mode, fvgn_stream_t stream);
FVGN_API int fvgn_edge_block_fwd_xbf16(const fvgn_mlp_t* mlp, const uint16_t* x_prime_bf16 /*[C,128]*/, float* e /*[F,128] in/out*/,
                                       const int32_t* neighbour_cell, int n_faces, int math_mode, fvgn_stream_t stream);

/* Inference fast path of one GnBlock in the fp32-equivalent mode (FVGN_MATH_BF16X3_TC): both latents are updated IN PLACE
 * (x += x', e += e'), and the two operands that only exist to feed a tensor-core tile travel as "split shadows" — per row the hi plane
 * fp16(8 v), then the lo plane fp16(8 v - hi), i.e. exactly the two-term tensor-core operand, moved from HBM into the operand tiles
 * by cp.async with no conversion work in the consuming kernel:
 *   fvgn_cell_mean_split      per-cell mean of the three node aggregates (GN_blocks.py:125-129)      -> [C][2][64] fp16
 *   fvgn_cell_block_fwd_split CellBlock MLP + LayerNorm (GN_blocks.py:78-145), x += x'; x' is written -> [C][2][128] fp16 only
 *   fvgn_edge_block_fwd_split EdgeBlock (GN_blocks.py:16-54) gathering x' from its shadow, e += e'
 * Results are bit-identical to fvgn_cell_block_fwd + fvgn_edge_block_fwd in FVGN_MATH_BF16X3_TC mode. */
FVGN_API int fvgn_cell_mean_split(const float* node_agg /*[N,64]*/, const int32_t* cells_node_T /*[3,C]*/, int n_cells,
                                  void* cell_mean_split /*[C,2,64] fp16 out*/, fvgn_stream_t stream);
/* The same per-cell mean as plain fp32 [C,64] (GN_blocks.py:125-129, same arithmetic): the training path computes it once per layer and
 * hands it to fvgn_cell_block_fwd / fvgn_mlp_wgrad_x3 as a per-cell aggregate with cells_node_T = NULL (bit-identical results; the
 * weight-gradient producers then read plain rows through their cp.async ring instead of gathering three node rows). */
FVGN_API int fvgn_cell_mean(const float* node_agg /*[N,64]*/, const int32_t* cells_node_T /*[3,C]*/, int n_cells,
                            float* cell_mean /*[C,64] out*/, fvgn_stream_t stream);
FVGN_API int fvgn_cell_block_fwd_split(const fvgn_mlp_t* mlp, float* x /*[C,128] in/out*/, const void* cell_mean_split /*[C,2,64] fp16*/,
                                       int n_cells, void* x_prime_split /*[C,2,128] fp16 out*/, fvgn_stream_t stream);
FVGN_API int fvgn_edge_block_fwd_split(const fvgn_mlp_t* mlp, const void* x_prime_split /*[C,2,128] fp16*/, float* e /*[F,128] in/out*/,
                                       const int32_t* neighbour_cell, int n_faces, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a10. face-area feature + BatchNorm1d(1) (FVGN.py:77-79, 209-227 / 286-304):
 *      raw = len*dt/((A[s]+A[r])/2);  training=1: batch statistics (biased var for the output, unbiased into running_var,
 *      momentum 0.1) ; training=0: running statistics ; training=2: caller-provided batch_stats=[mean, invstd] (DDP).  bn = [weight, bias, running_mean, running_var] device scalars. */
FVGN_API int fvgn_face_area_bn_fwd(const float* face_type_len /*[F,2]*/, const float* cell_area /*[C]*/, const int32_t* neighbour_cell,
                          int n_faces, float dt, int training, const float* bn_weight, const float* bn_bias, float* running_mean,
                          float* running_var, float* face_area /*[F]*/, float* raw /*[F] or NULL*/, float* batch_stats /*[2] mean,invstd or NULL*/,
                          float* workspace /*>= fvgn_norm_workspace_floats(1)*/, fvgn_stream_t stream);

/* Data-parallel training: local (sum, sum of squares) of the raw face-area feature (float64) for the cross-rank BatchNorm
 * statistics; the caller all-reduces them and passes [mean, invstd] back through fvgn_face_area_bn_fwd(training = 2). */
FVGN_API int fvgn_face_area_sums(const float* face_type_len, const float* cell_area, const int32_t* neighbour_cell, int n_faces, float dt,
                                 float* raw /*[F]*/, double* sums /*[2]*/, float* workspace, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a11. Intergrator.forward (EncoderProcesserDecoder.py:251-331): face flux -> cell divergence.
 *      decoded [F,5] = [u_f,v_f,p_f,Dx,Dy];  delta_u[C,2] = -(A + P/rho) + D;  continuity[C] (optional, loss_cont>0). */
FVGN_API int fvgn_integrator_fwd(const float* decoded /*[F,5]*/, const float* unv /*[C,3,2]*/, const float* face_area /*[F]*/,
                        const int32_t* cells_face_T /*[3,C]*/, int n_cells, float rho, float rhs_coef, float* delta_u /*[C,2]*/,
                        float* continuity /*[C] or NULL*/, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * a12. compute_FVM_loss (utils/loss_compute.py:24-115).  mode 0 = direct_mean_loss, 1 = global_mean_sum,
 *      2 = global_mean (mean over channels).  batch vectors are int32 graph ids.  out_host-free: writes
 *      out[5] = [loss, 0 (continuity, App. D.2), loss_mom(mean), loss_face_uv(mean), loss_face_p(mean)] on device. */
FVGN_API int fvgn_loss_fwd(const float* pred_delta_u, const float* target_delta_u, const float* pred_edge /*[F,5]*/,
                  const float* target_face /*[F,3]*/, const uint8_t* mask_interior, const int32_t* cell_batch,
                  const int32_t* face_batch, int n_cells, int n_faces, int n_graphs, int mode, float w_mom, float w_uv,
                  float w_p, float* out /*[5]*/, float* workspace /*>= fvgn_loss_workspace_floats(n_graphs)*/, fvgn_stream_t stream);
FVGN_API size_t fvgn_loss_workspace_floats(int n_graphs);
/* The two halves of fvgn_loss_fwd, for data-parallel training with direct_mean_loss (utils/loss_compute.py:92-113: ONE log over the
 * batch means): fvgn_loss_partials fills workspace with the per-(segment, chunk) partials, 6 doubles each =
 * {sum sq momentum error, n_cells, sum sq interior face uv error, n_interior, sum sq face p error, n_faces}, 32 chunks per segment
 * (the counts n_cells / n_faces are read from chunk 0); the caller all-reduces the column sums over the ranks and writes them back
 * (chunk 0 = the global sums, the other chunks 0) BEFORE fvgn_loss_finish takes the log — so every rank holds the single-process
 * loss, and fvgn_loss_bwd (which reads the same workspace) the single-process per-element gradient. */
FVGN_API int fvgn_loss_partials(const float* pred_delta_u, const float* target_delta_u, const float* pred_edge, const float* target_face,
                       const uint8_t* mask_interior, const int32_t* cell_batch, const int32_t* face_batch, int n_cells, int n_faces,
                       int n_graphs, int mode, float* workspace, fvgn_stream_t stream);
FVGN_API int fvgn_loss_finish(int n_graphs, int mode, float w_mom, float w_uv, float w_p, float* out /*[5]*/, float* workspace,
                     fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * The processor's layer loops of one TRAINING step as two calls (tensor-core path, FVGN_MATH_BF16X3_TC kernels) — what the reference
 * runs as 15 x GnBlock.forward (FVGN_model/EncoderProcesserDecoder.py:112-124, 224-238) under torch autograd (train.py:194-218).
 * train.py:119-160 draws a new batch topology every step, so the step cannot be replayed from one captured CUDA graph; these entry
 * points issue the step's kernels from C (same kernels, order, streams and bits as the per-call path).
 * Layer-major arrays: X [L+1,C,128] (X[0] = encoder output, X[l+1] = cell latent after layer l), E [L+1,F,128], XP [L,C,128] (x'),
 * NA [L,C,64] (per-cell aggregate), ZC [L,C,384] / ZE [L,F,384] (pre-activations of the Cell / EdgeBlock MLPs). */
typedef struct {
  int n_cells, n_faces, n_nodes;
  const int32_t* neighbour_cell;        /* [2,F] */
  const int32_t* face_node;             /* [2,F]           (mp_times >= 2) */
  const int32_t* cells_node_T;          /* [3,C]           (mp_times >= 2) */
  const int32_t* node_ptr;              /* [N+1] stable CSR of cat(face_node[0], face_node[1]) by node (mp_times >= 2) */
  const int32_t* node_items;            /* [2F] */
  const int32_t* cell_from_face_ptr;    /* [C+1] stable CSR of neighbour_cell.reshape(-1) by cell */
  const int32_t* cell_from_face_items;  /* [2F] */
  const int32_t* node_from_cell_ptr;    /* [N+1] stable CSR of cells_node_T.reshape(-1) by node (backward, mp_times >= 2) */
  const int32_t* node_from_cell_items;  /* [3C] */
} fvgn_train_tables_t;
FVGN_API int fvgn_processor_train_fwd(const fvgn_mlp_t* cell_mlps /*[L]*/, const fvgn_mlp_t* edge_mlps /*[L]*/, int n_layers, int mp_times,
                             const fvgn_train_tables_t* tables, float* X, float* E, float* XP, float* NA, float* ZC, float* ZE,
                             float* node_scratch /*[N,64]*/, fvgn_stream_t stream);
FVGN_API size_t fvgn_processor_train_bwd_scratch_floats(int n_cells, int n_faces, int n_nodes);
/* layers hi-1 ... lo; gradients in / out are different buffers; grads[2l], grads[2l+1] = flat gradient buffers of layer l's EdgeBlock /
 * CellBlock MLP (fvgn_mlp_grad_floats); weight gradients run on side_stream (may be NULL = same stream), joined into stream on return */
FVGN_API int fvgn_processor_train_bwd(const fvgn_mlp_t* cell_mlps, const fvgn_mlp_t* edge_mlps, int lo, int hi, int mp_times, int n_products,
                             const fvgn_train_tables_t* tables, const float* X, const float* E, const float* XP, const float* NA,
                             const float* ZC, const float* ZE, const float* g_e_in, float* g_e_out, const float* g_x_in /*or NULL*/,
                             float* g_x_out, float* const* grads /*host array [2L] of device pointers*/, float* scratch,
                             size_t scratch_floats, fvgn_stream_t side_stream, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tri / quad hybrid meshes (BASELINE configs[3]).  The reference processes triangles only (its quad branch,
 * Extract_mesh/parse_tfrecord_refactor.py:1459-1491, is never reached: SURVEY.md section 0 item 5); these entry points take cells
 * with k = 3 or 4 nodes as [C,4] tables whose 4th column is -1 on triangles, follow the reference's rules with "3" read as "k"
 * (oracle/fvgn_oracle_poly.py) and reduce to the triangle entry points bit for bit on all-triangle meshes.  Inference only: the
 * per-cell mean of the node aggregates (fvgn_cell_mean_poly, fp32 [C,64] and / or the split fp16 shadow [C,2,64]) feeds the
 * CellBlock entry points through their cells_node_T == NULL convention in every math mode. */
FVGN_API size_t fvgn_connectivity_poly_workspace_bytes(int n_cells);
FVGN_API int fvgn_connectivity_poly_faces(const int32_t* cells4 /*[C,4], -1 padded*/, int n_cells, int n_nodes, int32_t* cells_node4 /*[C,4]*/,
                                 int* n_faces_host, void* workspace, size_t workspace_bytes, fvgn_stream_t stream); /* synchronous */
FVGN_API int fvgn_connectivity_poly_tables(const float* mesh_pos, const int32_t* node_type, const int32_t* cells_node4, int n_cells, int n_nodes,
                                  int n_faces, int face_rule, int32_t* face /*[2,F]*/, int32_t* cells_face4 /*[C,4]*/,
                                  int32_t* neighbour_cell /*[2,F]*/, int32_t* face_type, int32_t* cells_type, float* face_length,
                                  float* unit_norm_v /*[C,4,2]*/, float* centroid, float* cells_area, float* cell_factor4 /*[C,4]*/,
                                  void* workspace, size_t workspace_bytes, fvgn_stream_t stream);
FVGN_API int fvgn_cell_mean_poly(const float* node_agg /*[N,64]*/, const int32_t* cells_node4_T /*[4,C]*/, int n_cells,
                        float* cell_mean /*[C,64] or NULL*/, void* cell_mean_split /*[C,2,64] fp16 or NULL*/, fvgn_stream_t stream);
FVGN_API int fvgn_interp_cells_poly(const float* node_field, int ld, int col0, int width, const int32_t* cells_node4_T /*[4,C]*/,
                           const float* cell_factor4 /*[C,4]*/, int n_cells, float* on_cell /*[C,width]*/, fvgn_stream_t stream);
FVGN_API int fvgn_integrator_fwd_poly(const float* decoded /*[F,5]*/, const float* unv4 /*[C,4,2]*/, const float* face_area,
                             const int32_t* cells_face4_T /*[4,C]*/, int n_cells, float rho, float rhs_coef, float* delta_u,
                             float* continuity /*or NULL*/, fvgn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward (training) path — the reference gets these from torch autograd over train.py:194-218.  Strict fp32 (FFMA).
 * MLP gradients come back in one buffer `grads` of fvgn_mlp_grad_floats(k_in) floats laid out as
 *   dW1[128,k_in] | dW2[128,128] | dW3[128,128] (rows >= n_out unused) | db1[128] | db2[128] | db3[128] | dLN_w[128] | dLN_b[128].
 * The kernels recompute the forward inside each tile from the layer INPUTS (nothing else is saved) and are deterministic.
 * ---------------------------------------------------------------------------------------------------------- */
FVGN_API size_t fvgn_mlp_grad_floats(int k_in);
FVGN_API size_t fvgn_mlp_bwd_workspace_floats(int k_in);
FVGN_API int fvgn_mlp_rows_bwd(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, const float* d_out, int ld_dout,
                               float* d_in /*[rows,k_in] or NULL*/, int ld_din, float* grads, float* workspace, fvgn_stream_t stream);
FVGN_API int fvgn_cell_block_bwd(const fvgn_mlp_t* mlp, const float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells,
                                 const float* d_x_prime /*[C,128]*/, const float* d_x_base /*[C,>=128] or NULL: residual-path gradient added onto
                                 d_in[:,0:128]*/, int ld_base, float* d_in /*[C,192] = d[x | cell_agg]*/, float* grads, float* workspace,
                                 fvgn_stream_t stream);
FVGN_API int fvgn_edge_block_bwd(const fvgn_mlp_t* mlp, const float* x_prime, const float* e, const int32_t* neighbour_cell, int n_faces,
                                 const float* d_e_prime /*[F,128]*/, float* d_in /*[F,384] = d[x'_s | x'_r | e]*/, float* grads, float* workspace,
                                 fvgn_stream_t stream);
/* Tensor-core data-gradient chain of one MLP (training with the FVGN_MATH_BF16X3_TC forward; needs mlp->z_save and mlp->packed_x3_bwd):
 *   fvgn_ln_bwd:        dz[rows,128] = LayerNorm backward of d_out (statistics from the saved z3); MLPs without LayerNorm pass d_out itself
 *   fvgn_mlp_dgrad_x3:  da[rows,256] = (da2 | da1);  dA[rows,k_in] (optional) = da1 . W1 (+ dA_base on the first 128 columns)
 * followed by fvgn_*_bwd with mlp->da_in = da for the weight gradients. */
FVGN_API size_t fvgn_mlp_packed_x3_bwd_bytes(int k_in);
FVGN_API int fvgn_mlp_pack_bf16x3_bwd(const fvgn_mlp_t* mlp, void* packed, size_t packed_bytes, fvgn_stream_t stream);
/* the same for an array of MLPs in one launch: mlps[i].packed_x3_bwd is the destination of MLP i */
FVGN_API int fvgn_mlp_pack_bf16x3_bwd_many(const fvgn_mlp_t* mlps, int n_mlps, fvgn_stream_t stream);
/*                       `ln_sums` (optional, fvgn_ln_bwd_sums_floats(rows) floats): per-row-range partial COLUMN SUMS of dz, d_out * xhat and
 *                       d_out — what db3, dLN_w and dLN_b are — formed in the same pass (d_out * xhat is never written, nothing is re-read) */
FVGN_API size_t fvgn_ln_bwd_sums_floats(int rows);
FVGN_API int fvgn_ln_bwd(const float* d_out, int ld_dout, const float* z_save /*[rows,384]*/, const float* ln_w, int rows, float* dz /*[rows,128]*/,
                         float* ln_sums /*or NULL*/, fvgn_stream_t stream);
/*   fvgn_mlp_wgrad_x3:  all parameter gradients of the MLP (same `grads` layout as fvgn_mlp_rows_bwd) from dz / ln_sums (fvgn_ln_bwd; the
 *                       parameter named gxhat below is that buffer),
 *                       mlp->da_in (fvgn_mlp_dgrad_x3), mlp->z_save and the layer inputs (mode 0: in/ld_in; 1 CellBlock: x, node_agg,
 *                       cells_node_T; 2 EdgeBlock: x', e, neighbour_cell): dW = P^T . Q as MN-major tcgen05 MMAs, deterministic. */
FVGN_API size_t fvgn_mlp_wgrad_x3_workspace_floats(int k_in);
FVGN_API int fvgn_mlp_wgrad_x3(const fvgn_mlp_t* mlp, int mode, const float* in, int ld_in, const float* src0, const float* src1,
                               const int32_t* idx, int rows, const float* d_out, int ld_dout, const float* dz, const float* gxhat,
                               float* grads, float* workspace, fvgn_stream_t stream);
FVGN_API int fvgn_mlp_dgrad_x3(const fvgn_mlp_t* mlp, const float* dz, int ld_dz, int rows, float* da /*[rows,256]*/, float* dA /*or NULL*/,
                               int ld_dA, const float* dA_base /*or NULL*/, int ld_base, fvgn_stream_t stream);

/* Transposes of the gathers as deterministic CSR segment sums:
 *   out[d, 0:width] = base[d] + scale * sum_{pos in CSR(d)} src[pos % n_src, col0 + (pos / n_src) * part_stride + 0:width] */
FVGN_API int fvgn_segment_sum_rows(const float* src, int ld_src, int n_src, int col0, int part_stride, int width, const int32_t* ptr,
                                   const int32_t* items, int n_dst, const float* base, int ld_base, float scale, float* out, int ld_out,
                                   fvgn_stream_t stream);
/* g_e_in[F,128] = g_e_out + dA_edge[:,256:384] + gather(g_node_agg[N,64] by face_node halves)  (transposes of GN_blocks.py:91-122) */
FVGN_API int fvgn_edge_grad_combine(const float* g_e_out, const float* dA_edge /*[F,384]*/, const float* g_node /*[N,64]*/,
                                    const int32_t* face_node /*[2,F]*/, int n_faces, float* g_e_in, fvgn_stream_t stream);
/* d loss / d predicted_delta_u [C,2] and d loss / d predicted_edge_attr [F,5] (direct terms only) from fvgn_loss_fwd's workspace */
FVGN_API int fvgn_loss_bwd(const float* pred_delta_u, const float* target_delta_u, const float* pred_edge, const float* target_face,
                           const uint8_t* mask_interior, const int32_t* cell_batch, const int32_t* face_batch, int n_cells, int n_faces,
                           int n_graphs, int mode, float w_mom, float w_uv, float w_p, const float* g_loss /*scalar or NULL (=1)*/,
                           const float* fwd_workspace, float* g_delta_u, float* g_edge, float* coef_scratch /*>= 3*n_graphs*/,
                           fvgn_stream_t stream);
/* Intergrator transpose: g_decoded[F,5] += (d/du, d/dv, 0, d/dDx, d/dDy); g_face_area[F] = d/d face_area.  (face_ptr, face_items) = CSR of
 * cells_face_T.reshape(-1) by face. */
FVGN_API int fvgn_integrator_bwd(const float* decoded, const float* unv, const float* face_area, const float* g_delta_u,
                                 const int32_t* face_ptr, const int32_t* face_items, int n_cells, int n_faces, float rho, float rhs_coef,
                                 float* g_decoded /*in/out*/, float* g_face_area, fvgn_stream_t stream);
/* the same over the [4,C] (-1 padded) face table of a tri / quad mesh: unv4 [C,4,2]; (face_ptr, face_items) = CSR of cells_face4_T.reshape(-1)
 * by face with the -1 entries collected in an extra row n_faces that is not walked.  No reference counterpart (triangles only): the
 * transpose of fvgn_integrator_fwd_poly, checked against torch autograd over oracle/fvgn_oracle_poly.py. */
FVGN_API int fvgn_integrator_bwd_poly(const float* decoded, const float* unv4, const float* face_area, const float* g_delta_u,
                                      const int32_t* face_ptr, const int32_t* face_items, int n_cells, int n_faces, float rho, float rhs_coef,
                                      float* g_decoded /*in/out*/, float* g_face_area, fvgn_stream_t stream);
/* BatchNorm1d(1) affine gradients: g_weight_bias[2] from g_face_area, the saved raw feature and (mean, invstd) */
FVGN_API int fvgn_face_area_bn_bwd(const float* g_face_area, const float* raw, const float* batch_stats, int n_faces, float* g_weight_bias,
                                   float* workspace, fvgn_stream_t stream);

/* ---- single-mesh domain decomposition: the exchange steps over PEER MEMORY (no reference counterpart — the
 * reference runs one mesh on one device).  One plan per exchange kind (interface-node partial sums, ghost-cell x', ghost next_uv) and rank:
 * peers in ascending rank order; every buffer / flag pointer is a device pointer valid on THIS device (the peers' ones are peer mappings of
 * a symmetric allocation, or plain pointers when all parts live on one device).  Receive regions exist twice (sequence parity). */
#define FVGN_EXCHANGE_MAX_PEERS 8
typedef struct {
  int n_peers, row_bytes /* multiple of 4, <= 512 */, n_if /* node kind: interface nodes */, self_pos /* node kind: peers with a lower rank */;
  long long local_parity_stride;      /* bytes between the two parity copies of my receive regions */
  unsigned long long* seq;            /* device [2]: pushes / waits completed (advanced by the kernels: graph replays keep counting) */
  unsigned int* counters;             /* device [2], zero: last-CTA counters */
  unsigned int* status;               /* device word: bit 0 = a wait gave up after ~2 s */
  void* local_buf;                    /* my receive buffer */
  const int32_t* if_rows;             /* node kind: [n_if] local rows of the interface nodes */
  void* remote_buf[FVGN_EXCHANGE_MAX_PEERS];               /* peer p's receive buffer */
  long long remote_off[FVGN_EXCHANGE_MAX_PEERS];           /* byte offset of MY region in it (parity 0) */
  long long remote_parity_stride[FVGN_EXCHANGE_MAX_PEERS];
  unsigned long long* remote_flag[FVGN_EXCHANGE_MAX_PEERS];      /* my flag in peer p's flag block */
  const unsigned long long* local_flag[FVGN_EXCHANGE_MAX_PEERS]; /* peer p's flag in mine */
  const int32_t* send_rows[FVGN_EXCHANGE_MAX_PEERS];       /* [n_send] local rows sent to peer p, in the order both sides agreed on */
  int n_send[FVGN_EXCHANGE_MAX_PEERS];
  const int32_t* recv_rows[FVGN_EXCHANGE_MAX_PEERS];       /* scatter kinds: [n_recv] local destination rows; node kind: [n_if] position of
                                                              each interface node in peer p's message, -1 = not shared with p */
  int n_recv[FVGN_EXCHANGE_MAX_PEERS];
  long long local_off[FVGN_EXCHANGE_MAX_PEERS];            /* byte offset of peer p's region in my buffer (parity 0) */
} fvgn_exchange_plan_t;
/* rows src[send_rows[p][i]] -> peer p's region (the gather is fused with the peer store), then the sequence flag is raised at every peer */
FVGN_API int fvgn_exchange_push(const fvgn_exchange_plan_t* plan, const void* src, long long src_row_stride_bytes, fvgn_stream_t stream);
/* waits for every peer's flag, then dst[recv_rows[p][i]] <- row i of peer p's region */
FVGN_API int fvgn_exchange_wait_scatter(const fvgn_exchange_plan_t* plan, void* dst, long long dst_row_stride_bytes, fvgn_stream_t stream);
/* waits, then every interface node's aggregate [.,64] <- own partial + the peers' partials in ascending rank order (deterministic, the same
 * bits on every rank that holds the node): the domain-decomposed form of the scatter-add of GN_blocks.py:111-122 */
FVGN_API int fvgn_exchange_wait_node_sum(const fvgn_exchange_plan_t* plan, float* node_agg, fvgn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FVGN_B200_H */
