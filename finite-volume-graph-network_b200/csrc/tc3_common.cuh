// tc3_common.cuh — constants, operand-split helpers, shared-memory / TMEM layouts of the fp32-equivalent tensor-core kernels
// (mlp_tc3.cu: forward; mlp_tc3_bwd.cu: data-gradient chain, weight gradients, LayerNorm backward, column sums).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

// Timing experiments (skip loads / epilogue math / MMAs to isolate pipeline stages; results are garbage): compiled in only with
// -DFVGN_TC_TIMING_EXPERIMENTS=1, then selected at run time by the FVGN_TC_DBG environment variable.  Product builds have no such path.
#ifndef FVGN_TC_TIMING_EXPERIMENTS
#define FVGN_TC_TIMING_EXPERIMENTS 0
#endif
#if FVGN_TC_TIMING_EXPERIMENTS
#define FVGN_DBG(p) ((p).dbg)
#else
#define FVGN_DBG(p) 0
#endif

namespace fvgn {
namespace x3 {

#if FVGN_TC_TIMING_EXPERIMENTS
// per-role wait-time accounting of the forward kernel (clock64 sums over all CTAs of all launches since the last read): tools/x3_prof.py
static __device__ unsigned long long g_prof[32];
#define PROF_DECL long long pt_ = 0; unsigned long long pacc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_T0() (pt_ = clock64())
#define PROF_ADD(i) do { const long long n_ = clock64(); pacc_[i] += (unsigned long long)(n_ - pt_); pt_ = n_; } while (0)
#define PROF_FLUSH(base) do { for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_prof[(base) + i_], pacc_[i_]); } while (0)
#else
#define PROF_DECL
#define PROF_T0()
#define PROF_ADD(i)
#define PROF_FLUSH(base)
#endif

constexpr int TM = 128, HIDN = 128, KTILE = 64;
constexpr int TILE_BYTES = TC_TILE_BYTES, SLOT_BYTES = 3 * TILE_BYTES;  // one K-tile step = three split sub-tiles
constexpr int ARING = 2, WRING = 2;                 // backward kernels
constexpr int FA_SLOT = 2 * TILE_BYTES, FARING = 3;  // forward: two fp16 split sub-tiles per K-tile step, 3-slot rings
constexpr int FW_SLOT = 2 * TILE_BYTES, FWRING = 3;
constexpr float ASCALE = 8.0f;                       // power-of-two pre-scale of the forward activation operand
constexpr int NEPI = 256, NPROD = 256, NPW = NPROD / 32;  // epilogue: 8 warps = two threads (column halves) per tile row
constexpr int WARP_PROD0 = NEPI / 32, WARP_MMA = WARP_PROD0 + NPW, WARP_LOAD = WARP_MMA + 1;
constexpr int NTHREADS = (WARP_LOAD + 1) * 32;  // 576
constexpr int RPT = TM / (NPROD / 16);          // 8 rows per producer thread per K-tile
constexpr uint32_t TMEM_COLS = 512;             // acc0 [0,128) | acc1 [128,256) | H split 0 [256,320) | 1 [320,384) | 2 [384,448)
constexpr uint32_t TMEM_H0 = 256;               // (forward kernel: see fwd_acc_col / fwd_h_col)
// Forward kernel TMEM map: tile jl of a CTA (parity a = jl & 1) owns the column regions 128 a and 256 + 128 a; one is its fp32
// accumulator, the other its two fp16 hidden-activation planes (hi | lo, 64 columns each), and the roles swap from one tile of a
// parity to the next: layer 1 of tile jl + 2 accumulates into the columns whose last reader was layer 3 of tile jl (the tensor pipe
// executes in order, so no barrier is needed), while the epilogue of tile jl is still reading that tile's accumulator.
__host__ __device__ constexpr uint32_t fwd_acc_col(int jl) { return ((jl >> 1) & 1) ? 256u + 128u * (uint32_t)(jl & 1) : 128u * (uint32_t)(jl & 1); }
__host__ __device__ constexpr uint32_t fwd_h_col(int jl) { return ((jl >> 1) & 1) ? 128u * (uint32_t)(jl & 1) : 256u + 128u * (uint32_t)(jl & 1); }
// kind::f16 instruction descriptor with fp16 operands (a_format bits [7,10) = 0, b_format [10,13) = 0); mixing an fp16 A with a
// bf16 B is an illegal instruction on sm_100a
constexpr uint32_t IDESC_F16_M128_N128 = IDESC_BF16_M128_N128 & ~0x480u;

enum { TC_ROWS = 0, TC_CELL = 1, TC_EDGE = 2 };

struct Params {
  int mode, rows, ntiles, kt1, k_in, n_out, dbg, nprod;
  const float* in; int ld_in;                               // ROWS
  const float* src0; const float* src1; const int32_t* idx;  // CELL: x, node_agg, cells_node_T ; EDGE: x', e, neighbour_cell
  float* out0; int ld_out;
  float* out1;
  float* zsave;                 // optional [rows, 384]: pre-activations (z1 | z2 | z3) for the backward
  // SH variant (inference): x', which only ever feeds the EdgeBlock's gathers, travels as a "split shadow" [C][hi plane 128 fp16 |
  // lo plane 128 fp16] (512 B per row, values pre-scaled by ASCALE): written by the CellBlock epilogue, moved by cp.async straight
  // into the EdgeBlock's A ring (no registers, no conversion work for 4 of its 6 K-tile steps)
  const unsigned char* sh_in0;  // EDGE: shadow of x' (gathered); CELL: split shadow [C][2][64] of the per-cell node mean (cell_mean_split_kernel)
  unsigned char* sh_prime;      // CELL: shadow of x' (out)
  const unsigned char* packed;  // [kt1 + 4] steps x [3 splits] x 16 KB
  const float* b1; const float* b2; const float* b3; const float* ln_w; const float* ln_b;
};

// exact 3-way split of two fp32 values into three packed bf16 pairs (a = s0 + s1 + s2 up to 2^-24 relative).  Packed
// cvt.rn.bf16x2.f32 conversions; a bf16 is the top half of its fp32, so widening back is a shift / mask.
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  p0 = cvt_bf16x2(a, b);
  const float ra = a - __uint_as_float(p0 << 16), rb = b - __uint_as_float(p0 & 0xffff0000u);  // exact
  p1 = cvt_bf16x2(ra, rb);
  const float sa = ra - __uint_as_float(p1 << 16), sb = rb - __uint_as_float(p1 & 0xffff0000u);  // exact
  p2 = cvt_bf16x2(sa, sb);
}

// the two leading bf16 terms only (backward kernels run with the three leading split products: the third term is never read)
__device__ __forceinline__ void split3_pair_lead2(float a, float b, uint32_t& p0, uint32_t& p1) {
  p0 = cvt_bf16x2(a, b);
  p1 = cvt_bf16x2(a - __uint_as_float(p0 << 16), b - __uint_as_float(p0 & 0xffff0000u));
}

// 2-way fp16 split of two fp32 values (forward activation operand): a = h0 + h1 up to 2^-22 relative
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t& p0, uint32_t& p1) {
  const __half2 h0 = __floats2half2_rn(a, b);
  const float2 f0 = __half22float2(h0);
  const __half2 h1 = __floats2half2_rn(a - f0.x, b - f0.y);  // the differences are exact
  p0 = *reinterpret_cast<const uint32_t*>(&h0);
  p1 = *reinterpret_cast<const uint32_t*>(&h1);
}

// SiLU for the fp32-equivalent mode: x * rcp(1 + 2^(-x log2 e)) with ex2.approx and rcp.approx (each <= 1 ulp-level, 2^-22/2^-23
// relative): <= ~4e-7 relative overall, well inside this mode's 1e-5 budget.  Both are single MUFU instructions: no slow-path
// branches (as in expf / IEEE division / __frcp_rn), so the 32 independent elements of a chunk pipeline through the SFU.
__device__ __forceinline__ float silu_x3(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

// ASCALE * silu(v) in one instruction less: ASCALE / (1 + e) = rcp((1 + e) / ASCALE), the power-of-two scale folded into the FMA that
// forms the reciprocal's argument (rcp.approx works on the mantissa: the same bits as scaling afterwards)
__device__ __forceinline__ float silu_x3_scaled(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(e, 1.0f / ASCALE, 1.0f / ASCALE)));
  return v * r;
}

// 4 consecutive fp32 (columns 4*hl .. 4*hl+3 of a 64-wide K-tile) -> row r of the three swizzled split sub-tiles of one ring slot
__device__ __forceinline__ void st_slot_f4(unsigned char* slot, int r, int hl, float4 v, bool third) {
  const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)hl >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)hl & 1u) * 8u;
  uint2 q0, q1, q2;
  if (!third) {
    split3_pair_lead2(v.x, v.y, q0.x, q1.x);
    split3_pair_lead2(v.z, v.w, q0.y, q1.y);
    *reinterpret_cast<uint2*>(slot + off) = q0;
    *reinterpret_cast<uint2*>(slot + TILE_BYTES + off) = q1;
    return;
  }
  split3_pair(v.x, v.y, q0.x, q1.x, q2.x);
  split3_pair(v.z, v.w, q0.y, q1.y, q2.y);
  *reinterpret_cast<uint2*>(slot + off) = q0;
  *reinterpret_cast<uint2*>(slot + TILE_BYTES + off) = q1;
  *reinterpret_cast<uint2*>(slot + 2 * TILE_BYTES + off) = q2;
}

// forward flavour: two fp16 split sub-tiles per slot
__device__ __forceinline__ void st_slot2_f4(unsigned char* slot, int r, int hl, float4 v) {
  const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)hl >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)hl & 1u) * 8u;
  uint2 q0, q1;
  split2_pair(v.x * ASCALE, v.y * ASCALE, q0.x, q1.x);
  split2_pair(v.z * ASCALE, v.w * ASCALE, q0.y, q1.y);
  *reinterpret_cast<uint2*>(slot + off) = q0;
  *reinterpret_cast<uint2*>(slot + TILE_BYTES + off) = q1;
}

// 4 consecutive latent values of one row -> the two planes of a split shadow row
__device__ __forceinline__ void st_shadow_f4(unsigned char* base, size_t row, int col, float4 v) {
  uint2 q0, q1;
  split2_pair(v.x * ASCALE, v.y * ASCALE, q0.x, q1.x);
  split2_pair(v.z * ASCALE, v.w * ASCALE, q0.y, q1.y);
  unsigned char* d = base + row * 512 + (size_t)col * 2;
  *reinterpret_cast<uint2*>(d) = q0;
  *reinterpret_cast<uint2*>(d + 256) = q1;
}

struct SmemLayout {
  uint32_t a_off, w_off, stage_off, par_off, red_off, bar_off, total;
};
__host__ __device__ inline SmemLayout smem_layout() {
  SmemLayout L;
  uint32_t o = 0;
  L.a_off = o; o += ARING * SLOT_BYTES;   // 96 KB
  L.w_off = o; o += WRING * SLOT_BYTES;   // 96 KB
  L.stage_off = o; o += 8 * 2048;         // 16 KB per-epilogue-warp output staging [32 rows][16 fp32]
  L.par_off = o; o += 5 * 128 * 4;
  L.red_off = o; o += 2 * 2 * 128 * 4;    // LayerNorm exchange between the two column halves of a row
  L.bar_off = o; o += 256;
  L.total = o;
  return L;
}

__host__ __device__ inline SmemLayout smem_layout_fwd() {
  SmemLayout L;
  uint32_t o = 0;
  L.a_off = o; o += FARING * FA_SLOT;     // 96 KB
  L.w_off = o; o += FWRING * FW_SLOT;     // 96 KB
  L.stage_off = o; o += 8 * 2048;
  L.par_off = o; o += 5 * 128 * 4 + 16;   // b1 b2 b3 ln_w ln_b | the three 1 / (ASCALE * weight scale) factors
  L.red_off = o; o += 2 * 2 * 2 * 128 * 4;  // LayerNorm exchange (mean, M2) between the two column halves of a row, double-buffered by tile parity
  L.bar_off = o; o += 256;
  L.total = o;
  return L;
}
enum { FB_AFULL = 0, FB_AEMPTY = FB_AFULL + FARING, FB_WFULL = FB_AEMPTY + FARING, FB_WEMPTY = FB_WFULL + FWRING, FB_DFULL = FB_WEMPTY + FWRING,
       FB_DFULL23 = FB_DFULL + 2, FB_HREADY = FB_DFULL23 + 2, FB_COUNT = FB_HREADY + 2 };
// FB_DFULL[a]: layer-1 accumulator of a tile of parity a complete (one phase per tile); FB_DFULL23[a]: its layer-2, then layer-3
// accumulator (two phases per tile).  Separate barriers: layer 1 of tile jl + 2 may complete before the epilogue has waited for layer 3
// of tile jl (no barrier orders them), and an mbarrier waiter must never fall two phases behind.
// the three forward split products, most significant first: (activation split, weight split)
__host__ __device__ constexpr int fprod_a(int q) { return q == 2 ? 1 : 0; }  // 0 0 1
__host__ __device__ constexpr int fprod_w(int q) { return q == 1 ? 1 : 0; }  // 0 1 0

enum { B_AFULL = 0, B_AEMPTY = B_AFULL + ARING, B_WFULL = B_AEMPTY + ARING, B_WEMPTY = B_WFULL + WRING, B_DFULL = B_WEMPTY + WRING,
       B_DFREE = B_DFULL + 2, B_HREADY = B_DFREE + 2, B_COUNT = B_HREADY + 1 };

// the six split products, most significant first: (A split, W split)
__host__ __device__ constexpr int prod_a(int q) { return q == 2 || q == 4 ? 1 : (q == 5 ? 2 : 0); }  // 0 0 1 0 1 2
__host__ __device__ constexpr int prod_w(int q) { return q == 1 || q == 4 ? 1 : (q == 3 ? 2 : 0); }  // 0 1 0 2 1 0

constexpr int PACK_MAX_JOBS = 48;  // MLPs per batched pack launch (forward and backward tiles)
static inline int kt_of(int k_in) { return (k_in + KTILE - 1) / KTILE; }

}  // namespace x3
}  // namespace fvgn
