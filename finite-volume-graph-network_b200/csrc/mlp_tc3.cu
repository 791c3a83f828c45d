// mlp_tc3.cu — fp32-accurate tensor-core path (FVGN_MATH_BF16X3_TC): every GEMM operand is split into narrow terms whose products
// the tensor cores evaluate exactly, accumulated in fp32 in TMEM.
//   backward (dgrad / wgrad): activations-gradients and weights are both split into three bf16 terms (8 + 8 + 8 significand bits =
//     fp32's 24, full fp32 exponent range):  a.w ~= a0.w0 + a0.w1 + a1.w0 + a0.w2 + a1.w1 + a2.w0   (six tcgen05.mma, dropped <= 2^-24)
//   forward: BOTH operands in two fp16 terms (11 + 11 significand bits), pre-scaled by exact powers of two so that the low terms stay
//     normal fp16 numbers: activations (gathered rows, hidden activations) by 2^3, each weight matrix by 2^s with max|w| 2^s in
//     [2^13, 2^14) (the scale is found on the device by the pack kernel and undone, exactly, in the epilogue):
//       a.w ~= h0.g0 + h0.g1 + h1.g0                      (three tcgen05.mma; dropped / split residuals <= ~2^-22: 7e-8 relative
//                                                          on N(0,1) data, the level of an fp32 GEMM's own rounding)
//     Half the MMAs, two thirds of the operand bytes of the 3 x bf16 scheme.  Forward activations are O(1) (normalised inputs,
//     LayerNorm'd latents, SiLU outputs): |a| >= 8190 overflows the fp16 term to inf and the output row to NaN (loud, not silent).
//     Gradients have no such bound, hence the backward keeps bf16.
// Bias, SiLU (ex2.approx / rcp.approx), LayerNorm (two-pass) and the residual are fp32 in the epilogue, so results stay within the
// reference's 1e-5 per-step relative tolerance (tests/test_gpu_x3.py holds every op to 5e-6 against the fp64 oracle).
//
// 120 MMAs per 128-row EdgeBlock tile; one persistent CTA per SM,
//   warps 0-7   epilogue  : two threads per row; tcgen05.ld -> bias + SiLU -> fp16 split -> tcgen05.st (hidden activations stay in TMEM
//                           as the A operand of the next layer); last layer: bias + LayerNorm + residual -> HBM via per-warp staging
//   warps 8-15  producers : gather / stream the operand rows (fp32), split, write two swizzled fp16 K-tiles per 64 columns
//   warp 16     MMA issuer: layer 1 of tile j+1 is issued before layers 2/3 of tile j (two accumulators)
//   warp 17     loader    : ALL weights are streamed (2 x 160 KB per EdgeBlock MLP cannot stay resident): one 32 KB cp.async.bulk
//                           per K-tile step through a 3-slot ring, in the MMA issuer's consumption order
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
__device__ unsigned long long g_prof[32];
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
  L.red_off = o; o += 2 * 2 * 128 * 4;
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

template <int MODE, bool RED, bool SH>
__global__ void __launch_bounds__(NTHREADS, 1) fused_mlp_tc3_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const SmemLayout L = smem_layout_fwd();
  unsigned char* sA = smem + L.a_off;
  unsigned char* sW = smem + L.w_off;
  float* sPar = reinterpret_cast<float*>(smem + L.par_off);
  const uint32_t bar0 = smem_u32(smem + L.bar_off);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * FB_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt1 = (MODE == TC_EDGE) ? 6 : (MODE == TC_CELL) ? 3 : p.kt1;
  const int n_local = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("fvgn mlp_tc3: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < FARING; ++i) { mbar_init(BAR(FB_AFULL + i), NPW); mbar_init(BAR(FB_AEMPTY + i), 1); }
    for (int i = 0; i < FWRING; ++i) { mbar_init(BAR(FB_WFULL + i), 1); mbar_init(BAR(FB_WEMPTY + i), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(BAR(FB_DFULL + a), 1); mbar_init(BAR(FB_DFULL23 + a), 1); }
    mbar_init(BAR(FB_HREADY), 8); mbar_init(BAR(FB_HREADY + 1), 8);
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  if (warp >= WARP_PROD0 && warp < WARP_MMA) {
    for (int i = tid - NEPI; i < 5 * 128; i += NPROD) {
      const int which = i >> 7, c = i & 127;
      float v = 0.f;
      if (which == 0) v = p.b1[c];
      else if (which == 1) v = p.b2[c];
      else if (which == 2) v = (c < p.n_out) ? p.b3[c] : 0.f;
      else if (which == 3) v = p.ln_w ? p.ln_w[c] : 1.f;
      else v = p.ln_b ? p.ln_b[c] : 0.f;
      sPar[i] = v;
    }
    if (tid - NEPI < 3) sPar[5 * 128 + tid - NEPI] = reinterpret_cast<const float*>(p.packed + (size_t)(kt1 + 4) * FW_SLOT)[tid - NEPI];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= WARP_PROD0 && warp < WARP_MMA) {
    // ================================================================= producers: gather (fp32) -> split -> A ring
    constexpr int NHW = NPROD / 16;
    const int ptid = tid - NEPI;
    const int hw = ptid >> 4, hl = ptid & 15;
    const int S = n_local * kt1;
    const int last = p.rows - 1;
    PROF_DECL;
    int nidx[RPT];  // EDGE: gather indices of the next gathered K-tile pair
    if (MODE == TC_EDGE && S > 0) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) nidx[i] = __ldg(p.idx + min((int)blockIdx.x * TM + hw + NHW * i, last));
    }
    auto load = [&](int s, float4* v) {
      if (FVGN_DBG(p) & 1) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
      }
      const int jl = s / kt1, t = s - jl * kt1;
      const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
      if (MODE == TC_EDGE) {
        if (t < 4) {  // x'[sender] (t = 0, 1), x'[receiver] (t = 2, 3): the indices were fetched two steps ago
          const float* base = p.src0 + (t & 1) * 64 + hl * 4;
#pragma unroll
          for (int i = 0; i < RPT; ++i) v[i] = ldg4(base + (size_t)nidx[i] * 128);
          if (t & 1) {  // this index set is used up: fetch the receivers of this tile / the senders of the next tile
            const int r0 = (t == 3) ? row0 + (int)gridDim.x * TM : row0;
            const int32_t* ip = p.idx + (t == 3 ? 0 : (size_t)p.rows);
#pragma unroll
            for (int i = 0; i < RPT; ++i) nidx[i] = __ldg(ip + min(r0 + hw + NHW * i, last));
          }
        } else {
          const float* base = p.src1 + (t & 1) * 64 + hl * 4;
#pragma unroll
          for (int i = 0; i < RPT; ++i) v[i] = ldg4(base + (size_t)min(row0 + hw + NHW * i, last) * 128);
        }
      } else if (MODE == TC_CELL) {
        const int C = p.rows;
        if (t < 2) {
          const float* base = p.src0 + t * 64 + hl * 4;
#pragma unroll
          for (int i = 0; i < RPT; ++i) v[i] = ldg4(base + (size_t)min(row0 + hw + NHW * i, last) * 128);
        } else {
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            // the reference's arithmetic: (a + b) + c, then a true division by 3.0 (GN_blocks.py:125-129); mp_times = 1: the row itself
            v[i] = cell_agg4(p.src1, p.idx, (size_t)C, min(row0 + hw + NHW * i, last), hl * 4);
          }
        }
      } else {  // TC_ROWS
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int gr = row0 + hw + NHW * i;
          float x4[4] = {0.f, 0.f, 0.f, 0.f};
          if (gr < p.rows) {
            const int k0 = t * KTILE + hl * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k0 + j < p.k_in) x4[j] = p.in[(size_t)gr * p.ld_in + k0 + j];
          }
          v[i] = make_float4(x4[0], x4[1], x4[2], x4[3]);
        }
      }
    };
    auto store = [&](int s, const float4* v) {
      const uint32_t slot = (uint32_t)s % FARING;
      PROF_ADD(0);  // gather issue
      if (s >= FARING) mbar_wait(BAR(FB_AEMPTY + slot), (((uint32_t)s / FARING) - 1) & 1, 1);
      PROF_ADD(1);  // wait for a free A slot
      unsigned char* dst = sA + slot * FA_SLOT;
      if (!(FVGN_DBG(p) & 64)) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) st_slot2_f4(dst, hw + NHW * i, hl, v[i]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(FB_AFULL + slot));
      PROF_ADD(2);  // split + store (includes the wait for the gathered data)
    };
    PROF_T0();
    if (SH && MODE == TC_CELL) {
      // x rows (fp32 master) through registers, loaded one tile ahead; the node mean comes ready-split from cell_mean_split_kernel
      const int plane = hl >> 3, chunk = hl & 7;
      auto publish = [&](int s) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(FB_AFULL + (uint32_t)s % FARING));
      };
      auto acquire = [&](int s) {
        if (s >= FARING) mbar_wait(BAR(FB_AEMPTY + (uint32_t)s % FARING), (((uint32_t)s / FARING) - 1) & 1, 1);
        return sA + ((uint32_t)s % FARING) * FA_SLOT;
      };
      auto ldx = [&](int jl, int half, float4* v) {
        const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
        const float* base = p.src0 + half * 64 + hl * 4;
#pragma unroll
        for (int i = 0; i < RPT; ++i) v[i] = ldg4(base + (size_t)min(row0 + hw + NHW * i, last) * 128);
      };
      float4 va[RPT], vb[RPT];
      if (n_local > 0) { ldx(0, 0, va); ldx(0, 1, vb); }
#pragma unroll 1
      for (int jl = 0; jl < n_local; ++jl) {
        const int s = 3 * jl;
        const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
        PROF_ADD(0);
        unsigned char* dst = acquire(s);
        PROF_ADD(1);
#pragma unroll
        for (int i = 0; i < RPT; ++i) st_slot2_f4(dst, hw + NHW * i, hl, va[i]);
        publish(s);
        dst = acquire(s + 1);
#pragma unroll
        for (int i = 0; i < RPT; ++i) st_slot2_f4(dst, hw + NHW * i, hl, vb[i]);
        publish(s + 1);
        dst = acquire(s + 2);
        const unsigned char* src = p.sh_in0 + plane * 128 + chunk * 16;
        const uint32_t d0 = smem_u32(dst) + (uint32_t)plane * TILE_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = hw + NHW * i;
          cp_async16(d0 + (uint32_t)r * 128u + (((uint32_t)chunk ^ ((uint32_t)r & 7u)) << 4), src + (size_t)min(row0 + r, last) * 256);
        }
        cp_async_commit();
        if (jl + 1 < n_local) { ldx(jl + 1, 0, va); ldx(jl + 1, 1, vb); }
        cp_async_wait<0>();
        publish(s + 2);
        PROF_ADD(2);
      }
    } else if (SH && MODE == TC_EDGE) {
      // gathered x' rows come from the split shadow: cp.async 16-byte chunks straight into the swizzled ring slots; 16 lanes move one
      // row (2 planes x 128 B), 8 rows per thread and K-tile step.  A step is published one step later (its cp.async group done).
      const int plane = hl >> 3, chunk = hl & 7;
      auto publish = [&](int s) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(FB_AFULL + (uint32_t)s % FARING));
      };
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        const uint32_t slot = (uint32_t)s % FARING;
        const int jl = s / kt1, t = s - jl * kt1;
        const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
        PROF_ADD(0);
        if (s >= FARING) mbar_wait(BAR(FB_AEMPTY + slot), (((uint32_t)s / FARING) - 1) & 1, 1);
        PROF_ADD(1);
        unsigned char* dst = sA + slot * FA_SLOT;
        if (t >= 4) {  // this tile's own e rows (fp32 master): through registers
          float4 v[RPT];
          const float* base = p.src1 + (t & 1) * 64 + hl * 4;
#pragma unroll
          for (int i = 0; i < RPT; ++i) v[i] = ldg4(base + (size_t)min(row0 + hw + NHW * i, last) * 128);
#pragma unroll
          for (int i = 0; i < RPT; ++i) st_slot2_f4(dst, hw + NHW * i, hl, v[i]);
        } else {
          const unsigned char* src = p.sh_in0 + plane * 256 + (t & 1) * 128 + chunk * 16;
          const uint32_t d0 = smem_u32(dst) + (uint32_t)plane * TILE_BYTES;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = hw + NHW * i;
            cp_async16(d0 + (uint32_t)r * 128u + (((uint32_t)chunk ^ ((uint32_t)r & 7u)) << 4), src + (size_t)nidx[i] * 512);
          }
          if (t & 1) {  // index set used up: the receivers of this tile / the senders of the next tile
            const int r0 = (t == 3) ? row0 + (int)gridDim.x * TM : row0;
            const int32_t* ip = p.idx + (t == 3 ? 0 : (size_t)p.rows);
#pragma unroll
            for (int i = 0; i < RPT; ++i) nidx[i] = __ldg(ip + min(r0 + hw + NHW * i, last));
          }
        }
        cp_async_commit();
        if (s >= 1) { cp_async_wait<1>(); publish(s - 1); }
        PROF_ADD(2);
      }
      if (S > 0) { cp_async_wait<0>(); publish(S - 1); }
    } else {
      float4 va[RPT], vb[RPT];
#pragma unroll 1
      for (int s = -1; s < S; s += 2) {
        if (s + 1 < S) load(s + 1, va);
        if (s >= 0) store(s, vb);
        if (s + 2 < S) load(s + 2, vb);
        if (s + 1 < S) store(s + 1, va);
      }
    }
    if (ptid == 0) PROF_FLUSH(16);
  } else if (warp == WARP_LOAD) {
    // ================================================================= weight loader: 48 KB per K-tile step, MMA consumption order
    if (lane == 0 && n_local > 0) {
      uint32_t wl = 0;
      PROF_DECL;
      PROF_T0();
      auto put = [&](int step) {
        const uint32_t slot = wl % FWRING;
        if (wl >= FWRING) mbar_wait(BAR(FB_WEMPTY + slot), ((wl / FWRING) - 1) & 1, 2);
        PROF_ADD(0);
        const uint32_t nbytes = (FVGN_DBG(p) & 16) ? TILE_BYTES : FW_SLOT;  // timing experiment: half of the weight bytes
        mbar_arrive_expect_tx(BAR(FB_WFULL + slot), nbytes);
        bulk_g2s(smem_u32(sW + slot * FW_SLOT), p.packed + (size_t)step * FW_SLOT, nbytes, BAR(FB_WFULL + slot));
        ++wl;
      };
      auto put_l1 = [&]() { for (int t = 0; t < kt1; ++t) put(t); };
      put_l1();
      if (n_local > 1) put_l1();
#pragma unroll 1
      for (int k = 0; k < n_local; k += 2) {  // the MMA issuer's order: tiles in pairs (A, B)
        const bool hasB = k + 1 < n_local;
        put(kt1); put(kt1 + 1);                    // W2 for A
        if (hasB) { put(kt1); put(kt1 + 1); }      //        B
        put(kt1 + 2); put(kt1 + 3);                // W3 for A
        if (hasB) { put(kt1 + 2); put(kt1 + 3); }  //        B
        if (k + 2 < n_local) put_l1();
        if (k + 3 < n_local) put_l1();
      }
      PROF_FLUSH(24);
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ================================================================= MMA issuer
    if (lane == 0 && n_local > 0) {
      uint32_t ak = 0, wk = 0, hphase[2] = {0, 0};
      PROF_DECL;
      PROF_T0();
      auto layer1 = [&](int jl) {
        const int a = jl & 1;
        PROF_ADD(0);
        const uint32_t d = tmem_base + fwd_acc_col(jl);
#pragma unroll 1
        for (int t = 0; t < kt1; ++t, ++ak, ++wk) {
          const uint32_t sa = ak % FARING, sw = wk % FWRING;
          mbar_wait(BAR(FB_AFULL + sa), (ak / FARING) & 1, 4);
          PROF_ADD(1);
          mbar_wait(BAR(FB_WFULL + sw), (wk / FWRING) & 1, 5);
          PROF_ADD(2);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + sa * FA_SLOT), b_base = smem_u32(sW + sw * FW_SLOT);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < p.nprod && (!(FVGN_DBG(p) & 8) || q == 0)) umma_bf16(d, umma_desc(a_base + fprod_a(q) * TILE_BYTES + kk * 32), umma_desc(b_base + fprod_w(q) * TILE_BYTES + kk * 32),
                        IDESC_F16_M128_N128, (t | kk | q) != 0);
          umma_commit(BAR(FB_AEMPTY + sa));
          umma_commit(BAR(FB_WEMPTY + sw));
          PROF_ADD(3);
        }
        umma_commit(BAR(FB_DFULL + a));
      };
      auto layer23 = [&](int jl) {  // layer 2 or 3 of tile jl (whichever hidden activation the epilogue has just published)
        const int a = jl & 1;
        const uint32_t d = tmem_base + fwd_acc_col(jl), hcol = tmem_base + fwd_h_col(jl);
        {
          mbar_wait(BAR(FB_HREADY + a), hphase[a], 7); hphase[a] ^= 1;
          PROF_ADD(4);
#pragma unroll 1
          for (int t = 0; t < 2; ++t, ++wk) {
            const uint32_t sw = wk % FWRING;
            mbar_wait(BAR(FB_WFULL + sw), (wk / FWRING) & 1, 6);
            PROF_ADD(5);
            tc_fence_after();
            const uint32_t b_base = smem_u32(sW + sw * FW_SLOT);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int q = 0; q < 3; ++q)
                if (q < p.nprod && (!(FVGN_DBG(p) & 8) || q == 0)) umma_bf16_ts(d, hcol + 64u * fprod_a(q) + 8u * (uint32_t)(4 * t + kk),
                             umma_desc(b_base + fprod_w(q) * TILE_BYTES + kk * 32), IDESC_F16_M128_N128, (t | kk | q) != 0);
            umma_commit(BAR(FB_WEMPTY + sw));
            PROF_ADD(6);
          }
          umma_commit(BAR(FB_DFULL23 + a));
        }
      };
      // Tiles are processed in pairs (A, B) with their own accumulator and hidden-activation columns; the epilogue warps alternate
      // between the two (ep1 A, ep1 B, ep2 A, ep2 B, ep3 A, ep3 B), so that every layer-2 / layer-3 product and its barrier round trips
      // run under the other tile's epilogue, and layer 1 of the next pair runs under this pair's last epilogues.
      layer1(0);
      if (n_local > 1) layer1(1);
#pragma unroll 1
      for (int k = 0; k < n_local; k += 2) {
        const bool hasB = k + 1 < n_local;
        layer23(k);
        if (hasB) layer23(k + 1);
        layer23(k);
        if (hasB) layer23(k + 1);
        if (k + 2 < n_local) layer1(k + 2);
        if (k + 3 < n_local) layer1(k + 3);
      }
      PROF_FLUSH(0);
#if FVGN_TC_TIMING_EXPERIMENTS
      atomicAdd(&g_prof[7], (unsigned long long)n_local);
#endif
    }
    __syncwarp();
  } else {
    // ================================================================= epilogue warps 0-7: two threads per row (column halves),
    // 16-column sub-chunks (register budget 96: 576 threads)
    const int q = warp & 3, h = warp >> 2;  // TMEM lane quadrant = warp id % 4; column half
    const int erow = 32 * q + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    unsigned char* stage = smem + L.stage_off + warp * 2048;  // [32 rows][64 B], 16-byte chunks XOR-swizzled by ((row >> 1) & 3)
    float* sRed = reinterpret_cast<float*>(smem + L.red_off);
    const bool has_ln = (p.ln_w != nullptr);
    const bool staged = (MODE != TC_ROWS) || (p.n_out == HIDN && (p.ld_out & 3) == 0);
    const float* b3 = sPar + 2 * 128;
    const float* gw = sPar + 3 * 128;
    const float* gb = sPar + 4 * 128;
    // 16 fp32 of this thread's row -> the warp's swizzled staging tile -> 64-byte row segments of a [rows, ld] matrix at column col0
    auto store_rows16 = [&](const float* v, float* dst, int ld, int col0, int row_base) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(stage + (uint32_t)lane * 64u + (((uint32_t)k ^ (((uint32_t)lane >> 1) & 3u)) << 4)) =
            make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      __syncwarp();
      const int rr0 = lane >> 2, cp = lane & 3;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = 8 * it + rr0, grow = row_base + rr;
        if (grow < p.rows)
          *reinterpret_cast<float4*>(dst + (size_t)grow * ld + col0 + 4 * (cp ^ ((rr >> 1) & 3))) =
              *reinterpret_cast<const float4*>(stage + (uint32_t)rr * 64u + ((uint32_t)cp << 4));
      }
      __syncwarp();
    };
    PROF_DECL;
    PROF_T0();
    // hidden-layer epilogue of tile jl: accumulator -> bias + SiLU -> split -> this tile's H columns
    auto hidden = [&](int jl, int layer) {
      const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
      const int acc = jl & 1;
      const uint32_t t_acc = t_lane + fwd_acc_col(jl);
      const uint32_t t_ha = t_lane + fwd_h_col(jl);
      uint32_t r[16];
      {
        // layer 0 reads the layer-1 accumulator (phase jl >> 1 of FB_DFULL), layer 1 the layer-2 one (even phases of FB_DFULL23)
        if (layer == 0) mbar_wait(BAR(FB_DFULL + acc), (uint32_t)(jl >> 1) & 1, 8);
        else mbar_wait(BAR(FB_DFULL23 + acc), 0, 8);
        PROF_ADD(layer);  // 0: wait for the layer-1 accumulator, 1: layer 2
        tc_fence_after();
        const float* bias = sPar + layer * 128;
        const float inv = sPar[5 * 128 + layer];  // undo the power-of-two operand scales (exact)
#pragma unroll 1
        for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {  // 16-column sub-chunk index 0..7
          tmem_ld16_issue(t_acc + 16 * sc, r);
          tmem_ld16_wait(r);
          uint32_t p0[8], p1[8];
          if (FVGN_DBG(p) & 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { p0[j] = r[j]; p1[j] = r[j + 8]; }
          } else
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + 16 * sc + 4 * j4);
            const float z0 = fmaf(__uint_as_float(r[4 * j4]), inv, b.x), z1 = fmaf(__uint_as_float(r[4 * j4 + 1]), inv, b.y);
            const float z2 = fmaf(__uint_as_float(r[4 * j4 + 2]), inv, b.z), z3 = fmaf(__uint_as_float(r[4 * j4 + 3]), inv, b.w);
            r[4 * j4] = __float_as_uint(z0); r[4 * j4 + 1] = __float_as_uint(z1); r[4 * j4 + 2] = __float_as_uint(z2); r[4 * j4 + 3] = __float_as_uint(z3);
            split2_pair(silu_x3(z0) * ASCALE, silu_x3(z1) * ASCALE, p0[2 * j4], p1[2 * j4]);
            split2_pair(silu_x3(z2) * ASCALE, silu_x3(z3) * ASCALE, p0[2 * j4 + 1], p1[2 * j4 + 1]);
          }
          if (p.zsave) store_rows16(reinterpret_cast<const float*>(r), p.zsave, 384, 128 * layer + 16 * sc, row0 + 32 * q);
          tmem_st8(t_ha + 8 * sc, p0);
          tmem_st8(t_ha + 64 + 8 * sc, p1);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(FB_HREADY + acc));
        PROF_ADD(3);  // hidden-layer epilogue work
      }
    };
    // last-layer epilogue of tile jl: bias + LayerNorm + residual -> HBM
    auto final_ep = [&](int jl) {
      const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
      const int acc = jl & 1;
      const uint32_t t_acc = t_lane + fwd_acc_col(jl);
      uint32_t r[16];
      mbar_wait(BAR(FB_DFULL23 + acc), 1, 9);  // layer-3 accumulator: the odd phases
      PROF_ADD(2);  // wait for the layer-3 accumulator
      tc_fence_after();
      const float inv3 = sPar[5 * 128 + 2];
      float mean = 0.f, rstd = 1.f;
      if (has_ln && !(FVGN_DBG(p) & 32)) {  // two-pass LayerNorm statistics, as torch; the two column halves exchange partial sums
        float s = 0.f;
#pragma unroll 1
        for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
          tmem_ld16_issue(t_acc + 16 * sc, r);
          tmem_ld16_wait(r);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(b3 + 16 * sc + 4 * j4);
            s += fmaf(__uint_as_float(r[4 * j4]), inv3, b.x); s += fmaf(__uint_as_float(r[4 * j4 + 1]), inv3, b.y);
            s += fmaf(__uint_as_float(r[4 * j4 + 2]), inv3, b.z); s += fmaf(__uint_as_float(r[4 * j4 + 3]), inv3, b.w);
          }
        }
        sRed[h * 128 + erow] = s;
        named_bar_sync(1, NEPI);
        mean = (sRed[erow] + sRed[128 + erow]) * (1.0f / 128.0f);
        float qv = 0.f;
#pragma unroll 1
        for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
          tmem_ld16_issue(t_acc + 16 * sc, r);
          tmem_ld16_wait(r);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(b3 + 16 * sc + 4 * j4);
            float d;
            d = fmaf(__uint_as_float(r[4 * j4]), inv3, b.x) - mean; qv = fmaf(d, d, qv);
            d = fmaf(__uint_as_float(r[4 * j4 + 1]), inv3, b.y) - mean; qv = fmaf(d, d, qv);
            d = fmaf(__uint_as_float(r[4 * j4 + 2]), inv3, b.z) - mean; qv = fmaf(d, d, qv);
            d = fmaf(__uint_as_float(r[4 * j4 + 3]), inv3, b.w) - mean; qv = fmaf(d, d, qv);
          }
        }
        sRed[256 + h * 128 + erow] = qv;
        named_bar_sync(1, NEPI);
        rstd = 1.0f / sqrtf((sRed[256 + erow] + sRed[256 + 128 + erow]) * (1.0f / 128.0f) + 1e-5f);
      }
      const float* res = (MODE == TC_CELL) ? p.src0 : p.src1;
#pragma unroll 1
      for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
        tmem_ld16_issue(t_acc + 16 * sc, r);
        tmem_ld16_wait(r);
        float* v = reinterpret_cast<float*>(r);
        if (FVGN_DBG(p) & 32) continue;
        if (p.zsave) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(b3 + 16 * sc + 4 * j4);
            v[4 * j4] = fmaf(v[4 * j4], inv3, b.x); v[4 * j4 + 1] = fmaf(v[4 * j4 + 1], inv3, b.y);
            v[4 * j4 + 2] = fmaf(v[4 * j4 + 2], inv3, b.z); v[4 * j4 + 3] = fmaf(v[4 * j4 + 3], inv3, b.w);
          }
          store_rows16(v, p.zsave, 384, 256 + 16 * sc, row0 + 32 * q);
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float y0 = v[4 * j4], y1 = v[4 * j4 + 1], y2 = v[4 * j4 + 2], y3 = v[4 * j4 + 3];
          if (!p.zsave) {  // (with z_save the bias and scale were applied above)
            const float4 b = *reinterpret_cast<const float4*>(b3 + 16 * sc + 4 * j4);
            y0 = fmaf(y0, inv3, b.x); y1 = fmaf(y1, inv3, b.y); y2 = fmaf(y2, inv3, b.z); y3 = fmaf(y3, inv3, b.w);
          }
          if (has_ln) {
            const float4 w = *reinterpret_cast<const float4*>(gw + 16 * sc + 4 * j4);
            const float4 o = *reinterpret_cast<const float4*>(gb + 16 * sc + 4 * j4);
            y0 = (y0 - mean) * rstd * w.x + o.x; y1 = (y1 - mean) * rstd * w.y + o.y;
            y2 = (y2 - mean) * rstd * w.z + o.z; y3 = (y3 - mean) * rstd * w.w + o.w;
          }
          v[4 * j4] = y0; v[4 * j4 + 1] = y1; v[4 * j4 + 2] = y2; v[4 * j4 + 3] = y3;
        }
        if (staged) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(stage + (uint32_t)lane * 64u + (((uint32_t)k ^ (((uint32_t)lane >> 1) & 3u)) << 4)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          __syncwarp();
          // coalesced read-out: 4 lanes per row (64 contiguous bytes), 8 rows per warp access
          const int rr0 = lane >> 2, cp = lane & 3;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = 8 * it + rr0, grow = row0 + 32 * q + rr;
            const float4 y = *reinterpret_cast<const float4*>(stage + (uint32_t)rr * 64u + ((uint32_t)cp << 4));
            const int col = 16 * sc + 4 * (cp ^ ((rr >> 1) & 3));
            if (grow < p.rows) {
              if (MODE == TC_ROWS) {
                *reinterpret_cast<float4*>(p.out0 + (size_t)grow * p.ld_out + col) = y;
              } else {
                const size_t off = (size_t)grow * 128 + col;
                if (p.out0) *reinterpret_cast<float4*>(p.out0 + off) = y;
                if (SH && MODE == TC_CELL) st_shadow_f4(p.sh_prime, (size_t)grow, col, y);
                if (RED) red_add_f4(p.out1 + off, y);
                else if (p.out1) {
                  const float4 rv = *reinterpret_cast<const float4*>(res + off);
                  *reinterpret_cast<float4*>(p.out1 + off) = make_float4(rv.x + y.x, rv.y + y.y, rv.z + y.z, rv.w + y.w);
                }
              }
            }
          }
          __syncwarp();
        } else {
          const int grow = row0 + erow;
          if (grow < p.rows) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = 16 * sc + j;
              if (col < p.n_out) p.out0[(size_t)grow * p.ld_out + col] = v[j];
            }
          }
        }
      }
      PROF_ADD(4);  // final epilogue work
    };
#pragma unroll 1
    for (int k = 0; k < n_local; k += 2) {
      const int nt = (k + 1 < n_local) ? 2 : 1;
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer)
#pragma unroll 1
        for (int b = 0; b < nt; ++b) hidden(k + b, layer);
#pragma unroll 1
      for (int b = 0; b < nt; ++b) final_ep(k + b);
      // the next pair's first hidden epilogue overwrites (as H planes) the accumulator columns this pair's last epilogue has been
      // reading, across the two column-half warps of a lane quadrant: all epilogue warps finish the pair first
      named_bar_sync(1, NEPI);
    }
    if (tid == 0) PROF_FLUSH(8);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================================
// Backward data-gradient chain of one MLP on the tensor cores (training with the bf16x3 forward):
//   dh2 = dz . W3            da2 = dh2 * silu'(z2)
//   dh1 = da2 . W2           da1 = dh1 * silu'(z1)
//   dA[:, seg] = da1 . W1[:, seg]   (+ residual-path gradient on the first 128 columns)
// dz [rows, n_out] is the gradient behind the (optional) LayerNorm (ln_bwd_kernel below), z1/z2 the saved pre-activations
// (fvgn_mlp_t.z_save).  Same pipeline as the forward kernel with the transposed weights packed in consumption order
// (W3^T: 2 K-tile steps, W2^T: 2, W1[:, seg]^T: 2 per 128-column segment); da2 | da1 are also written to HBM [rows, 256] for the
// weight-gradient kernel (mlp_bwd.cu), which then has no GEMM left but the TN products.
struct DgParams {
  int rows, ntiles, nseg, k_in, n_out, want_dA, nprod;
  const float* dz; int ld_dz;
  const float* zsave;
  float* da;
  float* dA; int ld_dA;
  const float* dA_base; int ld_base;
  const unsigned char* packed;
};

__device__ __forceinline__ float dsilu_x3(float z) {
  float e, s;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
  return s * (1.0f + z * (1.0f - s));
}

template <bool THIRD>  // THIRD: the operands carry all three split terms (six products); false: the two leading terms (<= 3 products)
__global__ void __launch_bounds__(NTHREADS, 1) dgrad_tc3_kernel(const DgParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const SmemLayout L = smem_layout();
  unsigned char* sA = smem + L.a_off;
  unsigned char* sW = smem + L.w_off;
  const uint32_t bar0 = smem_u32(smem + L.bar_off);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * B_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nseg3 = p.want_dA ? p.nseg : 0;
  constexpr bool third = THIRD;  // products 4-6 read the third split term of an operand: with <= 3 products it is neither made nor moved

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("fvgn dgrad_tc3: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < ARING; ++i) { mbar_init(BAR(B_AFULL + i), NPW); mbar_init(BAR(B_AEMPTY + i), 1); }
    for (int i = 0; i < WRING; ++i) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(BAR(B_DFULL + a), 1); mbar_init(BAR(B_DFREE + a), 8); }
    mbar_init(BAR(B_HREADY), 8);
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= WARP_PROD0 && warp < WARP_MMA) {
    // ================================================================= producers: dz rows -> split -> A ring (2 K-tile steps per tile)
    constexpr int NHW = NPROD / 16;
    const int ptid = tid - NEPI;
    const int hw = ptid >> 4, hl = ptid & 15;
    const int S = n_local * 2;
    const bool vec = ((p.ld_dz & 3) == 0) && p.n_out == 128 && ((reinterpret_cast<uintptr_t>(p.dz) & 15) == 0);
    auto load = [&](int s, float4* v) {
      const int jl = s >> 1, t = s & 1;
      const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int gr = min(row0 + hw + NHW * i, p.rows - 1);
        const int k0 = t * KTILE + hl * 4;
        if (vec) v[i] = ldg4(p.dz + (size_t)gr * p.ld_dz + k0);
        else {
          float x4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k0 + j < p.n_out) x4[j] = p.dz[(size_t)gr * p.ld_dz + k0 + j];
          v[i] = make_float4(x4[0], x4[1], x4[2], x4[3]);
        }
      }
    };
    auto store = [&](int s, const float4* v) {
      const uint32_t slot = (uint32_t)s % ARING;
      if (s >= ARING) mbar_wait(BAR(B_AEMPTY + slot), (((uint32_t)s / ARING) - 1) & 1, 1);
      unsigned char* dst = sA + slot * SLOT_BYTES;
#pragma unroll
      for (int i = 0; i < RPT; ++i) st_slot_f4(dst, hw + NHW * i, hl, v[i], third);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_AFULL + slot));
    };
    float4 va[RPT], vb[RPT];
#pragma unroll 1
    for (int s = -1; s < S; s += 2) {
      if (s + 1 < S) load(s + 1, va);
      if (s >= 0) store(s, vb);
      if (s + 2 < S) load(s + 2, vb);
      if (s + 1 < S) store(s + 1, va);
    }
  } else if (warp == WARP_LOAD) {
    if (lane == 0 && n_local > 0) {
      uint32_t wl = 0;
      const int nsteps = 4 + 2 * nseg3;
#pragma unroll 1
      for (int jl = 0; jl < n_local; ++jl)
#pragma unroll 1
        for (int step = 0; step < nsteps; ++step, ++wl) {
          const uint32_t slot = wl % WRING;
          if (wl >= WRING) mbar_wait(BAR(B_WEMPTY + slot), ((wl / WRING) - 1) & 1, 2);
          const uint32_t nbytes = third ? SLOT_BYTES : 2 * TILE_BYTES;  // (the sub-tiles of a step are contiguous, most significant first)
          mbar_arrive_expect_tx(BAR(B_WFULL + slot), nbytes);
          bulk_g2s(smem_u32(sW + slot * SLOT_BYTES), p.packed + (size_t)step * SLOT_BYTES, nbytes, BAR(B_WFULL + slot));
        }
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ================================================================= MMA issuer: batches S1 (acc0), S2 (acc1), S3[seg] (acc seg & 1)
    if (lane == 0 && n_local > 0) {
      uint32_t ak = 0, wk = 0, hphase = 0, use0 = 0, use1 = 0;
      auto acquire = [&](int a) {  // the previous batch on accumulator a has been read out
        uint32_t& u = a ? use1 : use0;
        if (u > 0) mbar_wait(BAR(B_DFREE + a), (u - 1) & 1, 3);
        ++u;
      };
      auto w_steps = [&](uint32_t d, bool a_from_ring) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t, ++wk) {
          const uint32_t sw = wk % WRING;
          uint32_t sa = 0;
          if (a_from_ring) { sa = ak % ARING; mbar_wait(BAR(B_AFULL + sa), (ak / ARING) & 1, 4); }
          mbar_wait(BAR(B_WFULL + sw), (wk / WRING) & 1, 5);
          tc_fence_after();
          const uint32_t b_base = smem_u32(sW + sw * SLOT_BYTES);
          if (a_from_ring) {
            const uint32_t a_base = smem_u32(sA + sa * SLOT_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int q = 0; q < (THIRD ? 6 : 3); ++q)
                if (q < p.nprod)
                  umma_bf16(d, umma_desc(a_base + prod_a(q) * TILE_BYTES + kk * 32), umma_desc(b_base + prod_w(q) * TILE_BYTES + kk * 32),
                            IDESC_BF16_M128_N128, (t | kk | q) != 0);
            umma_commit(BAR(B_AEMPTY + sa));
            ++ak;
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int q = 0; q < (THIRD ? 6 : 3); ++q)
                if (q < p.nprod)
                  umma_bf16_ts(d, tmem_base + TMEM_H0 + 64u * prod_a(q) + 8u * (uint32_t)(4 * t + kk),
                               umma_desc(b_base + prod_w(q) * TILE_BYTES + kk * 32), IDESC_BF16_M128_N128, (t | kk | q) != 0);
          }
          umma_commit(BAR(B_WEMPTY + sw));
        }
      };
#pragma unroll 1
      for (int jl = 0; jl < n_local; ++jl) {
        acquire(0);
        w_steps(tmem_base, true);  // S1: dz . W3
        umma_commit(BAR(B_DFULL + 0));
        mbar_wait(BAR(B_HREADY), hphase, 7); hphase ^= 1;
        acquire(1);
        w_steps(tmem_base + 128u, false);  // S2: da2 . W2
        umma_commit(BAR(B_DFULL + 1));
        mbar_wait(BAR(B_HREADY), hphase, 7); hphase ^= 1;
#pragma unroll 1
        for (int seg = 0; seg < nseg3; ++seg) {  // S3: da1 . W1[:, seg]
          acquire(seg & 1);
          w_steps(tmem_base + 128u * (uint32_t)(seg & 1), false);
          umma_commit(BAR(B_DFULL + (seg & 1)));
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================= epilogue warps 0-7: two threads per row, 16-column sub-chunks
    const int q = warp & 3, h = warp >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t t_h = t_lane + TMEM_H0;
    unsigned char* stage = smem + L.stage_off + warp * 2048;
    auto store_rows16 = [&](const float* v, float* dst, int ld, int col0, int row_base) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(stage + (uint32_t)lane * 64u + (((uint32_t)k ^ (((uint32_t)lane >> 1) & 3u)) << 4)) =
            make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      __syncwarp();
      const int rr0 = lane >> 2, cp = lane & 3;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = 8 * it + rr0, grow = row_base + rr;
        if (grow < p.rows)
          *reinterpret_cast<float4*>(dst + (size_t)grow * ld + col0 + 4 * (cp ^ ((rr >> 1) & 3))) =
              *reinterpret_cast<const float4*>(stage + (uint32_t)rr * 64u + ((uint32_t)cp << 4));
      }
      __syncwarp();
    };
    uint32_t e0 = 0, e1 = 0;  // DFULL completions consumed per accumulator
    auto wait_acc = [&](int a) {
      uint32_t& u = a ? e1 : e0;
      mbar_wait(BAR(B_DFULL + a), u & 1, 8);
      ++u;
      tc_fence_after();
    };
    auto release_acc = [&](int a) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_DFREE + a));
    };
    // The saved pre-activations z2 / z1 of a tile (their global-load latency was the epilogue's largest stall: 37 % of its samples,
    // profiles/r01c_ncu_backward_summary.txt).  Two-term variant: this thread's 256 bytes of a batch are fetched by cp.async into the
    // third sub-tile of ring slot q — 16 KB that variant neither writes nor multiplies — a whole phase before they are read: z1 of the
    // tile as soon as batch S1 has consumed z2 (the layer-2 MMAs run meanwhile), z2 of the next tile after batch S2 (three S3 batches
    // follow).  Slots are private to the thread ([16-byte chunk][64 threads]: conflict-free), so cp.async.wait_group is the only sync.
    // Six-product variant: one sub-chunk ahead through registers.
    const uint32_t z_me = smem_u32((q < 2 ? sA + q * SLOT_BYTES : sW + (q - 2) * SLOT_BYTES) + 2 * TILE_BYTES) + (uint32_t)(32 * h + lane) * 16u;
    auto issue_z = [&](int jl_, int b_) {
      if (jl_ < n_local) {
        const int row = min(((int)blockIdx.x + jl_ * (int)gridDim.x) * TM + 32 * q + lane, p.rows - 1);
        const float* zp = p.zsave + (size_t)row * 384 + (b_ == 0 ? 128 : 0) + 64 * h;
#pragma unroll
        for (int k = 0; k < 16; ++k)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(z_me + 1024u * (uint32_t)k), "l"(zp + 4 * k) : "memory");
      }
      cp_async_commit();
    };
    if (!THIRD) issue_z(0, 0);
#pragma unroll 1
    for (int jl = 0; jl < n_local; ++jl) {
      const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
      const int grow_c = min(row0 + 32 * q + lane, p.rows - 1);  // this thread's row (clamped for loads)
      uint32_t r[16];
      // ---- batches S1, S2: d(hidden) * silu'(z) -> da (HBM) + split planes (TMEM)
#pragma unroll 1
      for (int b = 0; b < 2; ++b) {
        const float* zrow = p.zsave + (size_t)grow_c * 384 + (b == 0 ? 128 : 0);  // S1 pairs with z2, S2 with z1
        float4 zn[4];
        if (THIRD) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) zn[j4] = ldg4(zrow + 16 * (4 * h) + 4 * j4);
        }
        wait_acc(b);
        if (!THIRD) cp_async_wait<0>();
#pragma unroll 1
        for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
          tmem_ld16_issue(t_lane + 128u * (uint32_t)b + 16 * sc, r);
          float4 zc[4];
          if (THIRD) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) zc[j4] = zn[j4];
            if (sc + 1 < 4 * h + 4) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) zn[j4] = ldg4(zrow + 16 * (sc + 1) + 4 * j4);
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint32_t a = z_me + 1024u * (uint32_t)(4 * (sc - 4 * h) + j4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(zc[j4].x), "=f"(zc[j4].y), "=f"(zc[j4].z), "=f"(zc[j4].w) : "r"(a));
            }
          }
          tmem_ld16_wait(r);
          if (sc == 4 * h + 3) release_acc(b);
          float v[16];
          uint32_t p0[8], p1[8], p2[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 z = zc[j4];
            v[4 * j4] = __uint_as_float(r[4 * j4]) * dsilu_x3(z.x); v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) * dsilu_x3(z.y);
            v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) * dsilu_x3(z.z); v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) * dsilu_x3(z.w);
            if (third) {
              split3_pair(v[4 * j4], v[4 * j4 + 1], p0[2 * j4], p1[2 * j4], p2[2 * j4]);
              split3_pair(v[4 * j4 + 2], v[4 * j4 + 3], p0[2 * j4 + 1], p1[2 * j4 + 1], p2[2 * j4 + 1]);
            } else {
              split3_pair_lead2(v[4 * j4], v[4 * j4 + 1], p0[2 * j4], p1[2 * j4]);
              split3_pair_lead2(v[4 * j4 + 2], v[4 * j4 + 3], p0[2 * j4 + 1], p1[2 * j4 + 1]);
            }
          }
          tmem_st8(t_h + 8 * sc, p0);
          tmem_st8(t_h + 64 + 8 * sc, p1);
          if (third) tmem_st8(t_h + 128 + 8 * sc, p2);
          store_rows16(v, p.da, 256, 128 * b + 16 * sc, row0 + 32 * q);
        }
        if (!THIRD) {  // this batch's z has been consumed: request the next batch's (z1 of this tile, or z2 of the next tile)
          if (b == 0) issue_z(jl, 1);
          else issue_z(jl + 1, 0);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HREADY));
      }
      // ---- batches S3[seg]: dA[:, 128 seg ..] (+ the residual-path gradient on segment 0)
#pragma unroll 1
      for (int seg = 0; seg < nseg3; ++seg) {
        const int a = seg & 1;
        wait_acc(a);
        const int kv = min(128, p.k_in - 128 * seg);
#pragma unroll 1
        for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
          tmem_ld16_issue(t_lane + 128u * (uint32_t)a + 16 * sc, r);
          tmem_ld16_wait(r);
          if (sc == 4 * h + 3) release_acc(a);
          if (16 * sc >= kv) continue;  // columns past the segment's valid width (k_in % 16 == 0 is checked on the host)
          float* v = reinterpret_cast<float*>(r);
          if (seg == 0 && p.dA_base != nullptr) {
            const float* brow = p.dA_base + (size_t)grow_c * p.ld_base + 16 * sc;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bb = ldg4(brow + 4 * j4);
              v[4 * j4] += bb.x; v[4 * j4 + 1] += bb.y; v[4 * j4 + 2] += bb.z; v[4 * j4 + 3] += bb.w;
            }
          }
          store_rows16(v, p.dA, p.ld_dA, 128 * seg + 16 * sc, row0 + 32 * q);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// LayerNorm backward of a row block: dz = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)), xhat from the saved z3.
// One warp per row (4 columns per lane); the affine-parameter gradients are formed by the weight-gradient kernel.
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ d_out, int ld_dout, const float* __restrict__ zsave,
                                                     const float* __restrict__ gamma, int rows, float* __restrict__ dz,
                                                     float* __restrict__ gxhat /*[rows,128] = d_out * xhat (for dLN_w) or NULL*/) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4 z = ldg4(zsave + (size_t)row * 384 + 256 + lane * 4);
  const float4 g = ldg4(d_out + (size_t)row * ld_dout + lane * 4);
  const float4 w = ldg4(gamma + lane * 4);
  float s = (z.x + z.y) + (z.z + z.w);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 128.0f);
  const float d0 = z.x - mean, d1 = z.y - mean, d2 = z.z - mean, d3 = z.w - mean;
  float qv = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
  const float rstd = 1.0f / sqrtf(qv * (1.0f / 128.0f) + 1e-5f);
  const float x0 = d0 * rstd, x1 = d1 * rstd, x2 = d2 * rstd, x3v = d3 * rstd;
  const float h0 = g.x * w.x, h1 = g.y * w.y, h2 = g.z * w.z, h3 = g.w * w.w;
  float m1 = (h0 + h1) + (h2 + h3), m2 = fmaf(h0, x0, fmaf(h1, x1, fmaf(h2, x2, h3 * x3v)));
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
  m1 *= (1.0f / 128.0f); m2 *= (1.0f / 128.0f);
  if (gxhat) *reinterpret_cast<float4*>(gxhat + (size_t)row * 128 + lane * 4) = make_float4(g.x * x0, g.y * x1, g.z * x2, g.w * x3v);
  *reinterpret_cast<float4*>(dz + (size_t)row * 128 + lane * 4) =
      make_float4(rstd * (h0 - m1 - x0 * m2), rstd * (h1 - m1 - x1 * m2), rstd * (h2 - m1 - x2 * m2), rstd * (h3 - m1 - x3v * m2));
}

// The same with the three column sums the parameter gradients need folded in (db3 = colsum(dz), dLN_w = colsum(d_out * xhat),
// dLN_b = colsum(d_out)): d_out * xhat is never written, and the column-sum kernel does not re-read the three matrices.  Block b owns the
// row range [b per, (b + 1) per); warp w walks rows w, w + 8, ... of it keeping 12 running sums per lane; the eight warps are combined
// through shared memory in a fixed order and the block writes sums[src][b][128] — the layout colsum_reduce_kernel finishes.  Fixed
// row -> (block, warp) mapping and fixed combination orders: deterministic.
constexpr int LN_BWD_NB = 592;  // at most this many row ranges, at least 64 rows each
__host__ __device__ inline int ln_bwd_blocks(int rows) { const int n = (rows + 63) / 64; return n < LN_BWD_NB ? (n > 0 ? n : 1) : LN_BWD_NB; }
__global__ void __launch_bounds__(256) ln_bwd_sums_kernel(const float* __restrict__ d_out, int ld_dout, const float* __restrict__ zsave,
                                                          const float* __restrict__ gamma, int rows, float* __restrict__ dz,
                                                          float* __restrict__ sums /*[3][gridDim.x][128]*/) {
  __shared__ float sh[3][8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (rows + (int)gridDim.x - 1) / (int)gridDim.x, r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  const float4 w = ldg4(gamma + lane * 4);
  float4 s_dz = make_float4(0.f, 0.f, 0.f, 0.f), s_gx = s_dz, s_g = s_dz;
  int row = r0 + warp;
  float4 z, g;
  if (row < r1) { z = ldg4(zsave + (size_t)row * 384 + 256 + lane * 4); g = ldg4(d_out + (size_t)row * ld_dout + lane * 4); }
#pragma unroll 1
  for (; row < r1; row += 8) {
    const float4 zc = z, gc = g;
    if (row + 8 < r1) {  // the next row of this warp is in flight while this one goes through its three reductions
      z = ldg4(zsave + (size_t)(row + 8) * 384 + 256 + lane * 4);
      g = ldg4(d_out + (size_t)(row + 8) * ld_dout + lane * 4);
    }
    float s = (zc.x + zc.y) + (zc.z + zc.w);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 128.0f);
    const float d0 = zc.x - mean, d1 = zc.y - mean, d2 = zc.z - mean, d3 = zc.w - mean;
    float qv = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
    const float rstd = 1.0f / sqrtf(qv * (1.0f / 128.0f) + 1e-5f);
    const float x0 = d0 * rstd, x1 = d1 * rstd, x2 = d2 * rstd, x3v = d3 * rstd;
    const float h0 = gc.x * w.x, h1 = gc.y * w.y, h2 = gc.z * w.z, h3 = gc.w * w.w;
    float m1 = (h0 + h1) + (h2 + h3), m2 = fmaf(h0, x0, fmaf(h1, x1, fmaf(h2, x2, h3 * x3v)));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    m1 *= (1.0f / 128.0f); m2 *= (1.0f / 128.0f);
    const float4 dzv = make_float4(rstd * (h0 - m1 - x0 * m2), rstd * (h1 - m1 - x1 * m2), rstd * (h2 - m1 - x2 * m2), rstd * (h3 - m1 - x3v * m2));
    *reinterpret_cast<float4*>(dz + (size_t)row * 128 + lane * 4) = dzv;
    s_dz.x += dzv.x; s_dz.y += dzv.y; s_dz.z += dzv.z; s_dz.w += dzv.w;
    s_gx.x += gc.x * x0; s_gx.y += gc.y * x1; s_gx.z += gc.z * x2; s_gx.w += gc.w * x3v;
    s_g.x += gc.x; s_g.y += gc.y; s_g.z += gc.z; s_g.w += gc.w;
  }
  *reinterpret_cast<float4*>(&sh[0][warp][lane * 4]) = s_dz;
  *reinterpret_cast<float4*>(&sh[1][warp][lane * 4]) = s_gx;
  *reinterpret_cast<float4*>(&sh[2][warp][lane * 4]) = s_g;
  __syncthreads();
  for (int t = threadIdx.x; t < 384; t += 256) {
    const int src = t >> 7, c = t & 127;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += sh[src][k][c];
    sums[((size_t)src * gridDim.x + blockIdx.x) * 128 + c] = a;
  }
}

// transposed packing for the backward: element (n, k) of the packed "weight" is W[k * ldw + n0 + n]  (W^T restricted to columns n0..)
// (blockIdx.y = job: all transposed blocks of one MLP in one launch)
struct PackTJob { const float* W; int ldw, n0, n_valid, k_valid, step0; };
struct PackTParams { PackTJob job[5]; int kt; unsigned char* out; };
__global__ void pack_tiles3_t_kernel(const PackTParams pp) {
  const PackTJob jb = pp.job[blockIdx.y];
  const float* __restrict__ W = jb.W;
  const int ldw = jb.ldw, n0 = jb.n0, n_valid = jb.n_valid, k_valid = jb.k_valid, kt = pp.kt;
  unsigned char* __restrict__ out = pp.out + (size_t)jb.step0 * SLOT_BYTES;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tile, n, chunk)
  if (idx >= kt * 128 * 8) return;
  const int c = idx & 7, n = (idx >> 3) & 127, t = idx >> 10;
  uint4 q0, q1, q2;
  uint32_t* a0 = reinterpret_cast<uint32_t*>(&q0);
  uint32_t* a1 = reinterpret_cast<uint32_t*>(&q1);
  uint32_t* a2 = reinterpret_cast<uint32_t*>(&q2);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = t * KTILE + c * 8 + 2 * j;
    const float x = (n < n_valid && k < k_valid) ? W[(size_t)k * ldw + n0 + n] : 0.f;
    const float y = (n < n_valid && k + 1 < k_valid) ? W[(size_t)(k + 1) * ldw + n0 + n] : 0.f;
    split3_pair(x, y, a0[j], a1[j], a2[j]);
  }
  const size_t off = (size_t)t * SLOT_BYTES + (size_t)n * 128 + (((uint32_t)c ^ ((uint32_t)n & 7u)) << 4);
  *reinterpret_cast<uint4*>(out + off) = q0;
  *reinterpret_cast<uint4*>(out + TILE_BYTES + off) = q1;
  *reinterpret_cast<uint4*>(out + 2 * TILE_BYTES + off) = q2;
}

// The same for an array of MLPs in one launch (blockIdx.z = MLP; the jobs of an MLP are derived from its dimensions): a training step
// re-packs every MLP's transposed tiles after the optimizer step.
constexpr int PACK_MAX_JOBS = 48;  // MLPs per batched pack launch (forward and backward tiles)
struct PackTMlp { const float* w1; const float* w2; const float* w3; unsigned char* out; int k_in, n_out; };
struct PackTBatch { PackTMlp m[PACK_MAX_JOBS]; };
__global__ void pack_tiles3_t_many_kernel(const __grid_constant__ PackTBatch b) {
  const PackTMlp& mm = b.m[blockIdx.z];
  const int nseg = (mm.k_in + 127) / 128, j = blockIdx.y;
  if (j >= 2 + nseg) return;
  PackTJob jb;
  if (j == 0) jb = {mm.w3, 128, 0, 128, mm.n_out, 0};
  else if (j == 1) jb = {mm.w2, 128, 0, 128, 128, 2};
  else { const int seg = j - 2; jb = {mm.w1, mm.k_in, 128 * seg, min(128, mm.k_in - 128 * seg), 128, 4 + 2 * seg}; }
  const float* __restrict__ W = jb.W;
  const int ldw = jb.ldw, n0 = jb.n0, n_valid = jb.n_valid, k_valid = jb.k_valid, kt = 2;
  unsigned char* __restrict__ out = mm.out + (size_t)jb.step0 * SLOT_BYTES;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tile, n, chunk)
  if (idx >= kt * 128 * 8) return;
  const int c = idx & 7, n = (idx >> 3) & 127, t = idx >> 10;
  uint4 q0, q1, q2;
  uint32_t* a0 = reinterpret_cast<uint32_t*>(&q0);
  uint32_t* a1 = reinterpret_cast<uint32_t*>(&q1);
  uint32_t* a2 = reinterpret_cast<uint32_t*>(&q2);
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int k = t * KTILE + c * 8 + 2 * jj;
    const float x = (n < n_valid && k < k_valid) ? W[(size_t)k * ldw + n0 + n] : 0.f;
    const float y = (n < n_valid && k + 1 < k_valid) ? W[(size_t)(k + 1) * ldw + n0 + n] : 0.f;
    split3_pair(x, y, a0[jj], a1[jj], a2[jj]);
  }
  const size_t off = (size_t)t * SLOT_BYTES + (size_t)n * 128 + (((uint32_t)c ^ ((uint32_t)n & 7u)) << 4);
  *reinterpret_cast<uint4*>(out + off) = q0;
  *reinterpret_cast<uint4*>(out + TILE_BYTES + off) = q1;
  *reinterpret_cast<uint4*>(out + 2 * TILE_BYTES + off) = q2;
}

// =====================================================================================================================
// Weight gradients of one MLP on the tensor cores (training with the bf16x3 forward):
//   dW3 = dz^T . h2      dW2 = da2^T . h1      dW1[:, 128 s ..] = da1^T . A[:, 128 s ..]        (sums over the rows)
// Both operands of P^T . Q are stored the way they come, [row][feature] (one 128-byte row per 64 features), and read by the MMA as
// MN-MAJOR operands (K = rows), 3-way split as everywhere in this mode.  A "job" is one 128 x 128 output block (dW3, dW2, one
// 128-column segment of dW1); CTA b works on job b % njobs over row range b / njobs, accumulating in ONE TMEM accumulator over all
// its 64-row chunks — no read-modify-write of partial gradients in memory (the FFMA kernel's cost) — and writes its partial block
// once; wgrad_reduce_kernel sums the ranges in a fixed order (deterministic).
constexpr int WG_CHUNK = 64;                 // rows per K chunk
constexpr int WG_HALF_BYTES = WG_CHUNK * 128;  // one [64 rows][64 features] bf16 tile
constexpr int WG_MAT_BYTES = 6 * WG_HALF_BYTES;  // 3 splits x 2 feature halves = 48 KB
constexpr int WG_STAGE_BYTES = 2 * WG_MAT_BYTES;  // P and Q
constexpr int WG_STAGES = 2;
constexpr int WG_NPW = 16;                       // producer warps (warps 0-3 also run the epilogue once the chunks are done)
constexpr int WG_NTHREADS = (WG_NPW + 1) * 32;   // + 1 MMA warp
// Two-term variant (THIRD = false): an operand block holds two split tiles per feature half (32 KB), a stage 64 KB; the 64 KB of shared
// memory the third terms no longer take + 32 KB hold a ring of three RAW half-chunks ([P | Q] x 32 rows x 128 fp32 = 32 KB each) that the
// producers fill by cp.async two steps ahead: the row loads need no registers while in flight, ~64 KB per SM instead of ~32-48 KB.
constexpr int WG_RAW_BYTES = WG_NPW * 32 * 64;   // one half-chunk: 512 producer threads x 4 x 16 bytes
constexpr int WG_RAW_DEPTH = 3;
template <bool THIRD> struct WgLayout {
  static constexpr int MAT = (THIRD ? 6 : 4) * WG_HALF_BYTES;
  static constexpr int STAGE = 2 * MAT;
  static constexpr int RAW_OFF = WG_STAGES * STAGE;
  static constexpr int BAR_OFF = RAW_OFF + (THIRD ? 0 : WG_RAW_DEPTH * WG_RAW_BYTES);
  static constexpr int TOTAL = BAR_OFF + 256;
};

struct WgParams {
  int mode, rows, k_in, n_out, njobs, nranges, chunks_per_range, nchunks, nprod;
  const float* dz; int ld_dz;   // gradient behind the LayerNorm [rows, n_out]
  const float* da;              // [rows, 256] = da2 | da1
  const float* zsave;           // [rows, 384]
  const float* in; int ld_in;   // ROWS: layer input
  const float* src0; const float* src1; const int32_t* idx;  // CELL / EDGE: as the forward
  float* part;                  // [njobs][nranges][128 * 128]
  float* csum;                  // [3][nranges * WG_NPW][128]: per-(range, producer warp) column sums of the P operand of jobs 0 / 1 / 2
  int csum_job0;                //   (= dz or d_out -> db3 when the MLP has no LayerNorm, da2 -> db2, da1 -> db1): the rows pass through anyway
};

// MN-major SWIZZLE_128B operand descriptor: 64 features per 128-byte row, 8-row groups 1024 B apart (SBO), the second 64-feature
// block of the operand one tile (8 KB) further (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(WG_HALF_BYTES >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t IDESC_BF16_MN_MN = IDESC_BF16_M128_N128 | (1u << 15) | (1u << 16);

// 4 consecutive fp32 (features 4*hl.. of one 64-feature half) -> row r of the three split tiles of a matrix block
__device__ __forceinline__ void st_wg_f4(unsigned char* mat, int half, int r, int hl, float4 v, bool third) {
  const uint32_t off = (uint32_t)half * WG_HALF_BYTES + (uint32_t)r * 128u + ((((uint32_t)hl >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)hl & 1u) * 8u;
  uint2 q0, q1, q2;
  if (!third) {
    split3_pair_lead2(v.x, v.y, q0.x, q1.x);
    split3_pair_lead2(v.z, v.w, q0.y, q1.y);
    *reinterpret_cast<uint2*>(mat + off) = q0;
    *reinterpret_cast<uint2*>(mat + 2 * WG_HALF_BYTES + off) = q1;
    return;
  }
  split3_pair(v.x, v.y, q0.x, q1.x, q2.x);
  split3_pair(v.z, v.w, q0.y, q1.y, q2.y);
  *reinterpret_cast<uint2*>(mat + off) = q0;
  *reinterpret_cast<uint2*>(mat + 2 * WG_HALF_BYTES + off) = q1;
  *reinterpret_cast<uint2*>(mat + 4 * WG_HALF_BYTES + off) = q2;
}

template <bool THIRD>
__global__ void __launch_bounds__(WG_NTHREADS, 1) wgrad_tc3_kernel(const WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using WL = WgLayout<THIRD>;
  unsigned char* sBar = smem + WL::BAR_OFF;
  const uint32_t bar0 = smem_u32(sBar);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };  // FULL[2], EMPTY[2], DONE
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sBar + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int job = (int)blockIdx.x % p.njobs, range = (int)blockIdx.x / p.njobs;
  const int c0 = range * p.chunks_per_range;
  const int nch = max(0, min(p.chunks_per_range, p.nchunks - c0));
  constexpr bool third = THIRD;  // the third split term is only read by products 4-6

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("fvgn wgrad_tc3: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(BAR(i), WG_NPW); mbar_init(BAR(2 + i), 1); }
    mbar_init(BAR(4), 1);
    fence_mbar_init();
  }
  if (warp == WG_NPW) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WG_NPW) {
    // ================================================================= producers: P (gradient rows) and Q (activation rows) chunks,
    // half a chunk (32 rows) per step with two register buffers: the next step's loads are in flight while this one is split / stored.
    // 16 warps: the split / store work is latency-bound per warp, so more warps in flight is what makes it faster.
    const int ptid = tid;
    const int hw = ptid >> 4, hl = ptid & 15;  // row 32 half + hw of the chunk, features 4 hl.. of each 64-feature half
    const int seg = job - 2;
    const int S = 2 * nch;
    auto load = [&](int s, float4 (*pv)[2], float4 (*qv)[2]) {
      const int row0 = (c0 + (s >> 1)) * WG_CHUNK + 32 * (s & 1);
#pragma unroll
      for (int i = 0; i < 1; ++i) {
        const int g = row0 + hw;
        const bool ok = g < p.rows;
        const int gc = min(g, p.rows - 1);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int col = 64 * hh + 4 * hl;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (ok) {
            // ---- P: dz (job 0), da2 (job 1), da1 (jobs >= 2)
            if (job == 0) {
              if (p.n_out == 128 && (p.ld_dz & 3) == 0) a = ldg4(p.dz + (size_t)gc * p.ld_dz + col);
              else {
                float t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = (col + j < p.n_out) ? p.dz[(size_t)gc * p.ld_dz + col + j] : 0.f;
                a = make_float4(t[0], t[1], t[2], t[3]);
              }
            } else {
              a = ldg4(p.da + (size_t)gc * 256 + (job == 1 ? 0 : 128) + col);
            }
            // ---- Q: z2 / z1 (SiLU applied at store time: jobs 0, 1), layer-input segment (jobs >= 2)
            if (job < 2) {
              b = ldg4(p.zsave + (size_t)gc * 384 + (job == 0 ? 128 : 0) + col);
            } else if (p.mode == TC_EDGE) {
              if (seg < 2) b = ldg4(p.src0 + (size_t)__ldg(p.idx + (size_t)seg * p.rows + gc) * 128 + col);
              else b = ldg4(p.src1 + (size_t)gc * 128 + col);
            } else if (p.mode == TC_CELL) {
              if (seg == 0) b = ldg4(p.src0 + (size_t)gc * 128 + col);
              else if (hh == 0) {
                b = cell_agg4(p.src1, p.idx, (size_t)p.rows, gc, 4 * hl);
              }
            } else {
              float t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = 128 * seg + col + j;
                t[j] = (k < p.k_in) ? p.in[(size_t)gc * p.ld_in + k] : 0.f;
              }
              b = make_float4(t[0], t[1], t[2], t[3]);
            }
          }
          pv[i][hh] = a; qv[i][hh] = b;
        }
      }
    };
    const bool do_cs = p.csum != nullptr && (job == 1 || job == 2 || (job == 0 && p.csum_job0 != 0));
    float4 cs0 = make_float4(0.f, 0.f, 0.f, 0.f), cs1 = cs0;  // this thread's 8 columns, rows hw, hw + 32, ... of the range in order
    auto store = [&](int s, float4 (*pv)[2], float4 (*qv)[2]) {
      if (do_cs) {
        cs0.x += pv[0][0].x; cs0.y += pv[0][0].y; cs0.z += pv[0][0].z; cs0.w += pv[0][0].w;
        cs1.x += pv[0][1].x; cs1.y += pv[0][1].y; cs1.z += pv[0][1].z; cs1.w += pv[0][1].w;
      }
      const int c = s >> 1, half = s & 1;
      const uint32_t stg = (uint32_t)c % WG_STAGES;
      if (half == 0 && c >= WG_STAGES) mbar_wait(BAR(2 + stg), (((uint32_t)c / WG_STAGES) - 1) & 1, 1);
      unsigned char* P = smem + stg * WL::STAGE;
      unsigned char* Q = P + WL::MAT;
#pragma unroll
      for (int i = 0; i < 1; ++i)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = 32 * half + hw;
          float4 b = qv[i][hh];
          if (job < 2) {  // Q = silu(z) of the rows that exist (rows past the end were loaded as zeros: silu(0) = 0)
            b = make_float4(silu_x3(b.x), silu_x3(b.y), silu_x3(b.z), silu_x3(b.w));
          }
          st_wg_f4(P, hh, r, hl, pv[i][hh], third);
          st_wg_f4(Q, hh, r, hl, b, third);
        }
      if (half == 1) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(stg));
      }
    };
    // (a third register buffer — rows requested two steps ahead — was measured and rejected: 17 warps cap the kernel at 96 registers, the
    // rotation spills and the launch gets 4 % slower; the producers wait for these loads a third of their samples,
    // profiles/r01c_ncu_backward_summary.txt: the next step is a cp.async staging ring in the shared memory the two-term slots free)
    // jobs whose P and Q rows are plain 16-byte-aligned global rows go through the cp.async ring (all of an EdgeBlock MLP's jobs, three of
    // a CellBlock MLP's four); the 3-gather mean segment, the narrow encoder / decoder operands and the six-product variant keep the
    // register double buffer
    const bool p_plain = job > 0 || (p.n_out == 128 && (p.ld_dz & 3) == 0);
    const bool q_plain = job < 2 || p.mode == TC_EDGE || (p.mode == TC_CELL && (seg == 0 || p.idx == nullptr)) ||
                         (p.mode == TC_ROWS && (p.k_in & 127) == 0 && (p.ld_in & 3) == 0);
    if (!THIRD && p_plain && q_plain) {
      unsigned char* raw = smem + WL::RAW_OFF;
      const uint32_t raw_me = smem_u32(raw) + (uint32_t)ptid * 16u;  // [buffer][k = P0 P1 Q0 Q1][thread]: conflict-free 16-byte slots
      const bool gather = p.mode == TC_EDGE && job >= 2 && seg < 2;
      auto row_of = [&](int s) { return (c0 + (s >> 1)) * WG_CHUNK + 32 * (s & 1) + hw; };
      int nidx = 0;  // EdgeBlock x' gather: the row index of step s + 1 is fetched while step s is issued
      if (gather && S > 0) nidx = __ldg(p.idx + (size_t)seg * p.rows + min(row_of(0), p.rows - 1));
      auto issue = [&](int s) {
        const int g = row_of(s);
        const uint32_t nbytes = g < p.rows ? 16u : 0u;  // rows past the end: zero fill
        const int gc = min(g, p.rows - 1);
        const float* ps = job == 0 ? p.dz + (size_t)gc * p.ld_dz : p.da + (size_t)gc * 256 + (job == 1 ? 0 : 128);
        const float* qs;
        bool q_half_only = false;
        if (job < 2) qs = p.zsave + (size_t)gc * 384 + (job == 0 ? 128 : 0);
        else if (p.mode == TC_EDGE) qs = seg < 2 ? p.src0 + (size_t)nidx * 128 : p.src1 + (size_t)gc * 128;
        else if (p.mode == TC_CELL) { qs = seg == 0 ? p.src0 + (size_t)gc * 128 : p.src1 + (size_t)gc * 64; q_half_only = seg != 0; }
        else qs = p.in + (size_t)gc * p.ld_in + 128 * seg;
        if (gather && s + 1 < S) nidx = __ldg(p.idx + (size_t)seg * p.rows + min(row_of(s + 1), p.rows - 1));
        const uint32_t dst = raw_me + (uint32_t)(s % WG_RAW_DEPTH) * WG_RAW_BYTES;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)hh * 8192u), "l"(ps + 64 * hh + 4 * hl), "r"(nbytes)
                       : "memory");
          // (mp_times = 1 CellBlock: the per-cell aggregate is 64 wide, the upper half of that segment is zero)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16384u + (uint32_t)hh * 8192u),
                       "l"(qs + (q_half_only ? 0 : 64 * hh) + 4 * hl), "r"((q_half_only && hh == 1) ? 0u : nbytes)
                       : "memory");
        }
      };
      if (0 < S) issue(0);
      cp_async_commit();
      if (1 < S) issue(1);
      cp_async_commit();
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        if (s + 2 < S) issue(s + 2);
        cp_async_commit();    // (an empty group when nothing is left keeps the group count uniform)
        cp_async_wait<2>();   // step s has landed; the slots are private to this thread: no barrier
        const unsigned char* src = raw + (size_t)(s % WG_RAW_DEPTH) * WG_RAW_BYTES + (size_t)ptid * 16;
        float4 pv[1][2], qv[1][2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          pv[0][hh] = *reinterpret_cast<const float4*>(src + hh * 8192);
          qv[0][hh] = *reinterpret_cast<const float4*>(src + 16384 + hh * 8192);
        }
        store(s, pv, qv);
      }
    } else {
      float4 pa[1][2], qa[1][2], pb[1][2], qb[1][2];
#pragma unroll 1
      for (int s = -1; s < S; s += 2) {
        if (s + 1 < S) load(s + 1, pa, qa);
        if (s >= 0) store(s, pb, qb);
        if (s + 2 < S) load(s + 2, pb, qb);
        if (s + 1 < S) store(s + 1, pa, qa);
      }
    }
    if (do_cs) {  // the two row lanes of a warp combined, then one 128-column partial per (range, warp): fixed order, no shared memory
      cs0.x += __shfl_xor_sync(0xffffffffu, cs0.x, 16); cs0.y += __shfl_xor_sync(0xffffffffu, cs0.y, 16);
      cs0.z += __shfl_xor_sync(0xffffffffu, cs0.z, 16); cs0.w += __shfl_xor_sync(0xffffffffu, cs0.w, 16);
      cs1.x += __shfl_xor_sync(0xffffffffu, cs1.x, 16); cs1.y += __shfl_xor_sync(0xffffffffu, cs1.y, 16);
      cs1.z += __shfl_xor_sync(0xffffffffu, cs1.z, 16); cs1.w += __shfl_xor_sync(0xffffffffu, cs1.w, 16);
      if (lane < 16) {
        float* d = p.csum + (((size_t)job * p.nranges + range) * WG_NPW + warp) * 128 + 4 * hl;
        *reinterpret_cast<float4*>(d) = cs0;
        *reinterpret_cast<float4*>(d + 64) = cs1;
      }
    }
  }
  if (warp == WG_NPW) {
    // ================================================================= MMA issuer: D[o][i] += sum_rows P[row][o] Q[row][i]
    if (lane == 0) {
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        const uint32_t stg = (uint32_t)c % WG_STAGES;
        mbar_wait(BAR(stg), ((uint32_t)c / WG_STAGES) & 1, 4);
        tc_fence_after();
        const uint32_t pb = smem_u32(smem + stg * WL::STAGE), qb = pb + WL::MAT;
#pragma unroll
        for (int kk = 0; kk < WG_CHUNK / 16; ++kk)
#pragma unroll
          for (int q = 0; q < (THIRD ? 6 : 3); ++q)
            if (q < p.nprod)
              umma_bf16(tmem_base, umma_desc_mn(pb + prod_a(q) * 2 * WG_HALF_BYTES + kk * 2048),
                        umma_desc_mn(qb + prod_w(q) * 2 * WG_HALF_BYTES + kk * 2048), IDESC_BF16_MN_MN, (c | kk | q) != 0);
        umma_commit(BAR(2 + stg));
      }
      umma_commit(BAR(4));
    }
    __syncwarp();
  } else if (warp < 4) {
    // ================================================================= epilogue (warps 0-3, after their producer work): lane = output row o
    float* dst = p.part + ((size_t)job * p.nranges + range) * 16384 + (size_t)(32 * warp + lane) * 128;
    if (nch > 0) {
      mbar_wait(BAR(4), 0, 8);
      tc_fence_after();
    }
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      if (nch > 0) {
        tmem_ld32_issue(t_lane + 32 * c, r);
        tmem_ld_wait(r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(dst + 32 * c + 4 * j) =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_NPW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// grads <- sum over row ranges (fixed order) of the per-CTA partial blocks, laid out as dW1[128, k_in] | dW2[128,128] | dW3[128,128]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int njobs, int nranges, int k_in, float* __restrict__ grads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (job, o, i)
  if (idx >= njobs * 16384) return;
  const int job = idx >> 14, o = (idx >> 7) & 127, i = idx & 127;
  float s = 0.f;
  for (int r = 0; r < nranges; ++r) s += part[((size_t)job * nranges + r) * 16384 + o * 128 + i];
  if (job == 0) grads[(size_t)128 * k_in + 16384 + o * 128 + i] = s;       // dW3
  else if (job == 1) grads[(size_t)128 * k_in + o * 128 + i] = s;          // dW2
  else {
    const int k = 128 * (job - 2) + i;
    if (k < k_in) grads[(size_t)o * k_in + k] = s;                          // dW1
  }
}

// deterministic column sums of up to five [rows, ncols <= 128] fp32 matrices: blockIdx.y = source, blockIdx.x = row range; a block is
// 16 row-lanes x 32 float4 column groups, every thread keeps 4 independent partial sums (fixed combination order), the row-lanes are
// combined through shared memory in a fixed order, and colsum_reduce_kernel sums the row ranges in a fixed order.
struct ColsumSrc { const float* src; int ld, ncols, dst_off; const float* pre; int pre_nb; };  // pre: partials [pre_nb][128] already formed (ln_bwd_sums_kernel)
struct ColsumParams { ColsumSrc s[5]; int rows, nb; float* partial; float* grads; };
__global__ void __launch_bounds__(512) colsum_kernel(const ColsumParams p) {
  __shared__ float sh[16][128];
  const ColsumSrc cs = p.s[blockIdx.y];
  if (cs.src == nullptr || cs.pre != nullptr) return;
  const int per = (p.rows + p.nb - 1) / p.nb, r0 = blockIdx.x * per, r1 = min(p.rows, r0 + per);
  const bool vec = cs.ncols == 128 && (cs.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(cs.src) & 15) == 0;
  if (vec) {
    // 32 float4 column groups x 16 row lanes; 4 independent 16-byte loads in flight per thread, combined in a fixed order
    const int c4 = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const float* base = cs.src + c4 * 4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    auto add4 = [](float4& a, const float4 v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; };
    int r = r0 + rl;
    for (; r + 48 < r1; r += 64) {
      const float4 v0 = ldg4(base + (size_t)r * cs.ld), v1 = ldg4(base + (size_t)(r + 16) * cs.ld);
      const float4 v2 = ldg4(base + (size_t)(r + 32) * cs.ld), v3 = ldg4(base + (size_t)(r + 48) * cs.ld);
      add4(a0, v0); add4(a1, v1); add4(a2, v2); add4(a3, v3);
    }
    for (; r < r1; r += 16) add4(a0, ldg4(base + (size_t)r * cs.ld));
    *reinterpret_cast<float4*>(&sh[rl][c4 * 4]) = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                                                              (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
  } else {
    // narrow / unaligned sources (the decoder's [rows, 5] gradient): one column per thread, 4 row lanes
    const int c = threadIdx.x & 127, rl = threadIdx.x >> 7;
    float a = 0.f;
    if (c < cs.ncols)
      for (int r = r0 + rl; r < r1; r += 4) a += cs.src[(size_t)r * cs.ld + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) sh[4 * k + rl][c] = (k == 0) ? a : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < 128 && (int)threadIdx.x < cs.ncols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += sh[k][threadIdx.x];
    p.partial[((size_t)blockIdx.y * p.nb + blockIdx.x) * 128 + threadIdx.x] = s;
  }
}
__global__ void __launch_bounds__(1024) colsum_reduce_kernel(const ColsumParams p) {
  __shared__ float sh[32][128];
  const ColsumSrc cs = p.s[blockIdx.x];
  if (cs.src == nullptr && cs.pre == nullptr) return;
  // 32 float4 column groups x 32 row lanes (up to 592 partial rows per source: 19 loads per thread), fixed combination order
  const int c4 = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const float* part = cs.pre ? cs.pre : p.partial + (size_t)blockIdx.x * p.nb * 128;
  const int nb = cs.pre ? cs.pre_nb : p.nb;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = rl; b < nb; b += 32) {
    const float4 v = *reinterpret_cast<const float4*>(part + (size_t)b * 128 + c4 * 4);  // (columns >= ncols of a narrow source: never stored)
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4*>(&sh[rl][c4 * 4]) = a;
  __syncthreads();
  if (threadIdx.x < 128 && (int)threadIdx.x < cs.ncols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += sh[k][threadIdx.x];
    p.grads[cs.dst_off + threadIdx.x] = s;
  }
}

// Per-cell mean of the three node aggregates (GN_blocks.py:125-129: (a + b) + c, then a true division by 3.0) as the split fp16 shadow
// [C][2][64] the CellBlock kernel (SH variant) moves into its third K-tile by cp.async.  16 threads per cell, 4 columns each.
__global__ void __launch_bounds__(256) cell_mean_split_kernel(const float* __restrict__ node_agg, const int32_t* __restrict__ cells_node_T, int C,
                                                              unsigned char* __restrict__ out) {
  const int gid = blockIdx.x * 256 + threadIdx.x;
  const int c = gid >> 4, hl = gid & 15;
  if (c >= C) return;
  const float4 v = cell_agg4(node_agg, cells_node_T, (size_t)C, c, hl * 4);  // cells_node_T == NULL (mp_times = 1): node_agg is per cell
  uint2 q0, q1;
  split2_pair(v.x * ASCALE, v.y * ASCALE, q0.x, q1.x);
  split2_pair(v.z * ASCALE, v.w * ASCALE, q0.y, q1.y);
  unsigned char* o = out + (size_t)c * 256 + hl * 8;
  *reinterpret_cast<uint2*>(o) = q0;
  *reinterpret_cast<uint2*>(o + 128) = q1;
}

// The same mean as plain fp32 [C][64] (training: the forward and the weight-gradient kernel then read it as a per-cell row — the
// cells_node_T = NULL convention — instead of gathering three node rows each).
__global__ void __launch_bounds__(256) cell_mean_kernel(const float* __restrict__ node_agg, const int32_t* __restrict__ cells_node_T, int C,
                                                        float* __restrict__ out) {
  const int gid = blockIdx.x * 256 + threadIdx.x;
  const int c = gid >> 4, hl = gid & 15;
  if (c >= C) return;
  *reinterpret_cast<float4*>(out + (size_t)c * 64 + hl * 4) = cell_agg4(node_agg, cells_node_T, (size_t)C, c, hl * 4);
}

// Forward weights: blockIdx.x = weight matrix, blockIdx.y = slice of its tiles (every CTA of a matrix finds the same max|W| itself:
// 196 KB at most, L2-resident).  W [n_valid, k_valid] fp32 row-major -> scale 2^s with max|W| 2^s in [2^13, 2^14)
// (found here, on the device: graph-capturable, no host round trip) -> steps [kt][2 fp16 splits][128 rows][64 k], each sub-tile a
// byte image of the swizzled layout; the epilogue factor 1 / (ASCALE 2^s) goes to the header behind the last step.
// blockIdx.z = MLP: all the MLPs whose weights an optimizer step changed are re-packed by ONE launch (33 per training step).
struct PackFwdJob { const float* w[3]; unsigned char* out; int k_in, n_out; };
struct PackFwdBatch { PackFwdJob job[PACK_MAX_JOBS]; };
__global__ void __launch_bounds__(1024) pack_fwd_h2_kernel(const __grid_constant__ PackFwdBatch batch) {
  __shared__ float s_max[32];
  __shared__ float s_scale;
  const PackFwdJob& jb = batch.job[blockIdx.z];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int kt1 = (jb.k_in + KTILE - 1) / KTILE;
  const float* W = jb.w[m];
  const int ldw = m == 0 ? jb.k_in : 128, nv = m == 2 ? jb.n_out : 128, kv = m == 0 ? jb.k_in : 128, kt = m == 0 ? kt1 : 2;
  if ((int)blockIdx.y >= (kt > 2 ? kt : 2)) return;  // the grid's y extent is that of the widest layer 1 in the batch
  struct { unsigned char* out; int nsteps; int step0[3]; } p = {jb.out, kt1 + 4, {0, kt1, kt1 + 2}};
  float mx = 0.f;
  for (int r = tid >> 5; r < nv; r += 32)  // a warp per row: coalesced, no integer division
    for (int c = tid & 31; c < kv; c += 32) mx = fmaxf(mx, fabsf(W[(size_t)r * ldw + c]));
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_max[tid >> 5] = mx;
  __syncthreads();
  if (tid < 32) {
    mx = s_max[tid];
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (tid == 0) {
      // 2^(13 - floor(log2 max)) from the exponent field; a zero / tiny (< 2^-95) / non-finite matrix is left unscaled
      const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu);
      const float sc = (e < 32 || e == 255) ? 1.0f : __uint_as_float((uint32_t)(127 + 13 - (e - 127)) << 23);
      s_scale = sc;
      if (blockIdx.y == 0) reinterpret_cast<float*>(p.out + (size_t)p.nsteps * FW_SLOT)[m] = 1.0f / (ASCALE * sc);  // exact: powers of two
    }
  }
  __syncthreads();
  const float sc = s_scale;
  unsigned char* out = p.out + (size_t)p.step0[m] * FW_SLOT;
  const int ny = kt > 2 ? kt : 2;  // CTAs of this matrix
  for (int idx = blockIdx.y * 1024 + tid; idx < kt * 1024; idx += ny * 1024) {  // (tile, n, 16-byte chunk)
    const int c = idx & 7, n = (idx >> 3) & 127, t = idx >> 10;
    uint4 q0, q1;
    uint32_t* a0 = reinterpret_cast<uint32_t*>(&q0);
    uint32_t* a1 = reinterpret_cast<uint32_t*>(&q1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = t * KTILE + c * 8 + 2 * j;
      const float x = (n < nv && k < kv) ? W[(size_t)n * ldw + k] * sc : 0.f;
      const float y = (n < nv && k + 1 < kv) ? W[(size_t)n * ldw + k + 1] * sc : 0.f;
      split2_pair(x, y, a0[j], a1[j]);
    }
    const size_t off = (size_t)t * FW_SLOT + (size_t)n * 128 + (((uint32_t)c ^ ((uint32_t)n & 7u)) << 4);
    *reinterpret_cast<uint4*>(out + off) = q0;
    *reinterpret_cast<uint4*>(out + TILE_BYTES + off) = q1;
  }
}

static int kt_of(int k_in) { return (k_in + KTILE - 1) / KTILE; }

}  // namespace x3

size_t tc3_packed_bytes(int k_in) { return (size_t)(x3::kt_of(k_in) + 4) * x3::SLOT_BYTES; }

int tc3_pack(const fvgn_mlp_t* m, void* packed, size_t bytes, cudaStream_t st) {
  using namespace x3;
  FVGN_CHECK_ARG(m && packed, "NULL pointer");
  FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128, "bad mlp dims");
  FVGN_CHECK_ARG(bytes >= tc3_packed_bytes(m->k_in), "packed buffer too small");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed buffer must be 16B aligned");
  const int kt1 = kt_of(m->k_in);
  PackFwdBatch b;
  b.job[0] = {{m->w1, m->w2, m->w3}, reinterpret_cast<unsigned char*>(packed), m->k_in, m->n_out};
  pack_fwd_h2_kernel<<<dim3(3, kt1 > 2 ? kt1 : 2, 1), 1024, 0, st>>>(b);
  FVGN_LAUNCH_CHECK();
  return 0;
}

// every MLP of the array into its own m->packed_x3 (each at least tc3_packed_bytes(k_in) large), PACK_MAX_JOBS per launch
int tc3_pack_many(const fvgn_mlp_t* mlps, int n, cudaStream_t st) {
  using namespace x3;
  FVGN_CHECK_ARG(n >= 0 && (n == 0 || mlps), "NULL pointer");
  for (int i0 = 0; i0 < n; i0 += PACK_MAX_JOBS) {
    const int nb = n - i0 < PACK_MAX_JOBS ? n - i0 : PACK_MAX_JOBS;
    PackFwdBatch b;
    int ktmax = 2;
    for (int i = 0; i < nb; ++i) {
      const fvgn_mlp_t* m = mlps + i0 + i;
      FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128, "bad mlp dims");
      FVGN_CHECK_ARG(m->packed_x3 && (reinterpret_cast<uintptr_t>(m->packed_x3) & 15) == 0, "packed_x3 must be set and 16B aligned");
      b.job[i] = {{m->w1, m->w2, m->w3}, reinterpret_cast<unsigned char*>(const_cast<void*>(m->packed_x3)), m->k_in, m->n_out};
      if (kt_of(m->k_in) > ktmax) ktmax = kt_of(m->k_in);
    }
    pack_fwd_h2_kernel<<<dim3(3, ktmax, nb), 1024, 0, st>>>(b);
    FVGN_LAUNCH_CHECK();
  }
  return 0;
}

#if FVGN_TC_TIMING_EXPERIMENTS
extern "C" __attribute__((visibility("default"))) int fvgn_dbg_x3_prof(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, x3::g_prof, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(x3::g_prof, z, sizeof(z)); }
  return 0;
}
#endif

static int launch_tc3(x3::Params& p, const fvgn_mlp_t* m, bool red, cudaStream_t st, bool sh = false) {
  using namespace x3;
  if (p.rows <= 0) return 0;
#if FVGN_TC_TIMING_EXPERIMENTS
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FVGN_TC_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
#endif
  FVGN_CHECK_ARG(m->packed_x3 != nullptr, "FVGN_MATH_BF16X3_TC needs split-packed weights: call fvgn_mlp_pack_bf16x3 first");
  p.packed = reinterpret_cast<const unsigned char*>(m->packed_x3);
  p.b1 = m->b1; p.b2 = m->b2; p.b3 = m->b3; p.ln_w = m->ln_w; p.ln_b = m->ln_b;
  p.k_in = m->k_in; p.n_out = m->n_out;
  p.nprod = m->n_products == 1 ? 1 : 3;
  p.zsave = m->z_save;
  if (p.zsave) FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(p.zsave) & 15) == 0, "z_save must be 16B aligned");
  p.kt1 = kt_of(m->k_in);
  p.ntiles = cdiv(p.rows, TM);
  const SmemLayout L = smem_layout_fwd();
  const int grid = p.ntiles < sm_count() ? p.ntiles : sm_count();
#define FVGN_TC3_LAUNCH(MODE, RED, SH)                                                                                   \
  do {                                                                                                                   \
    static bool configured = false;                                                                                      \
    if (!configured) {                                                                                                   \
      FVGN_CUDA(cudaFuncSetAttribute(fused_mlp_tc3_kernel<MODE, RED, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
      configured = true;                                                                                                 \
    }                                                                                                                    \
    fused_mlp_tc3_kernel<MODE, RED, SH><<<grid, NTHREADS, L.total, st>>>(p);                                             \
  } while (0)
  if (sh) {
    FVGN_CHECK_ARG(!p.zsave && red && p.mode != TC_ROWS, "the split-shadow (inference) variants update the latent in place and save nothing");
    if (p.mode == TC_CELL) FVGN_TC3_LAUNCH(TC_CELL, true, true);
    else FVGN_TC3_LAUNCH(TC_EDGE, true, true);
  } else if (p.mode == TC_ROWS) FVGN_TC3_LAUNCH(TC_ROWS, false, false);
  else if (p.mode == TC_CELL && red) FVGN_TC3_LAUNCH(TC_CELL, true, false);
  else if (p.mode == TC_CELL) FVGN_TC3_LAUNCH(TC_CELL, false, false);
  else if (red) FVGN_TC3_LAUNCH(TC_EDGE, true, false);
  else FVGN_TC3_LAUNCH(TC_EDGE, false, false);
#undef FVGN_TC3_LAUNCH
  FVGN_LAUNCH_CHECK();
  return 0;
}

static int check_mlp3(const fvgn_mlp_t* m, bool need_ln) {
  FVGN_CHECK_ARG(m != nullptr, "mlp is NULL");
  FVGN_CHECK_ARG(m->w1 && m->b1 && m->w2 && m->b2 && m->w3 && m->b3, "mlp weight pointer is NULL");
  if (m->ln_w || need_ln) FVGN_CHECK_ARG(m->ln_w && m->ln_b && m->n_out == 128, "LayerNorm needs weight/bias and n_out == 128");
  return 0;
}

int tc3_mlp_rows(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, float* out, int ld_out, cudaStream_t st) {
  if (int e = check_mlp3(mlp, false)) return e;
  FVGN_CHECK_ARG(mlp->k_in <= 384, "k_in too large");
  FVGN_CHECK_ARG(rows >= 0 && (rows == 0 || (in && out)), "in/out NULL");
  FVGN_CHECK_ARG(ld_in >= mlp->k_in && ld_out >= mlp->n_out, "leading dimension too small");
  if (mlp->n_out == 128 && (ld_out & 3) == 0) FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "output must be 16B aligned");
  x3::Params p{};
  p.mode = x3::TC_ROWS; p.rows = rows; p.in = in; p.ld_in = ld_in; p.out0 = out; p.ld_out = ld_out;
  return launch_tc3(p, mlp, false, st);
}

int tc3_cell_block(const fvgn_mlp_t* mlp, const float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells, float* x_prime,
                   float* x_out, cudaStream_t st) {
  if (int e = check_mlp3(mlp, true)) return e;
  FVGN_CHECK_ARG(mlp->k_in == 192, "CellBlock MLP must have k_in == 192");
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (x && node_agg)), "NULL input");  // cells_node_T == NULL: mp_times = 1
  x3::Params p{};
  p.mode = x3::TC_CELL; p.rows = n_cells; p.src0 = x; p.src1 = node_agg; p.idx = cells_node_T; p.out0 = x_prime; p.out1 = x_out; p.ld_out = 128;
  return launch_tc3(p, mlp, x_out != nullptr && x_out == x, st);
}

int tc3_edge_block(const fvgn_mlp_t* mlp, const float* x_prime, const float* e, const int32_t* neighbour_cell, int n_faces, float* e_prime,
                   float* e_out, cudaStream_t st) {
  if (int er = check_mlp3(mlp, true)) return er;
  FVGN_CHECK_ARG(mlp->k_in == 384, "EdgeBlock MLP must have k_in == 384");
  FVGN_CHECK_ARG(n_faces >= 0 && (n_faces == 0 || (x_prime && e && neighbour_cell)), "NULL input");
  x3::Params p{};
  p.mode = x3::TC_EDGE; p.rows = n_faces; p.src0 = x_prime; p.src1 = e; p.idx = neighbour_cell; p.out0 = e_prime; p.out1 = e_out; p.ld_out = 128;
  return launch_tc3(p, mlp, e_out != nullptr && e_out == e, st);
}

// ---- split-shadow (inference) variants: latents are fp32, updated in place; x' only exists as the fp16 hi/lo shadow the EdgeBlock gathers
static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

int tc3_cell_mean(const float* node_agg, const int32_t* cells_node_T, int n_cells, float* out, cudaStream_t st) {
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (node_agg && cells_node_T && out)), "NULL input");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "output must be 16B aligned");
  if (n_cells == 0) return 0;
  x3::cell_mean_kernel<<<cdiv((long long)n_cells * 16, 256), 256, 0, st>>>(node_agg, cells_node_T, n_cells, out);
  FVGN_LAUNCH_CHECK();
  return 0;
}

int tc3_cell_mean_split(const float* node_agg, const int32_t* cells_node_T, int n_cells, void* out, cudaStream_t st) {
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (node_agg && out)), "NULL input");  // cells_node_T == NULL: mp_times = 1
  FVGN_CHECK_ARG(al16(node_agg) && al16(out), "buffers must be 16B aligned");
  if (n_cells == 0) return 0;
  x3::cell_mean_split_kernel<<<cdiv(n_cells * 16, 256), 256, 0, st>>>(node_agg, cells_node_T, n_cells, reinterpret_cast<unsigned char*>(out));
  FVGN_LAUNCH_CHECK();
  return 0;
}

int tc3_cell_block_sh(const fvgn_mlp_t* mlp, float* x, const void* cell_mean_sh, int n_cells, void* x_prime_sh, cudaStream_t st) {
  if (int e = check_mlp3(mlp, true)) return e;
  FVGN_CHECK_ARG(mlp->k_in == 192, "CellBlock MLP must have k_in == 192");
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (x && cell_mean_sh && x_prime_sh)), "NULL input");
  FVGN_CHECK_ARG(al16(x) && al16(x_prime_sh) && al16(cell_mean_sh), "buffers must be 16B aligned");
  x3::Params p{};
  p.mode = x3::TC_CELL; p.rows = n_cells; p.src0 = x; p.out1 = x; p.ld_out = 128;
  p.sh_in0 = reinterpret_cast<const unsigned char*>(cell_mean_sh);
  p.sh_prime = reinterpret_cast<unsigned char*>(x_prime_sh);
  return launch_tc3(p, mlp, true, st, true);
}

int tc3_edge_block_sh(const fvgn_mlp_t* mlp, const void* x_prime_sh, float* e, const int32_t* neighbour_cell, int n_faces, cudaStream_t st) {
  if (int er = check_mlp3(mlp, true)) return er;
  FVGN_CHECK_ARG(mlp->k_in == 384, "EdgeBlock MLP must have k_in == 384");
  FVGN_CHECK_ARG(n_faces >= 0 && (n_faces == 0 || (x_prime_sh && e && neighbour_cell)), "NULL input");
  FVGN_CHECK_ARG(al16(x_prime_sh) && al16(e), "buffers must be 16B aligned");
  x3::Params p{};
  p.mode = x3::TC_EDGE; p.rows = n_faces; p.src1 = e; p.idx = neighbour_cell; p.out1 = e; p.ld_out = 128;
  p.sh_in0 = reinterpret_cast<const unsigned char*>(x_prime_sh);
  return launch_tc3(p, mlp, true, st, true);
}

// ---- backward (data-gradient chain) host side
size_t tc3_packed_bwd_bytes(int k_in) { return (size_t)(4 + 2 * ((k_in + 127) / 128)) * x3::SLOT_BYTES; }

int tc3_pack_bwd(const fvgn_mlp_t* m, void* packed, size_t bytes, cudaStream_t st) {
  using namespace x3;
  FVGN_CHECK_ARG(m && packed, "NULL pointer");
  FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128, "bad mlp dims");
  FVGN_CHECK_ARG(bytes >= tc3_packed_bwd_bytes(m->k_in), "packed buffer too small");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed buffer must be 16B aligned");
  unsigned char* o = reinterpret_cast<unsigned char*>(packed);
  const int nseg = (m->k_in + 127) / 128;
  PackTParams pp{};
  pp.kt = 2; pp.out = o;
  // dh2[., i] = sum_o dz[., o] W3[o][i]: "weight" rows n = i (128), K = o (n_out valid)
  pp.job[0] = {m->w3, 128, 0, 128, m->n_out, 0};
  pp.job[1] = {m->w2, 128, 0, 128, 128, 2};
  for (int seg = 0; seg < nseg; ++seg) {  // dA[., 128 seg + n] = sum_o da1[., o] W1[o][128 seg + n]
    const int nv = m->k_in - 128 * seg < 128 ? m->k_in - 128 * seg : 128;
    pp.job[2 + seg] = {m->w1, m->k_in, 128 * seg, nv, 128, 4 + 2 * seg};
  }
  pack_tiles3_t_kernel<<<dim3(cdiv(2 * 1024, 256), 2 + nseg), 256, 0, st>>>(pp);
  FVGN_LAUNCH_CHECK();
  return 0;
}

// every MLP of the array into its own m->packed_x3_bwd (each at least tc3_packed_bwd_bytes(k_in) large)
int tc3_pack_bwd_many(const fvgn_mlp_t* mlps, int n, cudaStream_t st) {
  using namespace x3;
  FVGN_CHECK_ARG(n >= 0 && (n == 0 || mlps), "NULL pointer");
  for (int i0 = 0; i0 < n; i0 += PACK_MAX_JOBS) {
    const int nb = n - i0 < PACK_MAX_JOBS ? n - i0 : PACK_MAX_JOBS;
    PackTBatch b;
    for (int i = 0; i < nb; ++i) {
      const fvgn_mlp_t* m = mlps + i0 + i;
      FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128, "bad mlp dims");
      FVGN_CHECK_ARG(m->packed_x3_bwd && (reinterpret_cast<uintptr_t>(m->packed_x3_bwd) & 15) == 0, "packed_x3_bwd must be set and 16B aligned");
      b.m[i] = {m->w1, m->w2, m->w3, reinterpret_cast<unsigned char*>(const_cast<void*>(m->packed_x3_bwd)), m->k_in, m->n_out};
    }
    pack_tiles3_t_many_kernel<<<dim3(cdiv(2 * 1024, 256), 5, nb), 256, 0, st>>>(b);
    FVGN_LAUNCH_CHECK();
  }
  return 0;
}

int tc3_ln_bwd(const float* d_out, int ld_dout, const float* z_save, const float* ln_w, int rows, float* dz, float* gxhat, cudaStream_t st) {
  FVGN_CHECK_ARG(d_out && z_save && ln_w && dz, "NULL pointer");
  FVGN_CHECK_ARG((ld_dout & 3) == 0, "d_out row stride must be a multiple of 4 floats");
  if (rows <= 0) return 0;
  // `gxhat` is the column-sum partial buffer [3][ln_bwd_blocks(rows)][128] (dz | d_out * xhat | d_out), or NULL for dz alone
  if (gxhat) x3::ln_bwd_sums_kernel<<<x3::ln_bwd_blocks(rows), 256, 0, st>>>(d_out, ld_dout, z_save, ln_w, rows, dz, gxhat);
  else x3::ln_bwd_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(d_out, ld_dout, z_save, ln_w, rows, dz, nullptr);
  FVGN_LAUNCH_CHECK();
  return 0;
}

int tc3_dgrad(const fvgn_mlp_t* m, const float* dz, int ld_dz, int rows, float* da, float* dA, int ld_dA, const float* dA_base, int ld_base,
              cudaStream_t st) {
  using namespace x3;
  FVGN_CHECK_ARG(m && dz && da, "NULL pointer");
  FVGN_CHECK_ARG(m->z_save != nullptr, "the data-gradient chain needs the saved pre-activations (fvgn_mlp_t.z_save)");
  FVGN_CHECK_ARG(m->packed_x3_bwd != nullptr, "call fvgn_mlp_pack_bf16x3_bwd first");
  FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128 && ld_dz >= m->n_out, "bad dims");
  FVGN_CHECK_ARG(!dA || ((m->k_in & 15) == 0 && (ld_dA & 3) == 0 && ld_dA >= m->k_in), "dA needs k_in % 16 == 0 and a 16B-aligned row stride");
  FVGN_CHECK_ARG(!dA_base || (ld_base & 3) == 0, "dA_base row stride must be a multiple of 4 floats");
  if (rows <= 0) return 0;
  DgParams p{};
  p.rows = rows; p.ntiles = cdiv(rows, TM); p.nseg = (m->k_in + 127) / 128; p.k_in = m->k_in; p.n_out = m->n_out; p.want_dA = dA ? 1 : 0;
  p.dz = dz; p.ld_dz = ld_dz; p.zsave = m->z_save; p.da = da; p.dA = dA; p.ld_dA = ld_dA; p.dA_base = dA_base; p.ld_base = ld_base;
  p.packed = reinterpret_cast<const unsigned char*>(m->packed_x3_bwd);
  p.nprod = m->n_products == 1 ? 1 : (m->n_products == 3 ? 3 : 6);
  const SmemLayout L = smem_layout();
  static bool configured = false;
  if (!configured) {
    FVGN_CUDA(cudaFuncSetAttribute(dgrad_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    FVGN_CUDA(cudaFuncSetAttribute(dgrad_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    configured = true;
  }
  const int grid = p.ntiles < sm_count() ? p.ntiles : sm_count();
  if (p.nprod > 3) dgrad_tc3_kernel<true><<<grid, NTHREADS, L.total, st>>>(p);
  else dgrad_tc3_kernel<false><<<grid, NTHREADS, L.total, st>>>(p);
  FVGN_LAUNCH_CHECK();
  return 0;
}


// ---- weight gradients on the tensor cores
static int wg_njobs(int k_in) { return 2 + (k_in + 127) / 128; }
static int wg_nranges(int k_in, int rows) {
  const int nchunks = cdiv(rows, x3::WG_CHUNK), by_sm = sm_count() / wg_njobs(k_in);
  return nchunks < by_sm ? nchunks : by_sm;
}
size_t tc3_ln_bwd_sums_floats(int rows) { return (size_t)3 * x3::ln_bwd_blocks(rows) * 128; }
constexpr int WG_COLSUM_NB = 592;  // at most this many row ranges (>= 256 rows each) for the column sums
size_t tc3_wgrad_workspace_floats(int k_in) {
  return (size_t)wg_njobs(k_in) * (size_t)(sm_count() / wg_njobs(k_in)) * 16384 + (size_t)5 * WG_COLSUM_NB * 128;
}

int tc3_wgrad(const fvgn_mlp_t* m, int mode, const float* in, int ld_in, const float* src0, const float* src1, const int32_t* idx, int rows,
              const float* d_out, int ld_dout, const float* dz, const float* gxhat, float* grads, float* workspace, cudaStream_t st) {
  using namespace x3;
  FVGN_CHECK_ARG(m && d_out && grads && workspace, "NULL pointer");
  FVGN_CHECK_ARG(m->z_save && m->da_in, "the tensor-core weight gradients need fvgn_mlp_t.z_save and .da_in (fvgn_mlp_dgrad_x3)");
  FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128, "bad mlp dims");
  FVGN_CHECK_ARG(mode == TC_ROWS ? in != nullptr : (src0 && src1 && (idx || mode == TC_CELL)), "layer input pointers");  // CELL, idx == NULL: mp_times = 1
  FVGN_CHECK_ARG(!m->ln_w || (dz && gxhat), "an MLP with LayerNorm needs dz and gxhat from fvgn_ln_bwd");
  FVGN_CHECK_ARG(rows > 0, "rows must be > 0");
  WgParams p{};
  p.mode = mode; p.rows = rows; p.k_in = m->k_in; p.n_out = m->n_out;
  p.njobs = wg_njobs(m->k_in); p.nchunks = cdiv(rows, WG_CHUNK); p.nranges = wg_nranges(m->k_in, rows);
  p.chunks_per_range = cdiv(p.nchunks, p.nranges);
  p.dz = m->ln_w ? dz : d_out; p.ld_dz = m->ln_w ? 128 : ld_dout;
  p.da = m->da_in; p.zsave = m->z_save; p.in = in; p.ld_in = ld_in; p.src0 = src0; p.src1 = src1; p.idx = idx;
  p.part = workspace;
  p.csum = workspace + (size_t)p.njobs * (size_t)(sm_count() / p.njobs) * 16384;  // (the column-sum kernel's partial region: it no longer runs)
  p.csum_job0 = m->ln_w ? 0 : 1;
  p.nprod = m->n_products == 1 ? 1 : (m->n_products == 3 ? 3 : 6);
  const int smem_bytes = p.nprod > 3 ? WgLayout<true>::TOTAL : WgLayout<false>::TOTAL;
  static bool configured = false;
  if (!configured) {
    FVGN_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgLayout<true>::TOTAL));
    FVGN_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgLayout<false>::TOTAL));
    configured = true;
  }
  if (p.nprod > 3) wgrad_tc3_kernel<true><<<p.njobs * p.nranges, WG_NTHREADS, smem_bytes, st>>>(p);
  else wgrad_tc3_kernel<false><<<p.njobs * p.nranges, WG_NTHREADS, smem_bytes, st>>>(p);
  FVGN_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<cdiv(p.njobs * 16384, 256), 256, 0, st>>>(workspace, p.njobs, p.nranges, m->k_in, grads);
  FVGN_LAUNCH_CHECK();
  // bias / LayerNorm-parameter gradients: deterministic column sums.  grads tail: db1 | db2 | db3 | dLN_w | dLN_b
  const int vec0 = 128 * m->k_in + 2 * 16384;
  ColsumParams c{};
  // ~4 resident 512-thread blocks per SM over the 5 sources, at least 256 rows per block
  c.rows = rows; c.nb = cdiv(rows, 256) < (4 * sm_count()) / 2 ? cdiv(rows, 256) : (4 * sm_count()) / 2; c.grads = grads;
  if (c.nb > WG_COLSUM_NB) c.nb = WG_COLSUM_NB;
  c.partial = workspace + (size_t)p.njobs * (size_t)(sm_count() / p.njobs) * 16384;
  // db2 = colsum(da2), db1 = colsum(da1) (and db3 = colsum(d_out) without LayerNorm): per-(range, warp) partials formed by the producers of
  // weight-gradient jobs 1, 2 (and 0) as the rows passed through them
  const int nr16 = p.nranges * WG_NPW;
  const size_t cs_stride = (size_t)nr16 * 128;
  c.s[0] = {nullptr, 0, 128, vec0 + 128, p.csum + cs_stride, nr16};
  c.s[1] = {nullptr, 0, 128, vec0, p.csum + 2 * cs_stride, nr16};
  if (m->ln_w) {
    // db3 = colsum(dz), dLN_w = colsum(d_out * xhat), dLN_b = colsum(d_out): per-block partials formed by fvgn_ln_bwd (`gxhat` = that buffer)
    const int lnb = ln_bwd_blocks(rows);
    c.s[2] = {nullptr, 0, 128, vec0 + 256, gxhat, lnb};
    c.s[3] = {nullptr, 0, 128, vec0 + 384, gxhat + (size_t)lnb * 128, lnb};
    c.s[4] = {nullptr, 0, 128, vec0 + 512, gxhat + (size_t)2 * lnb * 128, lnb};
  } else {
    c.s[2] = {nullptr, 0, m->n_out, vec0 + 256, p.csum, nr16};      // db3 = colsum(d_out)
    c.s[3] = {nullptr, 0, 0, 0, nullptr, 0};
    c.s[4] = {nullptr, 0, 0, 0, nullptr, 0};
  }
  colsum_reduce_kernel<<<5, 1024, 0, st>>>(c);
  FVGN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fvgn
