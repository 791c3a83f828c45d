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
// Bias, SiLU (ex2.approx / rcp.approx), LayerNorm (shifted sums in one read of the accumulator, the two column halves combined pairwise) and the residual are fp32 in the epilogue, so results stay within the
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
#include "tc3_common.cuh"
#ifndef FVGN_COLLECTOR
#define FVGN_COLLECTOR 3  // A-collector reuse between the two products of a K slice that share their A operand (UTCHMMA .A_KEEP / .A_REUSE): bit 0 = layer 1 (A in shared memory), bit 1 = layers 2 / 3 (A in TMEM)
#endif

namespace fvgn {
namespace x3 {

// Device status word (one per device: a __device__ variable is instantiated per context).  bit 0: a row of a fp32-equivalent forward
// tile came out non-finite — the fp16 operand terms overflow for |activation| >= ~8190 (tc3_common.cuh: ASCALE), inf turns the row into
// NaN, and NaN reaches the row's last-layer statistics.  One compare per row and launch; read / cleared by fvgn_status_read.
__device__ unsigned int g_status = 0;

// the four 16-column sub-chunks of this thread's column half (t_acc, h, r in scope): fn(sc, registers).  (Keeping the TMEM load of
// sub-chunk sc + 1 in flight while fn works on sc was measured and rejected: profiles/r02_forward_epilogue_variants.txt.)
#define FVGN_FOR_SUBCHUNKS(fn)                                                      \
  do {                                                                              \
    _Pragma("unroll 1") for (int sc_ = 4 * h; sc_ < 4 * h + 4; ++sc_) {             \
      tmem_ld16_issue(t_acc + 16 * sc_, r);                                         \
      tmem_ld16_wait(r);                                                            \
      fn(sc_, r);                                                                   \
    }                                                                               \
  } while (0)

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
    // (all 32 lanes walk the sequence — uniform control flow keeps the descriptors in uniform registers — and one elected lane issues)
    if (n_local > 0) {
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
          const uint32_t alo = umma_desc_lo(smem_u32(sA + sa * FA_SLOT)), blo = umma_desc_lo(smem_u32(sW + sw * FW_SLOT));
#if FVGN_COLLECTOR & 1
          if (p.nprod == 3) {
            // products 0 and 1 of a K slice share their A operand (a_hi): the first MMA keeps it in the A collector, the second takes it there
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a0 = alo + (uint32_t)((kk * 32) >> 4), a1 = alo + (uint32_t)((TILE_BYTES + kk * 32) >> 4);
              const uint32_t b0 = blo + (uint32_t)((kk * 32) >> 4), b1 = blo + (uint32_t)((TILE_BYTES + kk * 32) >> 4);
              umma_f16_lo_w_col<1>(d, a0, b0, IDESC_F16_M128_N128, (t | kk) != 0);
              umma_f16_lo_w_col<2>(d, a0, b1, IDESC_F16_M128_N128, 1);
              umma_f16_lo_w_col<0>(d, a1, b0, IDESC_F16_M128_N128, 1);
            }
          } else
#endif
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < p.nprod && (!(FVGN_DBG(p) & 8) || q == 0))
                umma_f16_lo_w(d, alo + (uint32_t)((fprod_a(q) * TILE_BYTES + kk * 32) >> 4), blo + (uint32_t)((fprod_w(q) * TILE_BYTES + kk * 32) >> 4),
                              IDESC_F16_M128_N128, (t | kk | q) != 0);
          umma_commit_w(BAR(FB_AEMPTY + sa));
          umma_commit_w(BAR(FB_WEMPTY + sw));
          PROF_ADD(3);
        }
        umma_commit_w(BAR(FB_DFULL + a));
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
            const uint32_t blo = umma_desc_lo(smem_u32(sW + sw * FW_SLOT));
#if FVGN_COLLECTOR >= 2
            if (p.nprod == 3) {
              // products 0 and 1 of a K slice share h_hi: read from TMEM once, the second MMA takes it from the A collector
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint32_t a0 = hcol + 8u * (uint32_t)(4 * t + kk), a1 = a0 + 64u;
                const uint32_t b0 = blo + (uint32_t)((kk * 32) >> 4), b1 = blo + (uint32_t)((TILE_BYTES + kk * 32) >> 4);
                umma_f16_ts_lo_w_col<1>(d, a0, b0, IDESC_F16_M128_N128, (t | kk) != 0);
                umma_f16_ts_lo_w_col<2>(d, a0, b1, IDESC_F16_M128_N128, 1);
                umma_f16_ts_lo_w_col<0>(d, a1, b0, IDESC_F16_M128_N128, 1);
              }
            } else
#endif
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int q = 0; q < 3; ++q)
                if (q < p.nprod && (!(FVGN_DBG(p) & 8) || q == 0))
                  umma_f16_ts_lo_w(d, hcol + 64u * fprod_a(q) + 8u * (uint32_t)(4 * t + kk), blo + (uint32_t)((fprod_w(q) * TILE_BYTES + kk * 32) >> 4),
                                   IDESC_F16_M128_N128, (t | kk | q) != 0);
            umma_commit_w(BAR(FB_WEMPTY + sw));
            PROF_ADD(6);
          }
          umma_commit_w(BAR(FB_DFULL23 + a));
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
      if (lane == 0) {
        PROF_FLUSH(0);
#if FVGN_TC_TIMING_EXPERIMENTS
        atomicAdd(&g_prof[7], (unsigned long long)n_local);
#endif
      }
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
        // one 16-column sub-chunk: accumulator registers r -> bias + SiLU -> fp16 split -> this tile's H columns
        auto sub = [&](int sc, uint32_t* r) {
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
              split2_pair(silu_x3_scaled(z0), silu_x3_scaled(z1), p0[2 * j4], p1[2 * j4]);
              split2_pair(silu_x3_scaled(z2), silu_x3_scaled(z3), p0[2 * j4 + 1], p1[2 * j4 + 1]);
            }
            if (p.zsave) store_rows16(reinterpret_cast<const float*>(r), p.zsave, 384, 128 * layer + 16 * sc, row0 + 32 * q);
            tmem_st8(t_ha + 8 * sc, p0);
            tmem_st8(t_ha + 64 + 8 * sc, p1);
        };
        FVGN_FOR_SUBCHUNKS(sub);
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
      if (has_ln && !(FVGN_DBG(p) & 32)) {
        // LayerNorm statistics in ONE read of the accumulator: shifted sums of this thread's 64 columns (shift = its first value, so the
        // subtraction s2 - s1^2/64 cancels nothing that matters), then the two column halves of a row are combined with the pairwise
        // (Chan) update — mean and variance to fp32 rounding, like the two-pass form it replaces.  Each accumulator read-out costs the
        // MMAs TMEM-port time (profiles/r02_tmem_port_sharing.txt): three reads per tile in this layer are now two.
        float s1 = 0.f, s2 = 0.f, c = 0.f;
        auto stat16 = [&](int sc, const uint32_t* r) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(b3 + 16 * sc + 4 * j4);
            const float y0 = fmaf(__uint_as_float(r[4 * j4]), inv3, b.x), y1 = fmaf(__uint_as_float(r[4 * j4 + 1]), inv3, b.y);
            const float y2 = fmaf(__uint_as_float(r[4 * j4 + 2]), inv3, b.z), y3 = fmaf(__uint_as_float(r[4 * j4 + 3]), inv3, b.w);
            if (j4 == 0 && sc == 4 * h) c = y0;
            float d;
            d = y0 - c; s1 += d; s2 = fmaf(d, d, s2);
            d = y1 - c; s1 += d; s2 = fmaf(d, d, s2);
            d = y2 - c; s1 += d; s2 = fmaf(d, d, s2);
            d = y3 - c; s1 += d; s2 = fmaf(d, d, s2);
          }
        };
        FVGN_FOR_SUBCHUNKS(stat16);
        const float mh = c + s1 * (1.0f / 64.0f);             // mean of this half
        const float m2 = s2 - s1 * s1 * (1.0f / 64.0f);       // sum of squared deviations from it
        float* red = sRed + 512 * acc;  // (double-buffered by tile parity: one barrier per tile orders writes and reads of a buffer)
        red[h * 128 + erow] = mh;
        red[256 + h * 128 + erow] = m2;
        named_bar_sync(1, NEPI);
        const float m0 = red[erow], m1 = red[128 + erow];
        const float dm = m0 - m1;
        mean = 0.5f * (m0 + m1);
        const float var = (red[256 + erow] + red[256 + 128 + erow] + dm * dm * 32.0f) * (1.0f / 128.0f);
        rstd = 1.0f / sqrtf((var < 0.0f ? 0.0f : var) + 1e-5f);  // (a NaN variance stays NaN: the status check below needs it)
        if (h == 0 && !(rstd * 0.0f == 0.0f) && row0 + erow < p.rows) atomicOr(&g_status, 1u);  // inf / NaN anywhere in the row
      }
      const float* res = (MODE == TC_CELL) ? p.src0 : p.src1;
      auto out16 = [&](int sc, uint32_t* r) {
        float* v = reinterpret_cast<float*>(r);
        if (FVGN_DBG(p) & 32) return;
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
            float acc0 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = 16 * sc + j;
              if (col < p.n_out) { p.out0[(size_t)grow * p.ld_out + col] = v[j]; acc0 += v[j] * 0.0f; }
            }
            if (!has_ln && !(acc0 == 0.0f)) atomicOr(&g_status, 1u);  // decoder rows (no LayerNorm statistics to look at)
          }
        }
      };
      FVGN_FOR_SUBCHUNKS(out16);
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

// tri / quad hybrid meshes: the same two means over a [4,C] table (cell_agg4_poly); the fused kernels then take the per-cell row
// (cells_node_T = NULL convention), so every math mode handles quad cells without a k-wide gather of its own
__global__ void __launch_bounds__(256) cell_mean_poly_kernel(const float* __restrict__ node_agg, const int32_t* __restrict__ cells_node4_T, int C,
                                                             float* __restrict__ out, unsigned char* __restrict__ out_split) {
  const int gid = blockIdx.x * 256 + threadIdx.x;
  const int c = gid >> 4, hl = gid & 15;
  if (c >= C) return;
  const float4 v = cell_agg4_poly(node_agg, cells_node4_T, (size_t)C, c, hl * 4);
  if (out) *reinterpret_cast<float4*>(out + (size_t)c * 64 + hl * 4) = v;
  if (out_split) {
    uint2 q0, q1;
    split2_pair(v.x * ASCALE, v.y * ASCALE, q0.x, q1.x);
    split2_pair(v.z * ASCALE, v.w * ASCALE, q0.y, q1.y);
    unsigned char* o = out_split + (size_t)c * 256 + hl * 8;
    *reinterpret_cast<uint2*>(o) = q0;
    *reinterpret_cast<uint2*>(o + 128) = q1;
  }
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

}  // namespace x3


}  // namespace fvgn
#include "mlp_tc3_pair.cuh"
namespace fvgn {

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

// synchronous: waits for the device, returns the status word and (optionally) clears it
int tc3_status_read(unsigned int* out_host, int clear) {
  FVGN_CHECK_ARG(out_host != nullptr, "NULL pointer");
  FVGN_CUDA(cudaDeviceSynchronize());
  FVGN_CUDA(cudaMemcpyFromSymbol(out_host, x3::g_status, sizeof(unsigned int)));
  if (clear && *out_host) {
    const unsigned int z = 0;
    FVGN_CUDA(cudaMemcpyToSymbol(x3::g_status, &z, sizeof(z)));
  }
  return 0;
}

// 1: the inference kernels of the fp32-equivalent mode run on CTA pairs (mlp_tc3_pair.cuh: tcgen05.mma.cta_group::2).  Default 0 — measured
// (profiles/r02_forward_cta_pairs.txt): bit-identical, 1 % faster on the 1M-cell mesh, 9 % slower at cylinder-batch size; FVGN_X3_PAIR=1 in
// the environment or fvgn_x3_pair_enable(1) selects them
static int g_pair = -1;
int tc3_pair_enable(int on) {
  const int prev = g_pair < 0 ? 0 : g_pair;
  g_pair = on ? 1 : 0;
  return prev;
}

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
    static ::fvgn::PerDeviceOnce configured;                                                                                      \
    if (configured.first()) {                                                                                                   \
      FVGN_CUDA(cudaFuncSetAttribute(fused_mlp_tc3_kernel<MODE, RED, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
    }                                                                                                                    \
    fused_mlp_tc3_kernel<MODE, RED, SH><<<grid, NTHREADS, L.total, st>>>(p);                                             \
  } while (0)
  if (sh) {
    FVGN_CHECK_ARG(!p.zsave && red && p.mode != TC_ROWS, "the split-shadow (inference) variants update the latent in place and save nothing");
    if (g_pair < 0) { const char* e = getenv("FVGN_X3_PAIR"); g_pair = (e && e[0] == '1') ? 1 : 0; }
    if (g_pair && !FVGN_DBG(p)) {
      const int rc = tc3_launch_pair(p, st);
      if (rc != FVGN_E_UNSUPPORTED) return rc;
    }
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

int tc3_cell_mean_poly(const float* node_agg, const int32_t* cells_node4_T, int n_cells, float* out, void* out_split, cudaStream_t st) {
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (node_agg && cells_node4_T && (out || out_split))), "NULL input");
  FVGN_CHECK_ARG(al16(node_agg) && al16(out) && al16(out_split), "buffers must be 16B aligned");
  if (n_cells == 0) return 0;
  x3::cell_mean_poly_kernel<<<cdiv((long long)n_cells * 16, 256), 256, 0, st>>>(node_agg, cells_node4_T, n_cells, out,
                                                                                reinterpret_cast<unsigned char*>(out_split));
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

}  // namespace fvgn
