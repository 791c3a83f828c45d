// mlp_tc3_bwd.cu — the tensor-core backward of one build_mlp in the fp32-equivalent mode (training with the bf16x3 forward):
// LayerNorm backward (+ column sums), the data-gradient chain, the weight gradients and their deterministic reductions, and the
// transposed weight packing.  Shares tc3_common.cuh with the forward kernels of mlp_tc3.cu.
#include "tc3_common.cuh"
#ifndef FVGN_BWD_COLLECTOR
#define FVGN_BWD_COLLECTOR 0  // dgrad: A-collector reuse between the two split products of a K slice that share their A operand (measured: +0.08 ms on the 10.1 ms training step, so off; the forward kernels gain 1 % from the same hint)
#endif

namespace fvgn {
namespace x3 {

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
    // (all 32 lanes walk the sequence in uniform control flow, one elected lane issues: tc_common.cuh::umma_bf16_w)
    if (n_local > 0) {
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
                if (q < p.nprod) {
                  // products 0 and 1 share their A operand (the leading split term): the second takes it from the A collector
                  const uint64_t ad = umma_desc(a_base + prod_a(q) * TILE_BYTES + kk * 32), bd = umma_desc(b_base + prod_w(q) * TILE_BYTES + kk * 32);
                  if (FVGN_BWD_COLLECTOR && q == 0 && p.nprod >= 2) umma_bf16_w_col<1>(d, ad, bd, IDESC_BF16_M128_N128, (t | kk | q) != 0);
                  else if (FVGN_BWD_COLLECTOR && q == 1) umma_bf16_w_col<2>(d, ad, bd, IDESC_BF16_M128_N128, 1);
                  else umma_bf16_w(d, ad, bd, IDESC_BF16_M128_N128, (t | kk | q) != 0);
                }
            umma_commit_w(BAR(B_AEMPTY + sa));
            ++ak;
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int q = 0; q < (THIRD ? 6 : 3); ++q)
                if (q < p.nprod) {
                  const uint32_t ta = tmem_base + TMEM_H0 + 64u * prod_a(q) + 8u * (uint32_t)(4 * t + kk);
                  const uint64_t bd = umma_desc(b_base + prod_w(q) * TILE_BYTES + kk * 32);
                  if (FVGN_BWD_COLLECTOR && q == 0 && p.nprod >= 2) umma_bf16_ts_w_col<1>(d, ta, bd, IDESC_BF16_M128_N128, (t | kk | q) != 0);
                  else if (FVGN_BWD_COLLECTOR && q == 1) umma_bf16_ts_w_col<2>(d, ta, bd, IDESC_BF16_M128_N128, 1);
                  else umma_bf16_ts_w(d, ta, bd, IDESC_BF16_M128_N128, (t | kk | q) != 0);
                }
          }
          umma_commit_w(BAR(B_WEMPTY + sw));
        }
      };
#pragma unroll 1
      for (int jl = 0; jl < n_local; ++jl) {
        acquire(0);
        w_steps(tmem_base, true);  // S1: dz . W3
        umma_commit_w(BAR(B_DFULL + 0));
        mbar_wait(BAR(B_HREADY), hphase, 7); hphase ^= 1;
        acquire(1);
        w_steps(tmem_base + 128u, false);  // S2: da2 . W2
        umma_commit_w(BAR(B_DFULL + 1));
        mbar_wait(BAR(B_HREADY), hphase, 7); hphase ^= 1;
#pragma unroll 1
        for (int seg = 0; seg < nseg3; ++seg) {  // S3: da1 . W1[:, seg]
          acquire(seg & 1);
          w_steps(tmem_base + 128u * (uint32_t)(seg & 1), false);
          umma_commit_w(BAR(B_DFULL + (seg & 1)));
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
    {  // (all 32 lanes walk the sequence in uniform control flow, one elected lane issues)
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
              umma_bf16_w(tmem_base, umma_desc_mn(pb + prod_a(q) * 2 * WG_HALF_BYTES + kk * 2048),
                        umma_desc_mn(qb + prod_w(q) * 2 * WG_HALF_BYTES + kk * 2048), IDESC_BF16_MN_MN, (c | kk | q) != 0);
        umma_commit_w(BAR(2 + stg));
      }
      umma_commit_w(BAR(4));
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

}  // namespace x3

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
  static ::fvgn::PerDeviceOnce configured;
  if (configured.first()) {
    FVGN_CUDA(cudaFuncSetAttribute(dgrad_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    FVGN_CUDA(cudaFuncSetAttribute(dgrad_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
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
  static ::fvgn::PerDeviceOnce configured;
  if (configured.first()) {
    FVGN_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgLayout<true>::TOTAL));
    FVGN_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgLayout<false>::TOTAL));
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
