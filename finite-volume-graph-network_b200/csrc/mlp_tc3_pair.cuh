// mlp_tc3_pair.cuh — the two INFERENCE kernels of the fp32-equivalent mode (FVGN_MATH_BF16X3_TC, split-shadow variants of the CellBlock and
// the EdgeBlock: GN_blocks.py:16-145 fused with their build_mlp + LayerNorm + residual) on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// Why (profiles/r02_forward_two_epilogue_groups.txt): the single-CTA kernel (mlp_tc3.cu) is bound by the SM's shared-memory port — per
// 128-row EdgeBlock tile 120 SS / TS-mode MMAs read 4 KB of A and 4 KB of B each, the producers write 192 KB of A tiles, the loader
// 320 KB of split weights, the epilogue stages 128 KB: ~1.4 MB per tile against 128 B/clk.  A 2-CTA MMA (M = 256: 128 rows per CTA,
// N = 128) takes HALF of every B operand from each CTA of the pair: each CTA loads and holds only rows [64 r, 64 r + 64) of every weight
// K-tile (r = its rank in the cluster), so the weight ring writes and the B-operand reads per CTA halve (~1.0 MB per tile).
//
// Structure = mlp_tc3.cu's SH kernels, per CTA: 8 epilogue warps, 8 producer warps, a weight-loader thread — all on the CTA's own rows,
// shared memory, TMEM and mbarriers (identical offsets in both CTAs) — and one thread in the MMA warp:
//   rank 0: the ISSUER of the pair's MMAs.  Before a K-tile step it waits for its own A slot / weight slot / hidden activations AND for
//           the peer's (one arrival per step on its PEER barrier ring); tcgen05.commit ... multicast::cluster then arrives on the slot /
//           accumulator barriers of BOTH CTAs.
//   rank 1: a RELAY: walks the issuer's step sequence, waits for the same preconditions in its own CTA and arrives (release.cluster) on
//           the issuer's PEER barrier.
// Arithmetic, tile order per CTA and results are those of the single-CTA kernels (same MMAs on the same operands): bit-identical
// (tests/test_gpu_x3.py::test_x3_pair_kernels_are_bit_identical).
// (included by mlp_tc3.cu: same translation unit as the single-CTA kernels and the device status word)
#pragma once
#include "tc3_common.cuh"

namespace fvgn {
namespace x3 {

constexpr int PW_HALF = TILE_BYTES / 2;      // one fp16 plane of half a weight K-tile: [64 rows][128 B] = 8 KB
constexpr int PW_SLOT = 2 * PW_HALF;         // hi | lo plane = 16 KB
#ifndef FVGN_PARING
#define FVGN_PARING 3
#endif
constexpr int PARING = FVGN_PARING;          // A-ring depth (the halved weight ring leaves room for a fourth 32 KB slot)
constexpr int PWRING = 4, PEER_DEPTH = 8;
enum { PB_AFULL = 0, PB_AEMPTY = PB_AFULL + PARING, PB_WFULL = PB_AEMPTY + PARING, PB_WEMPTY = PB_WFULL + PWRING, PB_DFULL = PB_WEMPTY + PWRING,
       PB_DFULL23 = PB_DFULL + 2, PB_HREADY = PB_DFULL23 + 2, PB_PEER = PB_HREADY + 2, PB_COUNT = PB_PEER + PEER_DEPTH };
static_assert(PB_COUNT * 8 + 4 <= 256, "barrier block");
// kind::f16, fp16 operands, M = 256 (bits [24,29) = M >> 4), N = 128: the pair's MMA
constexpr uint32_t IDESC_F16_M256_N128 = (IDESC_F16_M128_N128 & ~(0x1fu << 24)) | (16u << 24);

__host__ __device__ inline SmemLayout smem_layout_pair() {
  SmemLayout L;
  uint32_t o = 0;
  L.a_off = o; o += PARING * FA_SLOT;   // 96 KB (128 KB with four slots)
  L.w_off = o; o += PWRING * PW_SLOT;   // 64 KB
  L.stage_off = o; o += 8 * 2048;
  L.par_off = o; o += 5 * 128 * 4 + 16;
  L.red_off = o; o += 2 * 2 * 2 * 128 * 4;
  L.bar_off = o; o += 256;
  L.total = o;
  return L;
}

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], both CTAs of the pair (A: 128 rows per CTA at the same shared-memory offset; B: 64 of the 128 rows per CTA)
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// (whole-warp forms with low-word descriptors: see tc_common.cuh::umma_f16_lo_w)
__device__ __forceinline__ void umma2_f16_lo_w(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, {%6, %6, %6, %6, %6, %6, %6, %6}, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(0x40004040u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ts_lo_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, {%6, %6, %6, %6, %6, %6, %6, %6}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(idesc), "r"(accumulate), "r"(0x40004040u), "r"(0u)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have completed (one elected lane)
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int who) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity, who);
  }
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) fused_mlp_tc3_pair_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const SmemLayout L = smem_layout_pair();
  unsigned char* sA = smem + L.a_off;
  unsigned char* sW = smem + L.w_off;
  float* sPar = reinterpret_cast<float*>(smem + L.par_off);
  const uint32_t bar0 = smem_u32(smem + L.bar_off);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * PB_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  constexpr int kt1 = (MODE == TC_EDGE) ? 6 : 3;
  // both CTAs of a pair walk the same number of tiles (the leader's): the peer's last tile may lie beyond the rows (loads clamp, stores are guarded)
  const int lead = (int)blockIdx.x & ~1;
  const int n_local = (lead < p.ntiles) ? (p.ntiles - lead + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("fvgn mlp_tc3_pair: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < PARING; ++i) { mbar_init(BAR(PB_AFULL + i), NPW); mbar_init(BAR(PB_AEMPTY + i), 1); }
    for (int i = 0; i < PWRING; ++i) { mbar_init(BAR(PB_WFULL + i), 1); mbar_init(BAR(PB_WEMPTY + i), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(BAR(PB_DFULL + a), 1); mbar_init(BAR(PB_DFULL23 + a), 1); mbar_init(BAR(PB_HREADY + a), 8); }
    for (int i = 0; i < PEER_DEPTH; ++i) mbar_init(BAR(PB_PEER + i), 1);
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
  if (warp >= WARP_PROD0 && warp < WARP_MMA) {
    for (int i = tid - NEPI; i < 5 * 128; i += NPROD) {
      const int which = i >> 7, c = i & 127;
      float v = 0.f;
      if (which == 0) v = p.b1[c];
      else if (which == 1) v = p.b2[c];
      else if (which == 2) v = p.b3[c];
      else if (which == 3) v = p.ln_w[c];
      else v = p.ln_b[c];
      sPar[i] = v;
    }
    if (tid - NEPI < 3) sPar[5 * 128 + tid - NEPI] = reinterpret_cast<const float*>(p.packed + (size_t)(kt1 + 4) * FW_SLOT)[tid - NEPI];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised and its TMEM allocated before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= WARP_PROD0 && warp < WARP_MMA) {
    // ================================================================= producers (this CTA's rows -> this CTA's A ring), as mlp_tc3.cu
    constexpr int NHW = NPROD / 16;
    const int ptid = tid - NEPI;
    const int hw = ptid >> 4, hl = ptid & 15;
    const int plane = hl >> 3, chunk = hl & 7;
    const int last = p.rows - 1;
    auto publish = [&](int s) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(PB_AFULL + (uint32_t)s % PARING));
    };
    auto acquire = [&](int s) {
      if (s >= PARING) mbar_wait(BAR(PB_AEMPTY + (uint32_t)s % PARING), (((uint32_t)s / PARING) - 1) & 1, 1);
      return sA + ((uint32_t)s % PARING) * FA_SLOT;
    };
    if (MODE == TC_CELL) {
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
        unsigned char* dst = acquire(s);
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
      }
    } else {
      const int S = n_local * kt1;
      int nidx[RPT];
      if (S > 0) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) nidx[i] = __ldg(p.idx + min((int)blockIdx.x * TM + hw + NHW * i, last));
      }
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        const uint32_t slot = (uint32_t)s % PARING;
        const int jl = s / kt1, t = s - jl * kt1;
        const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
        if (s >= PARING) mbar_wait(BAR(PB_AEMPTY + slot), (((uint32_t)s / PARING) - 1) & 1, 1);
        unsigned char* dst = sA + slot * FA_SLOT;
        if (t >= 4) {
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
          if (t & 1) {
            const int r0 = (t == 3) ? row0 + (int)gridDim.x * TM : row0;
            const int32_t* ip = p.idx + (t == 3 ? 0 : (size_t)p.rows);
#pragma unroll
            for (int i = 0; i < RPT; ++i) nidx[i] = __ldg(ip + min(r0 + hw + NHW * i, last));
          }
        }
        cp_async_commit();
        if (s >= 1) { cp_async_wait<1>(); publish(s - 1); }
      }
      if (S > 0) { cp_async_wait<0>(); publish(S - 1); }
    }
  } else if (warp == WARP_LOAD) {
    // ================================================================= weight loader: THIS CTA's half (rows 64 r .. 64 r + 63) of every K-tile
    if (lane == 0 && n_local > 0) {
      uint32_t wl = 0;
      auto put = [&](int step) {
        const uint32_t slot = wl % PWRING;
        if (wl >= PWRING) mbar_wait(BAR(PB_WEMPTY + slot), ((wl / PWRING) - 1) & 1, 2);
        mbar_arrive_expect_tx(BAR(PB_WFULL + slot), PW_SLOT);
        const unsigned char* src = p.packed + (size_t)step * FW_SLOT + (size_t)rank * PW_HALF;
        const uint32_t dst = smem_u32(sW + slot * PW_SLOT);
        bulk_g2s(dst, src, PW_HALF, BAR(PB_WFULL + slot));                          // hi plane
        bulk_g2s(dst + PW_HALF, src + TILE_BYTES, PW_HALF, BAR(PB_WFULL + slot));   // lo plane
        ++wl;
      };
      auto put_l1 = [&]() { for (int t = 0; t < kt1; ++t) put(t); };
      put_l1();
      if (n_local > 1) put_l1();
#pragma unroll 1
      for (int k = 0; k < n_local; k += 2) {  // the issuer's order
        const bool hasB = k + 1 < n_local;
        put(kt1); put(kt1 + 1);
        if (hasB) { put(kt1); put(kt1 + 1); }
        put(kt1 + 2); put(kt1 + 3);
        if (hasB) { put(kt1 + 2); put(kt1 + 3); }
        if (k + 2 < n_local) put_l1();
        if (k + 3 < n_local) put_l1();
      }
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ================================================================= rank 0: issuer of the pair's MMAs; rank 1: relay of its CTA's readiness
    if (n_local > 0) {  // all 32 lanes walk the sequence (uniform control flow), one elected lane issues / arrives
      uint32_t ak = 0, wk = 0, pk = 0, hphase[2] = {0, 0};
      // one step = one weight K-tile slot (+ an A slot in layer 1, + the hidden activations at the start of layers 2 / 3)
      auto step_ready = [&](bool layer1, int a, bool first23, uint32_t& sa, uint32_t& sw) {
        if (first23) { mbar_wait(BAR(PB_HREADY + a), hphase[a], 7); hphase[a] ^= 1; }
        if (layer1) { sa = ak % PARING; mbar_wait(BAR(PB_AFULL + sa), (ak / PARING) & 1, 4); }
        sw = wk % PWRING;
        mbar_wait(BAR(PB_WFULL + sw), (wk / PWRING) & 1, 5);
        if (rank == 0) {
          mbar_wait_cluster(BAR(PB_PEER + pk % PEER_DEPTH), (pk / PEER_DEPTH) & 1, 10);  // the peer CTA is ready for this step too
        } else {
          tc_fence_after();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(BAR(PB_PEER + pk % PEER_DEPTH), 0);
        }
        ++pk;
      };
      auto layer1 = [&](int jl) {
        const int a = jl & 1;
        const uint32_t d = tmem_base + fwd_acc_col(jl);
#pragma unroll 1
        for (int t = 0; t < kt1; ++t, ++ak, ++wk) {
          uint32_t sa = 0, sw = 0;
          step_ready(true, a, false, sa, sw);
          if (rank != 0) continue;
          tc_fence_after();
          const uint32_t alo = umma_desc_lo(smem_u32(sA + sa * FA_SLOT)), blo = umma_desc_lo(smem_u32(sW + sw * PW_SLOT));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < p.nprod)
                umma2_f16_lo_w(d, alo + (uint32_t)((fprod_a(q) * TILE_BYTES + kk * 32) >> 4), blo + (uint32_t)((fprod_w(q) * PW_HALF + kk * 32) >> 4),
                               IDESC_F16_M256_N128, (t | kk | q) != 0);
          umma_commit2(BAR(PB_AEMPTY + sa));
          umma_commit2(BAR(PB_WEMPTY + sw));
        }
        if (rank == 0) umma_commit2(BAR(PB_DFULL + a));
      };
      auto layer23 = [&](int jl) {
        const int a = jl & 1;
        const uint32_t d = tmem_base + fwd_acc_col(jl), hcol = tmem_base + fwd_h_col(jl);
#pragma unroll 1
        for (int t = 0; t < 2; ++t, ++wk) {
          uint32_t sa = 0, sw = 0;
          step_ready(false, a, t == 0, sa, sw);
          if (rank != 0) continue;
          tc_fence_after();
          const uint32_t blo = umma_desc_lo(smem_u32(sW + sw * PW_SLOT));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < p.nprod)
                umma2_f16_ts_lo_w(d, hcol + 64u * fprod_a(q) + 8u * (uint32_t)(4 * t + kk), blo + (uint32_t)((fprod_w(q) * PW_HALF + kk * 32) >> 4),
                                  IDESC_F16_M256_N128, (t | kk | q) != 0);
          umma_commit2(BAR(PB_WEMPTY + sw));
        }
        if (rank == 0) umma_commit2(BAR(PB_DFULL23 + a));
      };
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
    }
    __syncwarp();
  } else {
    // ================================================================= epilogue warps 0-7 (this CTA's 128 rows), as mlp_tc3.cu's SH variants
    const int q = warp & 3, h = warp >> 2;
    const int erow = 32 * q + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    unsigned char* stage = smem + L.stage_off + warp * 2048;
    float* sRed = reinterpret_cast<float*>(smem + L.red_off);
    const float* b3 = sPar + 2 * 128;
    const float* gw = sPar + 3 * 128;
    const float* gb = sPar + 4 * 128;
    auto hidden = [&](int jl, int layer) {
      const int acc = jl & 1;
      const uint32_t t_acc = t_lane + fwd_acc_col(jl);
      const uint32_t t_ha = t_lane + fwd_h_col(jl);
      uint32_t r[16];
      if (layer == 0) mbar_wait(BAR(PB_DFULL + acc), (uint32_t)(jl >> 1) & 1, 8);
      else mbar_wait(BAR(PB_DFULL23 + acc), 0, 8);
      tc_fence_after();
      const float* bias = sPar + layer * 128;
      const float inv = sPar[5 * 128 + layer];
#pragma unroll 1
      for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
        tmem_ld16_issue(t_acc + 16 * sc, r);
        tmem_ld16_wait(r);
        uint32_t p0[8], p1[8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = *reinterpret_cast<const float4*>(bias + 16 * sc + 4 * j4);
          const float z0 = fmaf(__uint_as_float(r[4 * j4]), inv, b.x), z1 = fmaf(__uint_as_float(r[4 * j4 + 1]), inv, b.y);
          const float z2 = fmaf(__uint_as_float(r[4 * j4 + 2]), inv, b.z), z3 = fmaf(__uint_as_float(r[4 * j4 + 3]), inv, b.w);
          split2_pair(silu_x3_scaled(z0), silu_x3_scaled(z1), p0[2 * j4], p1[2 * j4]);
          split2_pair(silu_x3_scaled(z2), silu_x3_scaled(z3), p0[2 * j4 + 1], p1[2 * j4 + 1]);
        }
        tmem_st8(t_ha + 8 * sc, p0);
        tmem_st8(t_ha + 64 + 8 * sc, p1);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(PB_HREADY + acc));
    };
    auto final_ep = [&](int jl) {
      const int row0 = ((int)blockIdx.x + jl * (int)gridDim.x) * TM;
      const int acc = jl & 1;
      const uint32_t t_acc = t_lane + fwd_acc_col(jl);
      uint32_t r[16];
      mbar_wait(BAR(PB_DFULL23 + acc), 1, 9);
      tc_fence_after();
      const float inv3 = sPar[5 * 128 + 2];
      // one-read LayerNorm statistics: the arithmetic of mlp_tc3.cu's final epilogue, operation for operation (bit-identical results)
      float s1 = 0.f, s2 = 0.f, c = 0.f;
#pragma unroll 1
      for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
        tmem_ld16_issue(t_acc + 16 * sc, r);
        tmem_ld16_wait(r);
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
      }
      const float mh = c + s1 * (1.0f / 64.0f);
      const float m2 = s2 - s1 * s1 * (1.0f / 64.0f);
      float* red = sRed + 512 * acc;
      red[h * 128 + erow] = mh;
      red[256 + h * 128 + erow] = m2;
      named_bar_sync(1, NEPI);
      const float m0 = red[erow], m1 = red[128 + erow];
      const float dm = m0 - m1;
      const float mean = 0.5f * (m0 + m1);
      const float var = (red[256 + erow] + red[256 + 128 + erow] + dm * dm * 32.0f) * (1.0f / 128.0f);
      const float rstd = 1.0f / sqrtf((var < 0.0f ? 0.0f : var) + 1e-5f);
      if (h == 0 && !(rstd * 0.0f == 0.0f) && row0 + erow < p.rows) atomicOr(&g_status, 1u);
#pragma unroll 1
      for (int sc = 4 * h; sc < 4 * h + 4; ++sc) {
        tmem_ld16_issue(t_acc + 16 * sc, r);
        tmem_ld16_wait(r);
        float* v = reinterpret_cast<float*>(r);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = *reinterpret_cast<const float4*>(b3 + 16 * sc + 4 * j4);
          const float4 w = *reinterpret_cast<const float4*>(gw + 16 * sc + 4 * j4);
          const float4 o = *reinterpret_cast<const float4*>(gb + 16 * sc + 4 * j4);
          const float y0 = fmaf(v[4 * j4], inv3, b.x), y1 = fmaf(v[4 * j4 + 1], inv3, b.y);
          const float y2 = fmaf(v[4 * j4 + 2], inv3, b.z), y3 = fmaf(v[4 * j4 + 3], inv3, b.w);
          v[4 * j4] = (y0 - mean) * rstd * w.x + o.x; v[4 * j4 + 1] = (y1 - mean) * rstd * w.y + o.y;
          v[4 * j4 + 2] = (y2 - mean) * rstd * w.z + o.z; v[4 * j4 + 3] = (y3 - mean) * rstd * w.w + o.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float4*>(stage + (uint32_t)lane * 64u + (((uint32_t)k ^ (((uint32_t)lane >> 1) & 3u)) << 4)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        __syncwarp();
        const int rr0 = lane >> 2, cp = lane & 3;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = 8 * it + rr0, grow = row0 + 32 * q + rr;
          const float4 y = *reinterpret_cast<const float4*>(stage + (uint32_t)rr * 64u + ((uint32_t)cp << 4));
          const int col = 16 * sc + 4 * (cp ^ ((rr >> 1) & 3));
          if (grow < p.rows) {
            const size_t off = (size_t)grow * 128 + col;
            if (MODE == TC_CELL) st_shadow_f4(p.sh_prime, (size_t)grow, col, y);
            red_add_f4(p.out1 + off, y);
          }
        }
        __syncwarp();
      }
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
      named_bar_sync(1, NEPI);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA releases its TMEM / leaves while the other may still be signalled or read
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

}  // namespace x3

// mode: TC_CELL / TC_EDGE; returns FVGN_E_UNSUPPORTED when the pair kernels do not apply (the caller then uses the single-CTA kernels)
static int tc3_launch_pair(x3::Params& p, cudaStream_t st) {
  using namespace x3;
  const SmemLayout L = smem_layout_pair();
  const int nsm = sm_count() & ~1;
  const int want = (p.ntiles + 1) & ~1;
  const int grid = want < nsm ? want : nsm;
  if (grid < 2) return FVGN_E_UNSUPPORTED;
  if (p.mode == TC_CELL) {
    static ::fvgn::PerDeviceOnce configured;
    if (configured.first()) FVGN_CUDA(cudaFuncSetAttribute(fused_mlp_tc3_pair_kernel<TC_CELL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    fused_mlp_tc3_pair_kernel<TC_CELL><<<grid, NTHREADS, L.total, st>>>(p);
  } else {
    static ::fvgn::PerDeviceOnce configured;
    if (configured.first()) FVGN_CUDA(cudaFuncSetAttribute(fused_mlp_tc3_pair_kernel<TC_EDGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    fused_mlp_tc3_pair_kernel<TC_EDGE><<<grid, NTHREADS, L.total, st>>>(p);
  }
  FVGN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fvgn
