// mlp_tc.cu — tcgen05 / TMEM tensor-core path of the fused gather -> MLP -> LayerNorm -> residual kernels (sm_100a).
//
// One persistent CTA per SM, 128-row tiles, fully decoupled warp-specialised pipeline (14 warps):
//   warps 0-3   producers : gather the tile's operand rows from HBM (16 x 128-bit loads in flight per thread), convert to
//                           bf16 and write 64-column K-tiles into a 5-slot shared-memory ring in the canonical K-major
//                           SWIZZLE_128B UMMA layout.  They run up to one tile + 5 K-tiles ahead of the epilogues.
//   warps 4-11  epilogue  : tcgen05.ld the fp32 accumulator from TMEM (one row per thread), bias + SiLU -> bf16 A operand
//                           of the next layer (H tile), or bias + LayerNorm + residual -> HBM through a swizzled smem
//                           staging tile so that every global access is a full 128-byte line
//   warp 12     MMA       : one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM, M=128 N=128 K=16).
//                           Two TMEM accumulators ping-pong: layer 1 of tile i+1 runs while tile i is in its epilogues.
//   warp 13     loader    : streams the first-layer weight K-tiles through a 2-slot ring with cp.async.bulk (TMA bulk copy)
//                           on mbarriers; W2/W3 are loaded once and stay resident
// Weights are pre-packed once per weight version (fvgn_mlp_pack_bf16) into bf16 tiles that are byte images of the
// swizzled shared-memory layout, so the loader needs no tensor map: each K-tile is one 16 KB bulk copy.
// Every mbarrier wait has a watchdog that traps instead of hanging the device.
#include <cuda_bf16.h>
#include "common.cuh"

namespace fvgn {

constexpr int TM = 128;            // rows per tile (= UMMA M)
constexpr int HIDN = 128;          // N of every layer (padded)
constexpr int KTILE = 64;          // bf16 elements per 128-byte swizzle row
constexpr int TILE_BYTES = 16384;  // [128 rows][128 B]
constexpr int ARING = 5;           // A-operand K-tile ring slots
constexpr int WRING = 2;           // W1 K-tile ring slots
constexpr int NPROD = 256;         // producer threads (warps 0-7)
constexpr int NPW = NPROD / 32;    // producer warps
constexpr int NEPI = 256;          // epilogue threads (warps 8-15)
constexpr int WARP_EPI0 = NPW, WARP_MMA = NPW + 8, WARP_LOAD = NPW + 9;
constexpr int NTHREADS = (NPW + 10) * 32;
constexpr int RPT = TM / (NPROD / 16);  // rows per producer thread per K-tile (8)

enum { TC_ROWS = 0, TC_CELL = 1, TC_EDGE = 2 };

struct TcParams {
  int mode;
  int rows;
  int ntiles;
  int kt1;    // K-tiles of layer 1
  int k_in;   // ROWS: valid input width
  int n_out;  // valid outputs of layer 3
  const float* in; int ld_in;                               // ROWS
  const float* src0; const float* src1; const int32_t* idx;  // CELL: x, node_agg, cells_node_T ; EDGE: x', e, neighbour_cell
  float* out0; int ld_out;
  float* out1;
  const unsigned char* packed;  // [kt1 + 4] tiles: W1 tiles, W2 (2), W3 (2)
  const float* b1; const float* b2; const float* b3; const float* ln_w; const float* ln_b;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int who) {
  printf("fvgn mlp_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u site %d)\n", (int)blockIdx.x, (int)threadIdx.x, bar,
         parity, who);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity, who);  // ~2 s: a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32 | version 1<<46 | layout 2<<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major A and B, N=128 (16<<17), M=128 (8<<24)
constexpr uint32_t IDESC_BF16_M128_N128 = 0x10u | 0x80u | 0x400u | (16u << 17) | (8u << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane (lane_base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
      "%27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// x * sigmoid(x) = 0.5 x (1 + tanh(x/2)): one MUFU per element (tanh.approx rel. error ~2^-11, below the bf16 rounding of H)
__device__ __forceinline__ float silu_fast(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
  const float hv = 0.5f * v;
  return fmaf(hv, t, hv);
}

// store 4 consecutive fp32 (columns 4*hl .. 4*hl+3 of a 64-wide K-tile) as bf16 into row r of a swizzled tile
__device__ __forceinline__ void st_tile_f4(unsigned char* tile, int r, int hl, float4 v) {
  const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)hl >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)hl & 1u) * 8u;
  uint2 pk;
  pk.x = pack_bf16(v.x, v.y);
  pk.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + off) = pk;
}

// ------------------------------------------------------------------------------------------------ shared memory map
struct SmemLayout {
  uint32_t a_off, h_off, w2_off, w3_off, wring_off, par_off, red_off, bar_off, total;
};
__host__ __device__ inline SmemLayout smem_layout() {
  SmemLayout L;
  uint32_t o = 0;
  L.a_off = o; o += ARING * TILE_BYTES;      // 80 KB  A-operand K-tile ring
  L.h_off = o; o += 2 * TILE_BYTES;          // 32 KB  hidden activations (A operand of layers 2,3) / fp32 output staging
  L.w2_off = o; o += 2 * TILE_BYTES;         // 32 KB
  L.w3_off = o; o += 2 * TILE_BYTES;         // 32 KB
  L.wring_off = o; o += WRING * TILE_BYTES;  // 32 KB  W1 K-tile ring
  L.par_off = o; o += 5 * 128 * 4;           // b1 | b2 | b3 | ln_w | ln_b
  L.red_off = o; o += 2 * 2 * 128 * 4;       // LayerNorm cross-warp exchange
  L.bar_off = o; o += 256;
  L.total = o + 1024;                        // slack for the 1024-byte alignment of the dynamic smem base
  return L;
}

// barrier indices (8 bytes each)
enum { B_AFULL = 0, B_AEMPTY = B_AFULL + ARING, B_WFULL = B_AEMPTY + ARING, B_WEMPTY = B_WFULL + WRING, B_WRES = B_WEMPTY + WRING,
       B_DFULL = B_WRES + 1, B_DFREE = B_DFULL + 2, B_HREADY = B_DFREE + 2, B_COUNT = B_HREADY + 1 };

__global__ void __launch_bounds__(NTHREADS, 1) fused_mlp_tc_kernel(const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const SmemLayout L = smem_layout();
  unsigned char* sA = smem + L.a_off;
  unsigned char* sH = smem + L.h_off;
  unsigned char* sW2 = smem + L.w2_off;
  unsigned char* sW3 = smem + L.w3_off;
  unsigned char* sWr = smem + L.wring_off;
  float* sPar = reinterpret_cast<float*>(smem + L.par_off);
  float* sRed = reinterpret_cast<float*>(smem + L.red_off);
  const uint32_t bar0 = smem_u32(smem + L.bar_off);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt1 = p.kt1;

  if (tid == 0) {
    for (int i = 0; i < ARING; ++i) { mbar_init(BAR(B_AFULL + i), NPW); mbar_init(BAR(B_AEMPTY + i), 1); }
    for (int i = 0; i < WRING; ++i) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
    mbar_init(BAR(B_WRES), 1);
    for (int b = 0; b < 2; ++b) { mbar_init(BAR(B_DFULL + b), 1); mbar_init(BAR(B_DFREE + b), 8); }
    mbar_init(BAR(B_HREADY), 8);
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 256);
  if (warp >= WARP_EPI0 && warp < WARP_EPI0 + 8) {
    for (int i = tid - NPROD; i < 5 * 128; i += NEPI) {
      const int which = i >> 7, c = i & 127;
      float v = 0.f;
      if (which == 0) v = p.b1[c];
      else if (which == 1) v = p.b2[c];
      else if (which == 2) v = (c < p.n_out) ? p.b3[c] : 0.f;
      else if (which == 3) v = p.ln_w ? p.ln_w[c] : 1.f;
      else v = p.ln_b ? p.ln_b[c] : 0.f;
      sPar[i] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NPW) {
    // ================================================================= producers: gather -> A ring
    constexpr int NHW = NPROD / 16;          // half-warps: half-warp hw owns rows hw, hw+NHW, ...
    const int hw = tid >> 4, hl = tid & 15;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int row0 = tile * TM;
      for (int t = 0; t < kt1; ++t, ++g) {
        const uint32_t slot = g % ARING;
        // issue the loads of this K-tile before waiting for the slot: the wait overlaps the memory latency
        float4 v[RPT];
        if (p.mode == TC_EDGE) {
          const int F = p.rows;
          int src[RPT];
          if (t < 4) {
            const int32_t* ip = p.idx + (t >= 2 ? (size_t)F : 0);
#pragma unroll
            for (int i = 0; i < RPT; ++i) { const int f = row0 + hw + NHW * i; src[i] = (f < F) ? __ldg(ip + f) : -1; }
            const float* base = p.src0 + (t & 1) * 64 + hl * 4;
#pragma unroll
            for (int i = 0; i < RPT; ++i) v[i] = (src[i] >= 0) ? ldg4(base + (size_t)src[i] * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            const float* base = p.src1 + (t & 1) * 64 + hl * 4;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const int f = row0 + hw + NHW * i;
              v[i] = (f < F) ? *reinterpret_cast<const float4*>(base + (size_t)f * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        } else if (p.mode == TC_CELL) {
          const int C = p.rows;
          if (t < 2) {
            const float* base = p.src0 + t * 64 + hl * 4;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const int c = row0 + hw + NHW * i;
              v[i] = (c < C) ? *reinterpret_cast<const float4*>(base + (size_t)c * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const int c = row0 + hw + NHW * i;
              float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
              if (c < C) {
                const int n0 = __ldg(p.idx + c), n1 = __ldg(p.idx + (size_t)C + c), n2 = __ldg(p.idx + 2 * (size_t)C + c);
                const float4 a = ldg4(p.src1 + (size_t)n0 * 64 + hl * 4);
                const float4 b = ldg4(p.src1 + (size_t)n1 * 64 + hl * 4);
                const float4 d = ldg4(p.src1 + (size_t)n2 * 64 + hl * 4);
                o.x = ((a.x + b.x) + d.x) / 3.0f; o.y = ((a.y + b.y) + d.y) / 3.0f;
                o.z = ((a.z + b.z) + d.z) / 3.0f; o.w = ((a.w + b.w) + d.w) / 3.0f;
              }
              v[i] = o;
            }
          }
        } else {  // TC_ROWS: generic [rows, k_in] fp32 input, zero padded to 64*kt1
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
        if (g >= ARING) mbar_wait(BAR(B_AEMPTY + slot), ((g / ARING) - 1) & 1, 1);
        unsigned char* dst = sA + slot * TILE_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) st_tile_f4(dst, hw + NHW * i, hl, v[i]);
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_AFULL + slot));
      }
    }
  } else if (warp == WARP_LOAD) {
    // ================================================================= loader: W2/W3 once, W1 K-tiles through the ring
    if (lane == 0) {
      mbar_arrive_expect_tx(BAR(B_WRES), 4 * TILE_BYTES);
      bulk_g2s(smem_u32(sW2), p.packed + (size_t)kt1 * TILE_BYTES, 2 * TILE_BYTES, BAR(B_WRES));
      bulk_g2s(smem_u32(sW3), p.packed + (size_t)(kt1 + 2) * TILE_BYTES, 2 * TILE_BYTES, BAR(B_WRES));
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int t = 0; t < kt1; ++t, ++g) {
          const uint32_t slot = g % WRING;
          if (g >= WRING) mbar_wait(BAR(B_WEMPTY + slot), ((g / WRING) - 1) & 1, 2);
          mbar_arrive_expect_tx(BAR(B_WFULL + slot), TILE_BYTES);
          bulk_g2s(smem_u32(sWr + slot * TILE_BYTES), p.packed + (size_t)t * TILE_BYTES, TILE_BYTES, BAR(B_WFULL + slot));
        }
      }
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ================================================================= MMA issuer (single thread)
    if (lane == 0) {
      uint32_t g = 0;           // K-tile counter shared by the A ring and the W1 ring
      uint32_t hphase = 0;      // h_ready parity
      uint32_t dfree_uses[2] = {0, 0};
      // layer 1 of local tile j into accumulator j&1
      auto issue_layer1 = [&](int j) {
        const int b = j & 1;
        if (j >= 2) { mbar_wait(BAR(B_DFREE + b), dfree_uses[b] & 1, 3); ++dfree_uses[b]; }
        tc_fence_after();
        const uint32_t d = tmem_base + 128u * (uint32_t)b;
        for (int t = 0; t < kt1; ++t, ++g) {
          const uint32_t sa = g % ARING, sw = g % WRING;
          mbar_wait(BAR(B_AFULL + sa), (g / ARING) & 1, 4);
          mbar_wait(BAR(B_WFULL + sw), (g / WRING) & 1, 5);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + sa * TILE_BYTES), b_base = smem_u32(sWr + sw * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(d, umma_desc(a_base + kk * 32), umma_desc(b_base + kk * 32), IDESC_BF16_M128_N128, (t | kk) != 0);
          umma_commit(BAR(B_AEMPTY + sa));  // both ring slots are free once these MMAs have read them
          umma_commit(BAR(B_WEMPTY + sw));
        }
        umma_commit(BAR(B_DFULL + b));
      };
      int j = 0;
      const int first = blockIdx.x;
      if (first < p.ntiles) issue_layer1(0);
      mbar_wait(BAR(B_WRES), 0, 6);
      for (int tile = first; tile < p.ntiles; tile += gridDim.x, ++j) {
        const int b = j & 1;
        if (tile + (int)gridDim.x < p.ntiles) issue_layer1(j + 1);  // runs under tile j's epilogues
        const uint32_t d = tmem_base + 128u * (uint32_t)b;
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(BAR(B_HREADY), hphase, 7); hphase ^= 1;
          tc_fence_after();
          const unsigned char* w = layer == 0 ? sW2 : sW3;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t a_base = smem_u32(sH + t * TILE_BYTES), b_base = smem_u32(w + t * TILE_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(d, umma_desc(a_base + kk * 32), umma_desc(b_base + kk * 32), IDESC_BF16_M128_N128, (t | kk) != 0);
          }
          umma_commit(BAR(B_DFULL + b));
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================= epilogue warps 4..11
    const int ew = warp - WARP_EPI0;
    const int q = warp & 3, h = ew >> 2;  // TMEM lane quadrant (= warp id % 4), column half
    const int erow = 32 * q + lane;       // this thread's row inside the tile
    const int et = tid - NPROD;           // 0..255
    const int ht = et & 127;              // thread index inside its column-half team
    uint32_t dfull_uses[2] = {0, 0};
    int j = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++j) {
      const int row0 = tile * TM;
      const int b = j & 1;
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * q) << 16) + 128u * (uint32_t)b + (uint32_t)(64 * h);
      // ---------------------------------------------------------------- epilogues 1, 2: bias + SiLU -> bf16 H tiles
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(BAR(B_DFULL + b), dfull_uses[b] & 1, 8); ++dfull_uses[b];
        tc_fence_after();
        const float* bias = sPar + layer * 128 + 64 * h;
        unsigned char* hrow = sH + h * TILE_BYTES + (uint32_t)erow * 128u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tmem_ld32(t_addr + 32 * half, v);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {  // 8 columns = one 16-byte chunk
            float o[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) o[jj] = silu_fast(v[c4 * 8 + jj] + bias[half * 32 + c4 * 8 + jj]);
            uint4 pk;
            pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]); pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
            const uint32_t chunk = (uint32_t)(half * 4 + c4);
            *reinterpret_cast<uint4*>(hrow + ((chunk ^ ((uint32_t)erow & 7u)) << 4)) = pk;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HREADY));
      }
      // ---------------------------------------------------------------- epilogue 3: bias (+ LayerNorm) (+ residual) -> HBM
      mbar_wait(BAR(B_DFULL + b), dfull_uses[b] & 1, 9); ++dfull_uses[b];
      tc_fence_after();
      const float* b3 = sPar + 2 * 128 + 64 * h;
      float mean = 0.f, rstd = 1.f;
      const bool has_ln = (p.ln_w != nullptr);
      if (has_ln) {
        float s = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tmem_ld32(t_addr + 32 * half, v);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) s += v[jj] + b3[half * 32 + jj];
        }
        sRed[h * 128 + erow] = s;
        named_bar_sync(1, NEPI);
        mean = (sRed[erow] + sRed[128 + erow]) * (1.0f / 128.0f);
        float qv = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tmem_ld32(t_addr + 32 * half, v);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) { const float d = v[jj] + b3[half * 32 + jj] - mean; qv = fmaf(d, d, qv); }
        }
        sRed[256 + h * 128 + erow] = qv;
        named_bar_sync(1, NEPI);
        rstd = 1.0f / sqrtf((sRed[256 + erow] + sRed[256 + 128 + erow]) * (1.0f / 128.0f) + 1e-5f);
      }
      const bool staged = (p.mode != TC_ROWS) || (p.n_out == HIDN && (p.ld_out & 3) == 0);
      const float* gw = sPar + 3 * 128 + 64 * h;
      const float* gb = sPar + 4 * 128 + 64 * h;
      unsigned char* stage = sH + h * TILE_BYTES;  // [128 rows][32 fp32], 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(t_addr + 32 * half, v);
        if (half == 1) {  // last TMEM read of this tile: the accumulator may be overwritten by layer 1 of tile j+2
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_DFREE + b));
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          float y = v[jj] + b3[half * 32 + jj];
          if (has_ln) y = (y - mean) * rstd * gw[half * 32 + jj] + gb[half * 32 + jj];
          v[jj] = y;
        }
        if (staged) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stage + (uint32_t)erow * 128u + (((uint32_t)c ^ ((uint32_t)erow & 7u)) << 4)) =
                make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          named_bar_sync(2 + h, 128);
          // coalesced read-out: 8 threads per row -> every warp access covers 4 full 128-byte lines
          const int col0 = 64 * h + 32 * half;
#pragma unroll 4
          for (int pass = 0; pass < 8; ++pass) {
            const int r = pass * 16 + (ht >> 3), cp = ht & 7;
            const int grow = row0 + r;
            if (grow < p.rows) {
              const float4 y = *reinterpret_cast<const float4*>(stage + (uint32_t)r * 128u + ((uint32_t)cp << 4));
              const int col = col0 + 4 * (cp ^ (r & 7));
              if (p.mode == TC_ROWS) {
                *reinterpret_cast<float4*>(p.out0 + (size_t)grow * p.ld_out + col) = y;
              } else {
                if (p.out0) *reinterpret_cast<float4*>(p.out0 + (size_t)grow * 128 + col) = y;
                if (p.out1) {
                  const float4 rv = *reinterpret_cast<const float4*>((p.mode == TC_CELL ? p.src0 : p.src1) + (size_t)grow * 128 + col);
                  *reinterpret_cast<float4*>(p.out1 + (size_t)grow * 128 + col) = make_float4(rv.x + y.x, rv.y + y.y, rv.z + y.z, rv.w + y.w);
                }
              }
            }
          }
          named_bar_sync(2 + h, 128);  // staging tile is rewritten by the next half / the next tile's H
        } else {
          const int grow = row0 + erow;
          if (grow < p.rows) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int col = 64 * h + 32 * half + jj;
              if (col < p.n_out) p.out0[(size_t)grow * p.ld_out + col] = v[jj];
            }
          }
        }
      }
    }
  }
  // ---------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// W [n_valid, k_valid] fp32 row-major -> tiles [kt][128 rows][64 k] bf16, each tile a byte image of the swizzled smem layout
__global__ void pack_tiles_kernel(const float* __restrict__ W, int ldw, int n_valid, int k_valid, int kt, unsigned char* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tile, n, chunk)
  if (idx >= kt * 128 * 8) return;
  const int c = idx & 7, n = (idx >> 3) & 127, t = idx >> 10;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = t * KTILE + c * 8 + j;
    v[j] = (n < n_valid && k < k_valid) ? W[(size_t)n * ldw + k] : 0.f;
  }
  uint4 pk;
  pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]); pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(out + (size_t)t * TILE_BYTES + (size_t)n * 128 + (((uint32_t)c ^ ((uint32_t)n & 7u)) << 4)) = pk;
}

static int kt_of(int k_in) { return (k_in + KTILE - 1) / KTILE; }

size_t tc_packed_bytes(int k_in) { return (size_t)(kt_of(k_in) + 4) * TILE_BYTES; }

int tc_pack(const fvgn_mlp_t* m, void* packed, size_t bytes, cudaStream_t st) {
  FVGN_CHECK_ARG(m && packed, "NULL pointer");
  FVGN_CHECK_ARG(m->k_in >= 1 && m->k_in <= 384 && m->n_out >= 1 && m->n_out <= 128, "bad mlp dims");
  FVGN_CHECK_ARG(bytes >= tc_packed_bytes(m->k_in), "packed buffer too small");
  FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed buffer must be 16B aligned");
  const int kt1 = kt_of(m->k_in);
  unsigned char* o = reinterpret_cast<unsigned char*>(packed);
  pack_tiles_kernel<<<cdiv(kt1 * 1024, 256), 256, 0, st>>>(m->w1, m->k_in, 128, m->k_in, kt1, o);
  FVGN_LAUNCH_CHECK();
  pack_tiles_kernel<<<cdiv(2 * 1024, 256), 256, 0, st>>>(m->w2, 128, 128, 128, 2, o + (size_t)kt1 * TILE_BYTES);
  FVGN_LAUNCH_CHECK();
  pack_tiles_kernel<<<cdiv(2 * 1024, 256), 256, 0, st>>>(m->w3, 128, m->n_out, 128, 2, o + (size_t)(kt1 + 2) * TILE_BYTES);
  FVGN_LAUNCH_CHECK();
  return 0;
}

static int launch_tc(TcParams& p, const fvgn_mlp_t* m, cudaStream_t st) {
  if (p.rows <= 0) return 0;
  FVGN_CHECK_ARG(m->packed != nullptr, "tensor-core modes need packed weights: call fvgn_mlp_pack_bf16 first");
  p.packed = reinterpret_cast<const unsigned char*>(m->packed);
  p.b1 = m->b1; p.b2 = m->b2; p.b3 = m->b3; p.ln_w = m->ln_w; p.ln_b = m->ln_b;
  p.k_in = m->k_in; p.n_out = m->n_out;
  p.kt1 = kt_of(m->k_in);
  p.ntiles = cdiv(p.rows, TM);
  const SmemLayout L = smem_layout();
  static bool configured = false;
  if (!configured) {
    FVGN_CUDA(cudaFuncSetAttribute(fused_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    configured = true;
  }
  const int grid = p.ntiles < sm_count() ? p.ntiles : sm_count();
  fused_mlp_tc_kernel<<<grid, NTHREADS, L.total, st>>>(p);
  FVGN_LAUNCH_CHECK();
  return 0;
}

static int check_tc_mlp(const fvgn_mlp_t* m, bool need_ln, int math_mode) {
  FVGN_CHECK_ARG(m != nullptr, "mlp is NULL");
  FVGN_CHECK_ARG(math_mode == FVGN_MATH_BF16_TC, "only FVGN_MATH_BF16_TC is built into the tcgen05 path so far");
  FVGN_CHECK_ARG(m->w1 && m->b1 && m->w2 && m->b2 && m->w3 && m->b3, "mlp weight pointer is NULL");
  if (m->ln_w || need_ln) FVGN_CHECK_ARG(m->ln_w && m->ln_b && m->n_out == 128, "LayerNorm needs weight/bias and n_out == 128");
  return 0;
}

int tc_mlp_rows(const fvgn_mlp_t* mlp, const float* in, int ld_in, int rows, float* out, int ld_out, int math_mode, cudaStream_t st) {
  if (int e = check_tc_mlp(mlp, false, math_mode)) return e;
  FVGN_CHECK_ARG(mlp->k_in <= 384, "k_in too large");
  FVGN_CHECK_ARG(rows >= 0 && (rows == 0 || (in && out)), "in/out NULL");
  FVGN_CHECK_ARG(ld_in >= mlp->k_in && ld_out >= mlp->n_out, "leading dimension too small");
  if (mlp->n_out == 128 && (ld_out & 3) == 0) FVGN_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "output must be 16B aligned");
  TcParams p{};
  p.mode = TC_ROWS; p.rows = rows; p.in = in; p.ld_in = ld_in; p.out0 = out; p.ld_out = ld_out;
  return launch_tc(p, mlp, st);
}

int tc_cell_block(const fvgn_mlp_t* mlp, const float* x, const float* node_agg, const int32_t* cells_node_T, int n_cells, float* x_prime,
                  float* x_out, int math_mode, cudaStream_t st) {
  if (int e = check_tc_mlp(mlp, true, math_mode)) return e;
  FVGN_CHECK_ARG(mlp->k_in == 192, "CellBlock MLP must have k_in == 192");
  FVGN_CHECK_ARG(n_cells >= 0 && (n_cells == 0 || (x && node_agg && cells_node_T)), "NULL input");
  TcParams p{};
  p.mode = TC_CELL; p.rows = n_cells; p.src0 = x; p.src1 = node_agg; p.idx = cells_node_T; p.out0 = x_prime; p.out1 = x_out; p.ld_out = 128;
  return launch_tc(p, mlp, st);
}

int tc_edge_block(const fvgn_mlp_t* mlp, const float* x_prime, const float* e, const int32_t* neighbour_cell, int n_faces, float* e_prime,
                  float* e_out, int math_mode, cudaStream_t st) {
  if (int er = check_tc_mlp(mlp, true, math_mode)) return er;
  FVGN_CHECK_ARG(mlp->k_in == 384, "EdgeBlock MLP must have k_in == 384");
  FVGN_CHECK_ARG(n_faces >= 0 && (n_faces == 0 || (x_prime && e && neighbour_cell)), "NULL input");
  TcParams p{};
  p.mode = TC_EDGE; p.rows = n_faces; p.src0 = x_prime; p.src1 = e; p.idx = neighbour_cell; p.out0 = e_prime; p.out1 = e_out; p.ld_out = 128;
  return launch_tc(p, mlp, st);
}

}  // namespace fvgn
