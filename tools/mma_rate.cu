// mma_rate.cu — what one tcgen05.mma of the forward kernel's shape costs when nothing else runs on the SM.
// One CTA per SM, one warp issues R back-to-back kind::f16 MMAs (M = 128, N = 128 or 256, K = 16) on operands that sit in shared memory
// (SS mode) or with A in TMEM (TS mode), into ONE accumulator or alternating between two, then commits and waits; clock64 around the lot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I finite-volume-graph-network_b200/csrc -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace fvgn;

constexpr uint32_t IDESC_F16_N128 = 0x10u | (16u << 17) | (8u << 24);           // kind::f16, fp16 A/B, fp32 D, M = 128, N = 128
constexpr uint32_t IDESC_F16_N256 = 0x10u | (32u << 17) | (8u << 24);           // N = 256

__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) { umma_f16_lo_w(d, alo, blo, idesc, acc); }

// MODE 5 / 6: as 0 / 2 while `ld_warps` other warps keep reading (5) or reading and writing (6) OTHER TMEM columns with tcgen05.ld / .st —
// what the epilogue warps of the fused kernels do next to the MMAs
// 32 lanes x 64 consecutive 32-bit TMEM columns in one instruction
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, "
      "%51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\ttcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
        "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
        "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

template <int MODE>  // 0: SS one accumulator, 1: SS two accumulators alternating, 2: TS one accumulator, 3: SS N = 256, 4: SS distinct A / B slices per MMA
__global__ void __launch_bounds__(288, 1) rate_kernel(int reps, long long* out, int ld_warps, int ts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (warp == 0) {
    const uint32_t a0 = umma_desc_lo(smem_u32(smem)), b0 = umma_desc_lo(smem_u32(smem + 2 * 16384));
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t kk = (uint32_t)(r & 3) * 2;  // 32-byte K slices of one K-tile
      if (MODE == 0) mma_ss(tm, a0 + kk, b0 + kk, IDESC_F16_N128, r != 0);
      else if (MODE == 1) mma_ss(tm + 128u * (uint32_t)(r & 1), a0 + kk, b0 + kk, IDESC_F16_N128, r > 1);
      else if (MODE == 2) umma_f16_ts_lo_w(tm, tm + 256u + 8u * (uint32_t)(r & 7), b0 + kk, IDESC_F16_N128, r != 0);
      else if (MODE == 3) mma_ss(tm, a0 + kk, b0 + kk, IDESC_F16_N256, r != 0);
      else if (MODE == 5 || MODE == 6 || MODE == 7) {
        if (ts) umma_f16_ts_lo_w(tm, tm + 192u + 8u * (uint32_t)(r & 7), b0 + kk, IDESC_F16_N128, r != 0);
        else mma_ss(tm, a0 + kk, b0 + kk, IDESC_F16_N128, r != 0);
      }
      else mma_ss(tm, a0 + kk + (uint32_t)((r >> 2) & 1) * 1024u, b0 + kk + (uint32_t)((r >> 3) & 1) * 1024u, IDESC_F16_N128, r != 0);
    }
    long long t1 = clock64();
    umma_commit_w(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    if (threadIdx.x == 0) done = 1;
  } else if (MODE == 7 && warp <= ld_warps) {
    // the same TMEM bytes per loop as MODE 5, in one 64-column load instead of four 16-column loads
    const uint32_t t_lane = tm + ((uint32_t)(32 * (warp & 3)) << 16) + 256u + 64u * (uint32_t)((warp - 1) >> 2);
    uint32_t r[64];
    unsigned int sink = 0, n = 0;
    while (!done) {
      tmem_ld64(t_lane, r);
      sink += r[0] + r[63];
      ++n;
    }
    if (sink == 0x12345678u) out[2] = n;
    if (threadIdx.x == 32 && blockIdx.x == 0) out[3] = n;
  } else if ((MODE == 5 || MODE == 6) && warp <= ld_warps) {
    // "epilogue" warps: this warp's 32 TMEM lanes, columns 256.. (never touched by the MMAs above)
    const uint32_t t_lane = tm + ((uint32_t)(32 * (warp & 3)) << 16) + 256u + 64u * (uint32_t)((warp - 1) >> 2);
    uint32_t r[16];
    unsigned int sink = 0, n = 0;
    while (!done) {
#pragma unroll
      for (int sc = 0; sc < 4; ++sc) {
        tmem_ld16_issue(t_lane + 16 * sc, r);
        tmem_ld16_wait(r);
        sink += r[0] + r[15];
        if (MODE == 6) { tmem_st8(t_lane + 8 * sc, r); tmem_st8(t_lane + 32 + 8 * sc, r + 8); }
      }
      if (MODE == 6) tmem_wait_st();
      ++n;
    }
    if (sink == 0x12345678u) out[2] = n;
    if (threadIdx.x == 32 && blockIdx.x == 0) out[3] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE>
static void run(const char* name, int reps, int ld_warps = 0, int ts = 0) {
  long long* out;
  cudaMalloc(&out, 32);
  cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16384);
  for (int it = 0; it < 2; ++it) rate_kernel<MODE><<<148, 288, 6 * 16384>>>(reps, out, ld_warps, ts);
  long long h[4];
  cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-58s %5d MMAs: issue %7.1f clk/MMA, issue -> all complete %7.1f clk/MMA", name, reps, (double)h[0] / reps, (double)h[1] / reps);
  if (ld_warps) printf("   (warp 1: %lld loops of 8 KB = %.1f B/clk per warp)", h[3], (double)h[3] * 8192.0 / (double)h[1]);
  printf("   %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int reps : {24, 72, 720}) {
    run<0>("SS  M128 N128 K16, one accumulator", reps);
    run<1>("SS  M128 N128 K16, two accumulators alternating", reps);
    run<2>("TS  M128 N128 K16 (A in TMEM), one accumulator", reps);
    run<3>("SS  M128 N256 K16, one accumulator", reps);
    run<4>("SS  M128 N128 K16, A / B slices vary per MMA", reps);
  }
  for (int w : {4, 8}) {
    char nm[96];
    snprintf(nm, sizeof nm, "SS  N128 + %d warps in a tcgen05.ld loop", w); run<5>(nm, 720, w, 0);
    snprintf(nm, sizeof nm, "TS  N128 + %d warps in a tcgen05.ld loop", w); run<5>(nm, 720, w, 1);
    snprintf(nm, sizeof nm, "SS  N128 + %d warps in a tcgen05.ld.x64 loop", w); run<7>(nm, 720, w, 0);
    snprintf(nm, sizeof nm, "TS  N128 + %d warps in a tcgen05.ld.x64 loop", w); run<7>(nm, 720, w, 1);
    snprintf(nm, sizeof nm, "SS  N128 + %d warps in a tcgen05.ld + st loop", w); run<6>(nm, 720, w, 0);
    snprintf(nm, sizeof nm, "TS  N128 + %d warps in a tcgen05.ld + st loop", w); run<6>(nm, 720, w, 1);
  }
  return 0;
}
