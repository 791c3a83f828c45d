// halo.cu — the exchange steps of the single-mesh domain decomposition (host: domain.py) over PEER MEMORY: a rank
// stores the rows a neighbour needs straight into that neighbour's receive buffer (NVLink / NVSwitch peer mapping of a symmetric
// allocation; inside one process: another part's buffer on the same device) and raises a sequence flag there; the neighbour's next kernel
// waits for the flag and consumes the rows in the same launch.  Two launches per exchange and rank, no NCCL call, no host involvement,
// CUDA-graph replayable (the sequence counters live in device memory, so a replayed launch finds its own place in the sequence).
//
//   fvgn_exchange_push          rows src[send_rows[p][i]] -> peer p's buffer (gather fused with the store), then flag
//   fvgn_exchange_wait_scatter  wait for every peer's flag; dst[recv_rows[p][i]] <- my buffer          (ghost-cell x', ghost next_uv)
//   fvgn_exchange_wait_node_sum wait; interface nodes <- own partial + the peers' partials, added in ascending rank order on every rank
//                               (the same order everywhere: a node shared by three parts gets the same bits on all of them)
//
// Receive regions are double-buffered by the parity of the sequence number: a neighbour that shares only a node (no face) with this rank is
// throttled by nothing but the node exchange itself and may run one exchange ahead.
// Memory ordering: data stores, __threadfence_system(), "last CTA" counter, then st.release.sys of the flag; the waiter spins with
// ld.acquire.sys.  A wait that exceeds ~2 s sets bit 0 of the plan's status word and proceeds (a lost peer must not hang the GPU).
#include "common.cuh"

namespace fvgn {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// copies one row of `row_bytes` (a multiple of 4, <= 512) with `lanes` threads of `unit` bytes each
__device__ __forceinline__ void copy_row(unsigned char* dst, const unsigned char* src, int unit, int lane) {
  if (unit == 16) *reinterpret_cast<uint4*>(dst + 16 * lane) = *reinterpret_cast<const uint4*>(src + 16 * lane);
  else if (unit == 8) *reinterpret_cast<uint2*>(dst + 8 * lane) = *reinterpret_cast<const uint2*>(src + 8 * lane);
  else *reinterpret_cast<uint32_t*>(dst + 4 * lane) = *reinterpret_cast<const uint32_t*>(src + 4 * lane);
}
__device__ __forceinline__ int unit_of(int row_bytes) { return (row_bytes & 15) == 0 ? 16 : (row_bytes & 7) == 0 ? 8 : 4; }

// true for exactly one CTA of the grid: the one that finishes last (all other CTAs' stores are then visible to it)
__device__ __forceinline__ bool last_cta(unsigned int* counter) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    last = (prev == gridDim.x - 1);
    if (last) *counter = 0;
    __threadfence();
  }
  __syncthreads();
  return last;
}

__device__ __forceinline__ void wait_flags(const fvgn_exchange_plan_t& pl, unsigned long long expected) {
  if (threadIdx.x < (unsigned)pl.n_peers) {
    const unsigned long long* f = pl.local_flag[threadIdx.x];
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < expected) {
      if (clock64() - t0 > 4000000000LL) {
        atomicOr(pl.status, 1u);
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) exchange_push_kernel(const __grid_constant__ fvgn_exchange_plan_t pl, const unsigned char* __restrict__ src,
                                                            long long src_stride) {
  const unsigned long long seq = pl.seq[0];  // pushes completed so far (advanced by the last CTA only, after everyone has read it)
  const int unit = unit_of(pl.row_bytes), lanes = pl.row_bytes / unit;
  const int rows_per_cta = 256 / lanes;
  const int lr = threadIdx.x / lanes, lane = threadIdx.x % lanes;
  for (int p = 0; p < pl.n_peers; ++p) {
    unsigned char* dst = reinterpret_cast<unsigned char*>(pl.remote_buf[p]) + pl.remote_off[p] + (long long)(seq & 1) * pl.remote_parity_stride[p];
    const int n = pl.n_send[p];
    if (lr < rows_per_cta)
      for (int i = (int)blockIdx.x * rows_per_cta + lr; i < n; i += (int)gridDim.x * rows_per_cta) {
        const int r = __ldg(pl.send_rows[p] + i);
        copy_row(dst + (long long)i * pl.row_bytes, src + (long long)r * src_stride, unit, lane);
      }
  }
  if (last_cta(pl.counters)) {
    if (threadIdx.x == 0) pl.seq[0] = seq + 1;
    if (threadIdx.x < (unsigned)pl.n_peers) st_release_sys(pl.remote_flag[threadIdx.x], seq + 1);
  }
}

__global__ void __launch_bounds__(256) exchange_wait_scatter_kernel(const __grid_constant__ fvgn_exchange_plan_t pl, unsigned char* __restrict__ dst,
                                                                    long long dst_stride) {
  const unsigned long long seq = pl.seq[1];  // waits completed so far
  wait_flags(pl, seq + 1);
  const int unit = unit_of(pl.row_bytes), lanes = pl.row_bytes / unit;
  const int rows_per_cta = 256 / lanes;
  const int lr = threadIdx.x / lanes, lane = threadIdx.x % lanes;
  for (int p = 0; p < pl.n_peers; ++p) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(pl.local_buf) + pl.local_off[p] + (long long)(seq & 1) * pl.local_parity_stride;
    const int n = pl.n_recv[p];
    if (lr < rows_per_cta)
      for (int i = (int)blockIdx.x * rows_per_cta + lr; i < n; i += (int)gridDim.x * rows_per_cta) {
        const int r = __ldg(pl.recv_rows[p] + i);
        copy_row(dst + (long long)r * dst_stride, src + (long long)i * pl.row_bytes, unit, lane);
      }
  }
  if (last_cta(pl.counters + 1) && threadIdx.x == 0) pl.seq[1] = seq + 1;
}

// node partial sums (rows of 64 floats): half-warp per interface node, 16 B per lane
__global__ void __launch_bounds__(256) exchange_wait_node_sum_kernel(const __grid_constant__ fvgn_exchange_plan_t pl, float* __restrict__ na) {
  const unsigned long long seq = pl.seq[1];
  wait_flags(pl, seq + 1);
  const int hl = threadIdx.x & 15;
  const unsigned char* buf = reinterpret_cast<const unsigned char*>(pl.local_buf) + (long long)(seq & 1) * pl.local_parity_stride;
  for (int r = (int)blockIdx.x * 16 + (threadIdx.x >> 4); r < pl.n_if; r += (int)gridDim.x * 16) {
    const int row = __ldg(pl.if_rows + r);
    float4* own = reinterpret_cast<float4*>(na + (size_t)row * 64 + hl * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // ascending rank order, this rank's own partial at its place (peers are listed in ascending rank; self_pos = peers with a lower rank)
    for (int p = 0; p <= pl.n_peers; ++p) {
      float4 v;
      if (p == pl.self_pos) {
        v = *own;
        acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
      }
      if (p == pl.n_peers) break;
      const int i = __ldg(pl.recv_rows[p] + r);  // position of this node in peer p's message, -1: not shared with p
      if (i >= 0) {
        v = *reinterpret_cast<const float4*>(buf + pl.local_off[p] + (long long)i * 256 + hl * 16);
        acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
      }
    }
    *own = acc;
  }
  if (last_cta(pl.counters + 1) && threadIdx.x == 0) pl.seq[1] = seq + 1;
}

static int check_plan(const fvgn_exchange_plan_t* pl) {
  FVGN_CHECK_ARG(pl != nullptr, "plan is NULL");
  FVGN_CHECK_ARG(pl->n_peers >= 0 && pl->n_peers <= FVGN_EXCHANGE_MAX_PEERS, "too many peers");
  FVGN_CHECK_ARG(pl->row_bytes >= 4 && pl->row_bytes <= 512 && (pl->row_bytes & 3) == 0, "row_bytes must be a multiple of 4, <= 512");
  FVGN_CHECK_ARG(pl->seq && pl->counters && pl->status, "NULL counter pointer");
  return 0;
}

static int grid_for(long long rows, int rows_per_cta) {
  long long g = (rows + rows_per_cta - 1) / rows_per_cta;
  if (g < 1) g = 1;
  const int cap = 2 * sm_count();
  return (int)(g < cap ? g : cap);
}

}  // namespace fvgn

using namespace fvgn;

extern "C" int fvgn_exchange_push(const fvgn_exchange_plan_t* plan, const void* src, long long src_row_stride_bytes, fvgn_stream_t stream) {
  if (int e = check_plan(plan)) return e;
  if (plan->n_peers == 0) return 0;
  FVGN_CHECK_ARG(src != nullptr, "src is NULL");
  long long rows = 0;
  for (int p = 0; p < plan->n_peers; ++p) rows = rows > plan->n_send[p] ? rows : plan->n_send[p];
  const int unit = (plan->row_bytes & 15) == 0 ? 16 : (plan->row_bytes & 7) == 0 ? 8 : 4;
  exchange_push_kernel<<<grid_for(rows, 256 / (plan->row_bytes / unit)), 256, 0, as_stream(stream)>>>(*plan, reinterpret_cast<const unsigned char*>(src),
                                                                                                     src_row_stride_bytes);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_exchange_wait_scatter(const fvgn_exchange_plan_t* plan, void* dst, long long dst_row_stride_bytes, fvgn_stream_t stream) {
  if (int e = check_plan(plan)) return e;
  if (plan->n_peers == 0) return 0;
  FVGN_CHECK_ARG(dst != nullptr, "dst is NULL");
  long long rows = 0;
  for (int p = 0; p < plan->n_peers; ++p) rows = rows > plan->n_recv[p] ? rows : plan->n_recv[p];
  const int unit = (plan->row_bytes & 15) == 0 ? 16 : (plan->row_bytes & 7) == 0 ? 8 : 4;
  exchange_wait_scatter_kernel<<<grid_for(rows, 256 / (plan->row_bytes / unit)), 256, 0, as_stream(stream)>>>(*plan, reinterpret_cast<unsigned char*>(dst),
                                                                                                             dst_row_stride_bytes);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" int fvgn_exchange_wait_node_sum(const fvgn_exchange_plan_t* plan, float* node_agg, fvgn_stream_t stream) {
  if (int e = check_plan(plan)) return e;
  if (plan->n_peers == 0) return 0;
  FVGN_CHECK_ARG(node_agg != nullptr && plan->if_rows != nullptr, "NULL pointer");
  FVGN_CHECK_ARG(plan->row_bytes == 256, "node partial sums are rows of 64 floats");
  exchange_wait_node_sum_kernel<<<grid_for(plan->n_if, 16), 256, 0, as_stream(stream)>>>(*plan, node_agg);
  FVGN_LAUNCH_CHECK();
  return 0;
}
