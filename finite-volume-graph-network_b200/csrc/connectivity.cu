// connectivity.cu — device build of the face/cell connectivity + geometry tables (a1) and of the destination-sorted
// CSR incidence tables the segment-sum kernels walk.
//
// Replaces extract_mesh_state (Extract_mesh/parse_tfrecord_refactor.py:419-971), which loops over cells and faces in
// Python with dicts keyed by str(ndarray).  Here: pack every (max,min) edge into a 64-bit key, LSD radix sort
// (cub::DeviceRadixSort — stable, so equal keys stay in packed order j*C+c), head-flag scan = torch.unique(dim=0)'s
// lexicographic order, then flat per-face / per-cell kernels.  fp32 geometry uses explicitly rounded operations in the
// reference's evaluation order (this file is compiled with --fmad=false) so the sign tests of reorder_face and of the
// outward-normal flip see the same bits as numpy/torch.
#include <cub/cub.cuh>
#include "common.cuh"

namespace fvgn {

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct ConnWs {
  unsigned long long* keys_in;
  unsigned long long* keys_out;
  int* vals_in;
  int* vals_out;
  int* face_id;    // [3C] face index of sorted position i
  unsigned* minmax;  // [2] ordered-uint encoded min / max of face-centre x
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static size_t cub_conn_bytes(int n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, n);
  cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, n);
  return a > b ? a : b;
}

static ConnWs carve_conn(void* base, int C, int K = 3) {
  const size_t n = (size_t)K * C;
  ConnWs w{};
  size_t off = 0;
  char* p = reinterpret_cast<char*>(base);
  w.keys_in = reinterpret_cast<unsigned long long*>(p + off); off += align_up(n * 8);
  w.keys_out = reinterpret_cast<unsigned long long*>(p + off); off += align_up(n * 8);
  w.vals_in = reinterpret_cast<int*>(p + off); off += align_up(n * 4);
  w.vals_out = reinterpret_cast<int*>(p + off); off += align_up(n * 4);
  w.face_id = reinterpret_cast<int*>(p + off); off += align_up(n * 4);
  w.minmax = reinterpret_cast<unsigned*>(p + off); off += 256;
  w.cub_bytes = cub_conn_bytes((int)n);
  w.cub_tmp = p + off; off += align_up(w.cub_bytes);
  w.total = off;
  return w;
}

// ------------------------------------------------------------------------------------------------ phase 1
__global__ void sort_cells_pack_edges(const int32_t* __restrict__ cells, int C, int32_t* __restrict__ cells_node,
                                      unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int a = cells[3 * c], b = cells[3 * c + 1], d = cells[3 * c + 2];
  // np.sort(cells, axis=1)  (:431)
  if (a > b) { int t = a; a = b; b = t; }
  if (b > d) { int t = b; b = d; d = t; }
  if (a > b) { int t = a; a = b; b = t; }
  cells_node[3 * c] = a; cells_node[3 * c + 1] = b; cells_node[3 * c + 2] = d;
  // triangles_to_faces (:1417-1431): edges (0,1) (1,2) (2,0) of the sorted row, stacked [3C], packed (max,min)
  const int e0[3] = {a, b, d}, e1[3] = {b, d, a};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const unsigned hi = (unsigned)max(e0[j], e1[j]), lo = (unsigned)min(e0[j], e1[j]);
    keys[(size_t)j * C + c] = ((unsigned long long)hi << 32) | lo;
    vals[(size_t)j * C + c] = j * C + c;
  }
}

__global__ void head_flags(const unsigned long long* __restrict__ keys, int n, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ phase 2
__device__ __forceinline__ float2 centroid_of(const float* __restrict__ pos, const int32_t* __restrict__ cn, int c) {
  const int n0 = cn[3 * c], n1 = cn[3 * c + 1], n2 = cn[3 * c + 2];
  // ((p0 + p1) + p2) / 3.0 in fp32  (:437-445)
  float2 r;
  r.x = __fdiv_rn(__fadd_rn(__fadd_rn(pos[2 * n0], pos[2 * n1]), pos[2 * n2]), 3.0f);
  r.y = __fdiv_rn(__fadd_rn(__fadd_rn(pos[2 * n0 + 1], pos[2 * n1 + 1]), pos[2 * n2 + 1]), 3.0f);
  return r;
}

__global__ void cell_geometry(const float* __restrict__ pos, const int32_t* __restrict__ cn, int C, float* __restrict__ centroid,
                              float* __restrict__ cell_factor) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float2 g = centroid_of(pos, cn, c);
  centroid[2 * c] = g.x; centroid[2 * c + 1] = g.y;
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // ||p[cn_k] - centroid||^2  (:795-823)
    const int n = cn[3 * c + k];
    const float dx = __fsub_rn(pos[2 * n], g.x), dy = __fsub_rn(pos[2 * n + 1], g.y);
    d[k] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  }
  const float tot = __fadd_rn(__fadd_rn(d[0], d[1]), d[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) cell_factor[3 * c + k] = __fdiv_rn(d[k], tot);
}

// one thread per sorted edge instance; group heads also emit the face row and the (reordered) neighbour cells
__global__ void faces_from_sorted(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, const int* __restrict__ face_id,
                                  int n, int C, int F, const float* __restrict__ pos, const int32_t* __restrict__ cn,
                                  int32_t* __restrict__ face, int32_t* __restrict__ cells_face, int32_t* __restrict__ neighbour_cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int fid = face_id[i] - 1;
  const int packed = vals[i];
  const int j = packed / C, c = packed - j * C;
  cells_face[3 * c + j] = fid;  // cells_face[c,j] = index of packed[j*C+c]  (:627-631,692-695)
  const unsigned long long key = keys[i];
  if (i == 0 || keys[i - 1] != key) {
    face[fid] = (int)(key >> 32);                    // senders = max node id   (:452-456)
    face[(size_t)F + fid] = (int)(key & 0xffffffffu);  // receivers = min node id
    // neighbour_cell (:741-781): the dict keeps [first cell, latest cell] in scan order c=0..C-1, j inner.
    int first = c * 3 + j, last = first;
    for (int t = i + 1; t < n && keys[t] == key; ++t) {
      const int pk = vals[t];
      const int jj = pk / C, cc = pk - jj * C;
      const int s = cc * 3 + jj;
      first = min(first, s);
      last = max(last, s);
    }
    const int ca = first / 3, cb = last / 3;  // boundary face: ca == cb (self-loop, :777-780)
    // reorder_face on centroids (:1343-1375, 784-789): keep (a,b) iff dx > 0 or (dx == 0 and dy > 0)
    const float2 ga = centroid_of(pos, cn, ca), gb = centroid_of(pos, cn, cb);
    const float dx = __fsub_rn(ga.x, gb.x), dy = __fsub_rn(ga.y, gb.y);
    const bool keep = (dx > 0.f) || (dx == 0.f && dy > 0.f);
    neighbour_cell[fid] = keep ? ca : cb;
    neighbour_cell[(size_t)F + fid] = keep ? cb : ca;
  }
}

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void face_geometry(const float* __restrict__ pos, const int32_t* __restrict__ face, int F, float* __restrict__ face_length,
                              unsigned* __restrict__ minmax) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  float fcx = 0.f;
  const bool ok = f < F;
  if (ok) {
    const int s = face[f], r = face[(size_t)F + f];
    const float dx = __fsub_rn(pos[2 * s], pos[2 * r]), dy = __fsub_rn(pos[2 * s + 1], pos[2 * r + 1]);
    // torch.norm(dim=-1) on CPU fp32 reduces as acc = dx*dx ; acc = fma(dy, dy, acc) ; sqrt   (:459-467)
    face_length[f] = __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    fcx = __fdiv_rn(__fadd_rn(pos[2 * s], pos[2 * r]), 2.0f);  // face centre x (:480-488)
  }
  // block reduce min / max of fcx, then one atomic pair per block (min/max are order independent => deterministic)
  unsigned lo = ok ? f2ord(fcx) : 0xffffffffu, hi = ok ? f2ord(fcx) : 0u;
  for (int o = 16; o >= 1; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}

__global__ void face_typing(const float* __restrict__ pos, const int32_t* __restrict__ node_type, const int32_t* __restrict__ face, int F,
                            int rule, const unsigned* __restrict__ minmax, int32_t* __restrict__ face_type) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int s = face[f], r = face[(size_t)F + f];
  const int a = node_type[s], b = node_type[r];
  int t = 0;
  if (rule == 0) {  // "compressible" branch (:491-509) — the one name="incompressible" selects (substring test)
    if (a == b && (a == 2 || a == 0 || a == 4)) t = a;
  } else {          // else branch (:511-620)
    if (a == b && (a == 6 || a == 4 || a == 5 || a == 0)) t = a;
    const float fcx = __fdiv_rn(__fadd_rn(pos[2 * s], pos[2 * r]), 2.0f);
    const float inlet = ord2f(minmax[0]), outlet = ord2f(minmax[1]);
    // limits: np.float32 min/max +- 1e-4 (weak python float => float32 arithmetic under NumPy >= 2)
    if (((a == 6 && b == 4) || (b == 6 && a == 4)) && fcx < __fadd_rn(inlet, 1e-4f) && fcx > __fsub_rn(inlet, 1e-4f)) t = 4;
    if (((a == 6 && b == 5) || (b == 6 && a == 5)) && fcx < __fadd_rn(outlet, 1e-4f) && fcx > __fsub_rn(outlet, 1e-4f)) t = 5;
  }
  face_type[f] = t;
}

__global__ void cell_finalize(const float* __restrict__ pos, const int32_t* __restrict__ face, const int32_t* __restrict__ cells_face,
                              const int32_t* __restrict__ face_type, const float* __restrict__ face_length, const float* __restrict__ centroid,
                              int C, int F, int32_t* __restrict__ cells_type, float* __restrict__ unv, float* __restrict__ cells_area) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int n_in = 0, n_wall = 0, n_out = 0, n_foil = 0, n_norm = 0;
  float len[3];
  const float gx = centroid[2 * c], gy = centroid[2 * c + 1];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int f = cells_face[3 * c + j];
    const int t = face_type[f];
    if (t == 4) ++n_in; else if (t == 6) ++n_wall; else if (t == 5) ++n_out; else if (t == 2) ++n_foil; else ++n_norm;
    len[j] = face_length[f];
    // unit normal (:699-706): (-dy, dx) / ||.||, norm reduced like torch.norm: first^2 then fma(second, second, .)
    const int s = face[f], r = face[(size_t)F + f];
    const float dx = __fsub_rn(pos[2 * s], pos[2 * r]), dy = __fsub_rn(pos[2 * s + 1], pos[2 * r + 1]);
    const float nx0 = -dy, ny0 = dx;
    const float nrm = __fsqrt_rn(__fmaf_rn(ny0, ny0, __fmul_rn(nx0, nx0)));
    float nx = __fdiv_rn(nx0, nrm), ny = __fdiv_rn(ny0, nrm);
    // outward flip (:721-738): flip when n . (centroid - face_centre) > 0
    const float fcx = __fdiv_rn(__fadd_rn(pos[2 * s], pos[2 * r]), 2.0f), fcy = __fdiv_rn(__fadd_rn(pos[2 * s + 1], pos[2 * r + 1]), 2.0f);
    const float vx = __fsub_rn(gx, fcx), vy = __fsub_rn(gy, fcy);
    const float dot = __fadd_rn(__fmul_rn(nx, vx), __fmul_rn(ny, vy));
    if (dot > 0.f) { nx = -nx; ny = -ny; }
    unv[(size_t)c * 6 + 2 * j] = nx;
    unv[(size_t)c * 6 + 2 * j + 1] = ny;
  }
  // cells_type (:639-691)
  int t = 0;
  if (n_wall > 0 && ((n_norm > 0 && n_in == 0 && n_out == 0) || (n_norm == 0 && n_in > 0 && n_out == 0) ||
                     (n_norm == 0 && n_in == 0 && n_out > 0) || (n_norm > 0 && n_in > 0 && n_out == 0) ||
                     (n_norm > 0 && n_in == 0 && n_out > 0)))
    t = 6;
  else if (n_foil > 0 && n_norm > 0 && n_in == 0 && n_out == 0) t = 2;
  else if (n_in > 0 && n_norm > 0 && n_wall == 0 && n_out == 0) t = 4;
  else if (n_out > 0 && n_norm > 0 && n_wall == 0 && n_in == 0) t = 5;
  cells_type[c] = t;
  // Heron (:876-884)
  const float p = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(len[0], len[1]), len[2]));
  cells_area[c] = __fsqrt_rn(__fmul_rn(__fmul_rn(__fmul_rn(p, __fsub_rn(p, len[0])), __fsub_rn(p, len[1])), __fsub_rn(p, len[2])));
}


// ------------------------------------------------------------------------------------------------ tri / quad hybrid meshes
// Cells with k = 3 or 4 nodes ([C,4] tables, 4th column -1 on triangles) — BASELINE configs[3].  The reference is triangles only
// (SURVEY.md §0 item 5); these kernels follow oracle/fvgn_oracle_poly.py, i.e. the reference's rules with "3" read as "k", and reduce to
// the triangle kernels above bit for bit when every cell is a triangle (tests/test_gpu_poly.py).
constexpr unsigned long long POLY_NOKEY = ~0ull;

__global__ void canon_cells_pack_edges_poly(const int32_t* __restrict__ cells, int C, int32_t* __restrict__ cells_node,
                                            unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int n[4] = {cells[4 * c], cells[4 * c + 1], cells[4 * c + 2], cells[4 * c + 3]};
  const bool quad = n[3] >= 0;
  if (!quad) {  // triangles: sorted ascending (:431)
    if (n[0] > n[1]) { int t = n[0]; n[0] = n[1]; n[1] = t; }
    if (n[1] > n[2]) { int t = n[1]; n[1] = n[2]; n[2] = t; }
    if (n[0] > n[1]) { int t = n[0]; n[0] = n[1]; n[1] = t; }
  } else {      // quads: cyclic order kept, smallest id first, reflected so that node[1] < node[3]
    int r = 0;
    for (int j = 1; j < 4; ++j) if (n[j] < n[r]) r = j;
    const int m[4] = {n[r], n[(r + 1) & 3], n[(r + 2) & 3], n[(r + 3) & 3]};
    n[0] = m[0]; n[2] = m[2];
    if (m[1] > m[3]) { n[1] = m[3]; n[3] = m[1]; } else { n[1] = m[1]; n[3] = m[3]; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) cells_node[4 * c + j] = n[j];
  const int k = quad ? 4 : 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unsigned long long key = POLY_NOKEY;
    if (j < k) {
      const int a = n[j], b = n[(j + 1 == k) ? 0 : j + 1];
      key = ((unsigned long long)(unsigned)max(a, b) << 32) | (unsigned)min(a, b);
    }
    keys[(size_t)j * C + c] = key;
    vals[(size_t)j * C + c] = j * C + c;
  }
}

__global__ void head_flags_poly(const unsigned long long* __restrict__ keys, int n, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (keys[i] != POLY_NOKEY && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}

__device__ __forceinline__ float2 centroid_poly(const float* __restrict__ pos, const int32_t* __restrict__ cn, int c) {
  const int n0 = cn[4 * c], n1 = cn[4 * c + 1], n2 = cn[4 * c + 2], n3 = cn[4 * c + 3];
  float sx = __fadd_rn(__fadd_rn(pos[2 * n0], pos[2 * n1]), pos[2 * n2]);
  float sy = __fadd_rn(__fadd_rn(pos[2 * n0 + 1], pos[2 * n1 + 1]), pos[2 * n2 + 1]);
  float2 r;
  if (n3 >= 0) { r.x = __fdiv_rn(__fadd_rn(sx, pos[2 * n3]), 4.0f); r.y = __fdiv_rn(__fadd_rn(sy, pos[2 * n3 + 1]), 4.0f); }
  else { r.x = __fdiv_rn(sx, 3.0f); r.y = __fdiv_rn(sy, 3.0f); }
  return r;
}

__global__ void cell_geometry_poly(const float* __restrict__ pos, const int32_t* __restrict__ cn, int C, float* __restrict__ centroid,
                                   float* __restrict__ cell_factor) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float2 g = centroid_poly(pos, cn, c);
  centroid[2 * c] = g.x; centroid[2 * c + 1] = g.y;
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  const int k = cn[4 * c + 3] >= 0 ? 4 : 3;
  for (int j = 0; j < k; ++j) {
    const int n = cn[4 * c + j];
    const float dx = __fsub_rn(pos[2 * n], g.x), dy = __fsub_rn(pos[2 * n + 1], g.y);
    d[j] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  }
  float tot = __fadd_rn(__fadd_rn(d[0], d[1]), d[2]);
  if (k == 4) tot = __fadd_rn(tot, d[3]);
#pragma unroll
  for (int j = 0; j < 4; ++j) cell_factor[4 * c + j] = __fdiv_rn(d[j], tot);
}

__global__ void faces_from_sorted_poly(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, const int* __restrict__ face_id,
                                       int n, int C, int F, const float* __restrict__ pos, const int32_t* __restrict__ cn,
                                       int32_t* __restrict__ face, int32_t* __restrict__ cells_face, int32_t* __restrict__ neighbour_cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int packed = vals[i];
  const int j = packed / C, c = packed - j * C;
  const unsigned long long key = keys[i];
  if (key == POLY_NOKEY) { cells_face[4 * c + j] = -1; return; }
  const int fid = face_id[i] - 1;
  cells_face[4 * c + j] = fid;
  if (i == 0 || keys[i - 1] != key) {
    face[fid] = (int)(key >> 32);
    face[(size_t)F + fid] = (int)(key & 0xffffffffu);
    int first = c * 4 + j, last = first;  // scan order c = 0..C-1, j = 0..k-1
    for (int t = i + 1; t < n && keys[t] == key; ++t) {
      const int pk = vals[t];
      const int jj = pk / C, cc = pk - jj * C;
      const int s = cc * 4 + jj;
      first = min(first, s);
      last = max(last, s);
    }
    const int ca = first / 4, cb = last / 4;
    const float2 ga = centroid_poly(pos, cn, ca), gb = centroid_poly(pos, cn, cb);
    const float dx = __fsub_rn(ga.x, gb.x), dy = __fsub_rn(ga.y, gb.y);
    const bool keep = (dx > 0.f) || (dx == 0.f && dy > 0.f);
    neighbour_cell[fid] = keep ? ca : cb;
    neighbour_cell[(size_t)F + fid] = keep ? cb : ca;
  }
}

__global__ void cell_finalize_poly(const float* __restrict__ pos, const int32_t* __restrict__ cn, const int32_t* __restrict__ face,
                                   const int32_t* __restrict__ cells_face, const int32_t* __restrict__ face_type,
                                   const float* __restrict__ face_length, const float* __restrict__ centroid, int C, int F,
                                   int32_t* __restrict__ cells_type, float* __restrict__ unv, float* __restrict__ cells_area) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int n_in = 0, n_wall = 0, n_out = 0, n_foil = 0, n_norm = 0;
  float len[3] = {0.f, 0.f, 0.f};
  const float gx = centroid[2 * c], gy = centroid[2 * c + 1];
  const int k = cn[4 * c + 3] >= 0 ? 4 : 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float nx = 0.f, ny = 0.f;
    if (j < k) {
      const int f = cells_face[4 * c + j];
      const int t = face_type[f];
      if (t == 4) ++n_in; else if (t == 6) ++n_wall; else if (t == 5) ++n_out; else if (t == 2) ++n_foil; else ++n_norm;
      if (j < 3) len[j] = face_length[f];
      const int s = face[f], r = face[(size_t)F + f];
      const float dx = __fsub_rn(pos[2 * s], pos[2 * r]), dy = __fsub_rn(pos[2 * s + 1], pos[2 * r + 1]);
      const float nx0 = -dy, ny0 = dx;
      const float nrm = __fsqrt_rn(__fmaf_rn(ny0, ny0, __fmul_rn(nx0, nx0)));
      nx = __fdiv_rn(nx0, nrm); ny = __fdiv_rn(ny0, nrm);
      const float fcx = __fdiv_rn(__fadd_rn(pos[2 * s], pos[2 * r]), 2.0f), fcy = __fdiv_rn(__fadd_rn(pos[2 * s + 1], pos[2 * r + 1]), 2.0f);
      const float vx = __fsub_rn(gx, fcx), vy = __fsub_rn(gy, fcy);
      const float dot = __fadd_rn(__fmul_rn(nx, vx), __fmul_rn(ny, vy));
      if (dot > 0.f) { nx = __fmul_rn(nx, -1.0f); ny = __fmul_rn(ny, -1.0f); }
    }
    unv[(size_t)c * 8 + 2 * j] = nx;
    unv[(size_t)c * 8 + 2 * j + 1] = ny;
  }
  int t = 0;
  if (n_wall > 0 && ((n_norm > 0 && n_in == 0 && n_out == 0) || (n_norm == 0 && n_in > 0 && n_out == 0) ||
                     (n_norm == 0 && n_in == 0 && n_out > 0) || (n_norm > 0 && n_in > 0 && n_out == 0) ||
                     (n_norm > 0 && n_in == 0 && n_out > 0)))
    t = 6;
  else if (n_foil > 0 && n_norm > 0 && n_in == 0 && n_out == 0) t = 2;
  else if (n_in > 0 && n_norm > 0 && n_wall == 0 && n_out == 0) t = 4;
  else if (n_out > 0 && n_norm > 0 && n_wall == 0 && n_in == 0) t = 5;
  cells_type[c] = t;
  if (k == 3) {  // Heron (:876-884)
    const float p = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(len[0], len[1]), len[2]));
    cells_area[c] = __fsqrt_rn(fmaxf(__fmul_rn(__fmul_rn(__fmul_rn(p, __fsub_rn(p, len[0])), __fsub_rn(p, len[1])), __fsub_rn(p, len[2])), 0.0f));
  } else {       // shoelace over the cyclic node order
    float x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int nd = cn[4 * c + j]; x[j] = pos[2 * nd]; y[j] = pos[2 * nd + 1]; }
    auto cr = [&](int a, int b) { return __fsub_rn(__fmul_rn(x[a], y[b]), __fmul_rn(x[b], y[a])); };
    const float sh = __fadd_rn(__fadd_rn(__fadd_rn(cr(0, 1), cr(1, 2)), cr(2, 3)), cr(3, 0));
    cells_area[c] = __fmul_rn(0.5f, fabsf(sh));
  }
}

static int bits_for(long long n) {
  int b = 1;
  while ((1LL << b) < n) ++b;
  return b;
}

}  // namespace fvgn

using namespace fvgn;

extern "C" size_t fvgn_connectivity_workspace_bytes(int n_cells) {
  if (n_cells <= 0) return 256;
  return carve_conn(nullptr, n_cells).total;
}

extern "C" int fvgn_connectivity_faces(const int32_t* cells, int n_cells, int n_nodes, int32_t* cells_node, int* n_faces_host,
                                       void* workspace, size_t workspace_bytes, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(n_cells > 0 && n_nodes > 0, "empty mesh");
  FVGN_CHECK_ARG(cells && cells_node && n_faces_host && workspace, "NULL pointer");
  FVGN_CHECK_ARG((long long)n_cells * 3 < 2147483647LL, "3*C must fit int32");
  ConnWs w = carve_conn(workspace, n_cells);
  if (workspace_bytes < w.total) { set_error("connectivity workspace too small: %zu < %zu", workspace_bytes, w.total); return FVGN_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int n = 3 * n_cells;
  sort_cells_pack_edges<<<cdiv(n_cells, 256), 256, 0, st>>>(cells, n_cells, cells_node, w.keys_in, w.vals_in);
  FVGN_LAUNCH_CHECK();
  const int nb = bits_for(n_nodes);
  size_t tmp = w.cub_bytes;
  FVGN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.vals_out, n, 0, 32 + nb, st));
  // keys_in is free now: reuse its storage for the head flags
  int* flags = reinterpret_cast<int*>(w.keys_in);
  head_flags<<<cdiv(n, 256), 256, 0, st>>>(w.keys_out, n, flags);
  FVGN_LAUNCH_CHECK();
  tmp = w.cub_bytes;
  FVGN_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, flags, w.face_id, n, st));
  FVGN_CUDA(cudaMemcpyAsync(n_faces_host, w.face_id + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  FVGN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int fvgn_connectivity_tables(const float* mesh_pos, const int32_t* node_type, const int32_t* cells_node, int n_cells,
                                        int n_nodes, int n_faces, int face_rule, int32_t* face, int32_t* cells_face,
                                        int32_t* neighbour_cell, int32_t* face_type, int32_t* cells_type, float* face_length,
                                        float* unit_norm_v, float* centroid, float* cells_area, float* cell_factor,
                                        void* workspace, size_t workspace_bytes, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(n_cells > 0 && n_nodes > 0 && n_faces > 0, "empty mesh");
  FVGN_CHECK_ARG(face_rule == 0 || face_rule == 1, "face_rule must be 0 (A) or 1 (B)");
  FVGN_CHECK_ARG(mesh_pos && node_type && cells_node && face && cells_face && neighbour_cell && face_type && cells_type &&
                 face_length && unit_norm_v && centroid && cells_area && cell_factor && workspace, "NULL pointer");
  ConnWs w = carve_conn(workspace, n_cells);
  if (workspace_bytes < w.total) { set_error("connectivity workspace too small"); return FVGN_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int n = 3 * n_cells;
  const unsigned init[2] = {0xffffffffu, 0u};
  FVGN_CUDA(cudaMemcpyAsync(w.minmax, init, sizeof(init), cudaMemcpyHostToDevice, st));
  cell_geometry<<<cdiv(n_cells, 256), 256, 0, st>>>(mesh_pos, cells_node, n_cells, centroid, cell_factor);
  FVGN_LAUNCH_CHECK();
  faces_from_sorted<<<cdiv(n, 256), 256, 0, st>>>(w.keys_out, w.vals_out, w.face_id, n, n_cells, n_faces, mesh_pos, cells_node, face,
                                                  cells_face, neighbour_cell);
  FVGN_LAUNCH_CHECK();
  face_geometry<<<cdiv(n_faces, 256), 256, 0, st>>>(mesh_pos, face, n_faces, face_length, w.minmax);
  FVGN_LAUNCH_CHECK();
  face_typing<<<cdiv(n_faces, 256), 256, 0, st>>>(mesh_pos, node_type, face, n_faces, face_rule, w.minmax, face_type);
  FVGN_LAUNCH_CHECK();
  cell_finalize<<<cdiv(n_cells, 256), 256, 0, st>>>(mesh_pos, face, cells_face, face_type, face_length, centroid, n_cells, n_faces,
                                                    cells_type, unit_norm_v, cells_area);
  FVGN_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t fvgn_connectivity_poly_workspace_bytes(int n_cells) {
  if (n_cells <= 0) return 256;
  return carve_conn(nullptr, n_cells, 4).total;
}

extern "C" int fvgn_connectivity_poly_faces(const int32_t* cells4, int n_cells, int n_nodes, int32_t* cells_node4, int* n_faces_host,
                                            void* workspace, size_t workspace_bytes, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(n_cells > 0 && n_nodes > 0, "empty mesh");
  FVGN_CHECK_ARG(cells4 && cells_node4 && n_faces_host && workspace, "NULL pointer");
  FVGN_CHECK_ARG((long long)n_cells * 4 < 2147483647LL, "4*C must fit int32");
  ConnWs w = carve_conn(workspace, n_cells, 4);
  if (workspace_bytes < w.total) { set_error("connectivity workspace too small: %zu < %zu", workspace_bytes, w.total); return FVGN_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int n = 4 * n_cells;
  canon_cells_pack_edges_poly<<<cdiv(n_cells, 256), 256, 0, st>>>(cells4, n_cells, cells_node4, w.keys_in, w.vals_in);
  FVGN_LAUNCH_CHECK();
  size_t tmp = w.cub_bytes;
  FVGN_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.vals_out, n, 0, 64, st));  // (the padding key is all ones)
  int* flags = reinterpret_cast<int*>(w.keys_in);
  head_flags_poly<<<cdiv(n, 256), 256, 0, st>>>(w.keys_out, n, flags);
  FVGN_LAUNCH_CHECK();
  tmp = w.cub_bytes;
  FVGN_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, flags, w.face_id, n, st));
  FVGN_CUDA(cudaMemcpyAsync(n_faces_host, w.face_id + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  FVGN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int fvgn_connectivity_poly_tables(const float* mesh_pos, const int32_t* node_type, const int32_t* cells_node4, int n_cells,
                                             int n_nodes, int n_faces, int face_rule, int32_t* face, int32_t* cells_face4,
                                             int32_t* neighbour_cell, int32_t* face_type, int32_t* cells_type, float* face_length,
                                             float* unit_norm_v, float* centroid, float* cells_area, float* cell_factor4,
                                             void* workspace, size_t workspace_bytes, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(n_cells > 0 && n_nodes > 0 && n_faces > 0, "empty mesh");
  FVGN_CHECK_ARG(face_rule == 0 || face_rule == 1, "face_rule must be 0 (A) or 1 (B)");
  FVGN_CHECK_ARG(mesh_pos && node_type && cells_node4 && face && cells_face4 && neighbour_cell && face_type && cells_type &&
                 face_length && unit_norm_v && centroid && cells_area && cell_factor4 && workspace, "NULL pointer");
  ConnWs w = carve_conn(workspace, n_cells, 4);
  if (workspace_bytes < w.total) { set_error("connectivity workspace too small"); return FVGN_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int n = 4 * n_cells;
  const unsigned init[2] = {0xffffffffu, 0u};
  FVGN_CUDA(cudaMemcpyAsync(w.minmax, init, sizeof(init), cudaMemcpyHostToDevice, st));
  cell_geometry_poly<<<cdiv(n_cells, 256), 256, 0, st>>>(mesh_pos, cells_node4, n_cells, centroid, cell_factor4);
  FVGN_LAUNCH_CHECK();
  faces_from_sorted_poly<<<cdiv(n, 256), 256, 0, st>>>(w.keys_out, w.vals_out, w.face_id, n, n_cells, n_faces, mesh_pos, cells_node4, face,
                                                       cells_face4, neighbour_cell);
  FVGN_LAUNCH_CHECK();
  face_geometry<<<cdiv(n_faces, 256), 256, 0, st>>>(mesh_pos, face, n_faces, face_length, w.minmax);
  FVGN_LAUNCH_CHECK();
  face_typing<<<cdiv(n_faces, 256), 256, 0, st>>>(mesh_pos, node_type, face, n_faces, face_rule, w.minmax, face_type);
  FVGN_LAUNCH_CHECK();
  cell_finalize_poly<<<cdiv(n_cells, 256), 256, 0, st>>>(mesh_pos, cells_node4, face, cells_face4, face_type, face_length, centroid, n_cells,
                                                         n_faces, cells_type, unit_norm_v, cells_area);
  FVGN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ CSR
namespace fvgn {
__global__ void iota_kernel(int* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void csr_ptr_kernel(const int* __restrict__ sorted_dst, int n_items, int n_dst, int* __restrict__ ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_items) return;
  const int prev = (i == 0) ? -1 : sorted_dst[i - 1];
  const int cur = (i == n_items) ? n_dst : sorted_dst[i];
  for (int d = prev + 1; d <= cur && d <= n_dst; ++d) ptr[d] = i;
}
static size_t cub_csr_bytes(int n) {
  size_t a = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, n);
  return a;
}
}  // namespace fvgn

extern "C" size_t fvgn_csr_workspace_bytes(int n_items) {
  if (n_items <= 0) return 256;
  return 2 * align_up((size_t)n_items * 4) + align_up(cub_csr_bytes(n_items));
}

extern "C" int fvgn_build_csr(const int32_t* dst, int n_items, int n_dst, int32_t* ptr, int32_t* items, void* workspace,
                              size_t workspace_bytes, fvgn_stream_t stream) {
  FVGN_CHECK_ARG(n_items >= 0 && n_dst > 0, "bad sizes");
  FVGN_CHECK_ARG(ptr && (n_items == 0 || (dst && items && workspace)), "NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (n_items == 0) { FVGN_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int) * ((size_t)n_dst + 1), st)); return 0; }
  if (workspace_bytes < fvgn_csr_workspace_bytes(n_items)) { set_error("csr workspace too small"); return FVGN_E_WORKSPACE; }
  char* p = reinterpret_cast<char*>(workspace);
  int* sorted = reinterpret_cast<int*>(p);
  int* iota = reinterpret_cast<int*>(p + align_up((size_t)n_items * 4));
  void* tmp = p + 2 * align_up((size_t)n_items * 4);
  size_t tmp_bytes = cub_csr_bytes(n_items);
  iota_kernel<<<cdiv(n_items, 256), 256, 0, st>>>(iota, n_items);
  FVGN_LAUNCH_CHECK();
  FVGN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dst, sorted, iota, items, n_items, 0, bits_for(n_dst), st));
  csr_ptr_kernel<<<cdiv((long long)n_items + 1, 256), 256, 0, st>>>(sorted, n_items, n_dst, ptr);
  FVGN_LAUNCH_CHECK();
  return 0;
}
