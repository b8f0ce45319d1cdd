// P1 -> P2 midpoint enrichment on the device (reference: create_midpoints_P2,
// tsx-tunnel/pythonFEM.py:1508-1626).
//
// The reference walks the elements and, per edge, searches the whole element table for the neighbour (O(n_e^2)).  Its
// numbering is fully determined by the ORDER in which it meets the edges: occurrence o = 3*element + edge (edges in the
// order V2-V3, V3-V1, V1-V2), and an edge is numbered when its first occurrence is visited.  So
//     midpoint index of an edge = number of edges whose first occurrence is smaller
//                               = exclusive prefix sum over the occurrences of the flag "o is the first occurrence of its edge".
// No sort is needed: one pass inserts the occurrences into an open-addressing hash table keyed by the vertex pair (atomicCAS
// on 64-bit keys; atomicMin/atomicMax keep the first and last occurrence, atomicAdd the count), one scan numbers the edges,
// a second scan numbers the boundary edges (count == 1) in midpoint order, and a last pass writes the outputs.  The atomics
// only decide WHERE an edge lives in the table; every output depends on min/max/count alone, so the result is
// deterministic and bit-identical to the reference's (the coordinates are (x_a + x_b) / 2 of the first occurrence's
// (from, to) vertices, the reference's operand order).
#include "common.cuh"

#include <climits>

int fem_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, cudaStream_t st);  // plan.cu

namespace {

constexpr unsigned long long MID_EMPTY = ~0ull;

struct Midpoints {
  int64_t n_n, n_e, n_occ, cap, n_mid, n_bnd;
  const int32_t* elem;  // caller's [3][n_e]; must stay alive until fem_midpoints_p2_fill
  unsigned long long* keys;
  int32_t *first, *last, *cnt;  // per table slot
  int32_t* slot_of;             // per occurrence
  int32_t* flag;                // per occurrence, then per midpoint (boundary flag)
  int32_t* idx;                 // [n_occ + 1] exclusive scan of the first-occurrence flags
  int32_t* bidx;                // [n_mid + 1] exclusive scan of the boundary flags
  int32_t* status;              // bit 0: edge shared by > 2 triangles; bit 1: inconsistent orientation
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// occurrence o -> (from, to) vertex
__device__ __forceinline__ void edge_of(const int32_t* __restrict__ elem, int64_t n_e, int64_t o, int32_t& a, int32_t& b) {
  const int64_t e = o / 3;
  const int s = (int)(o - 3 * e);
  const int ia = s == 0 ? 1 : (s == 1 ? 2 : 0), ib = s == 0 ? 2 : (s == 1 ? 0 : 1);
  a = elem[ia * n_e + e];
  b = elem[ib * n_e + e];
}

__global__ void mid_insert_kernel(Midpoints M) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= M.n_occ) return;
  int32_t a, b;
  edge_of(M.elem, M.n_e, o, a, b);
  const unsigned long long key = (unsigned long long)min(a, b) * (unsigned long long)M.n_n + (unsigned long long)max(a, b);
  const unsigned long long mask = (unsigned long long)M.cap - 1;
  unsigned long long h = mix64(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(&M.keys[h], MID_EMPTY, key);
    if (prev == MID_EMPTY || prev == key) break;
    h = (h + 1) & mask;
  }
  atomicMin(&M.first[h], (int32_t)o);
  atomicMax(&M.last[h], (int32_t)o);
  atomicAdd(&M.cnt[h], 1);
  M.slot_of[o] = (int32_t)h;
}

__global__ void mid_first_flag_kernel(Midpoints M) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= M.n_occ) return;
  const int32_t h = M.slot_of[o];
  const bool is_first = M.first[h] == (int32_t)o;
  M.flag[o] = is_first ? 1 : 0;
  if (!is_first) return;
  const int32_t c = M.cnt[h];
  if (c > 2) atomicOr(M.status, 1);
  if (c == 2) {  // the reference's neighbour-slot rule presumes the two triangles traverse the edge in opposite directions
    int32_t a, b, a2, b2;
    edge_of(M.elem, M.n_e, o, a, b);
    edge_of(M.elem, M.n_e, M.last[h], a2, b2);
    if (a != b2 || b != a2) atomicOr(M.status, 2);
  }
}

// per first occurrence: boundary flag of its midpoint
__global__ void mid_boundary_flag_kernel(Midpoints M) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= M.n_occ) return;
  const int32_t h = M.slot_of[o];
  if (M.first[h] != (int32_t)o) return;
  M.flag[M.idx[o]] = M.cnt[h] == 1 ? 1 : 0;
}

__global__ void mid_fill_kernel(Midpoints M, const double* __restrict__ coord, double* __restrict__ coord_mid,
                                int32_t* __restrict__ elem_mid, int32_t* __restrict__ edge_el, int32_t* __restrict__ surf) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= M.n_occ) return;
  const int32_t h = M.slot_of[o];
  const int32_t f = M.first[h];
  const int32_t m = M.idx[f];
  const int64_t e = o / 3;
  elem_mid[(o - 3 * e) * M.n_e + e] = m;
  if (f != (int32_t)o) return;
  int32_t a, b;
  edge_of(M.elem, M.n_e, o, a, b);
  coord_mid[m] = (coord[a] + coord[b]) / 2;
  coord_mid[M.n_mid + m] = (coord[M.n_n + a] + coord[M.n_n + b]) / 2;
  const bool shared = M.cnt[h] == 2;
  edge_el[m] = (int32_t)e;
  edge_el[M.n_mid + m] = shared ? M.last[h] / 3 : 0;
  if (!shared) {
    const int32_t k = M.bidx[m];
    surf[k] = b;
    surf[M.n_bnd + k] = a;
    surf[2 * M.n_bnd + k] = m + (int32_t)M.n_n;
  }
}

void mid_free(Midpoints* M) {
  cudaFree(M->keys);
  cudaFree(M->first);
  cudaFree(M->last);
  cudaFree(M->cnt);
  cudaFree(M->slot_of);
  cudaFree(M->flag);
  cudaFree(M->idx);
  cudaFree(M->bidx);
  cudaFree(M->status);
  delete M;
}

}  // namespace

extern "C" int fem_midpoints_p2_count(int64_t n_n, int64_t n_e, const int32_t* elem, void** handle, int64_t* n_mid, int64_t* n_bnd,
                                      int* status, fem_stream stream) {
  FEM_REQUIRE(elem && handle && n_mid && n_bnd && status, "null pointer");
  FEM_REQUIRE(n_n > 0 && n_e > 0 && 3 * n_e < INT_MAX && n_n < INT_MAX, "mesh size outside the 32-bit occurrence index");
  cudaStream_t st = (cudaStream_t)stream;
  Midpoints* M = new Midpoints();
  M->n_n = n_n;
  M->n_e = n_e;
  M->n_occ = 3 * n_e;
  M->elem = elem;
  M->cap = 1024;
  while (M->cap < 3 * n_e) M->cap *= 2;  // unique edges ~ 1.5 n_e: load factor <= 1/2
  int rc = FEM_OK;
  do {
#define MID_TRY(expr)                                                                                   \
  {                                                                                                     \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      fem_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e));        \
      rc = FEM_ERR_CUDA;                                                                                \
      break;                                                                                            \
    }                                                                                                   \
  }
    MID_TRY(cudaMalloc(&M->keys, sizeof(unsigned long long) * M->cap));
    MID_TRY(cudaMalloc(&M->first, sizeof(int32_t) * M->cap));
    MID_TRY(cudaMalloc(&M->last, sizeof(int32_t) * M->cap));
    MID_TRY(cudaMalloc(&M->cnt, sizeof(int32_t) * M->cap));
    MID_TRY(cudaMalloc(&M->slot_of, sizeof(int32_t) * M->n_occ));
    MID_TRY(cudaMalloc(&M->flag, sizeof(int32_t) * M->n_occ));
    MID_TRY(cudaMalloc(&M->idx, sizeof(int32_t) * (M->n_occ + 1)));
    MID_TRY(cudaMalloc(&M->status, sizeof(int32_t)));
    MID_TRY(cudaMemsetAsync(M->keys, 0xFF, sizeof(unsigned long long) * M->cap, st));
    MID_TRY(cudaMemsetAsync(M->first, 0x7F, sizeof(int32_t) * M->cap, st));  // 0x7f7f7f7f > any occurrence
    MID_TRY(cudaMemsetAsync(M->last, 0xFF, sizeof(int32_t) * M->cap, st));   // -1
    MID_TRY(cudaMemsetAsync(M->cnt, 0, sizeof(int32_t) * M->cap, st));
    MID_TRY(cudaMemsetAsync(M->status, 0, sizeof(int32_t), st));
    const unsigned grid = (unsigned)fem_div_up(M->n_occ, 256);
    mid_insert_kernel<<<grid, 256, 0, st>>>(*M);
    mid_first_flag_kernel<<<grid, 256, 0, st>>>(*M);
    MID_TRY(cudaGetLastError());
    if ((rc = fem_exclusive_scan_i32(M->flag, M->idx, M->n_occ, st)) != FEM_OK) break;
    int32_t h_mid = 0, h_status = 0;
    MID_TRY(cudaMemcpyAsync(&h_mid, M->idx + M->n_occ, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MID_TRY(cudaMemcpyAsync(&h_status, M->status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MID_TRY(cudaStreamSynchronize(st));
    M->n_mid = h_mid;
    *status = h_status;
    MID_TRY(cudaMalloc(&M->bidx, sizeof(int32_t) * (M->n_mid + 1)));
    mid_boundary_flag_kernel<<<grid, 256, 0, st>>>(*M);
    MID_TRY(cudaGetLastError());
    if ((rc = fem_exclusive_scan_i32(M->flag, M->bidx, M->n_mid, st)) != FEM_OK) break;
    int32_t h_bnd = 0;
    MID_TRY(cudaMemcpyAsync(&h_bnd, M->bidx + M->n_mid, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MID_TRY(cudaStreamSynchronize(st));
    M->n_bnd = h_bnd;
#undef MID_TRY
  } while (0);
  if (rc != FEM_OK) {
    mid_free(M);
    return rc;
  }
  *n_mid = M->n_mid;
  *n_bnd = M->n_bnd;
  *handle = M;
  return FEM_OK;
}

extern "C" int fem_midpoints_p2_fill(void* handle, const double* coord, double* coord_mid, int32_t* elem_mid, int32_t* edge_el,
                                     int32_t* surf, fem_stream stream) {
  FEM_REQUIRE(handle && coord && coord_mid && elem_mid && edge_el, "null pointer");
  Midpoints* M = (Midpoints*)handle;
  FEM_REQUIRE(surf || M->n_bnd == 0, "surf is null but the mesh has boundary edges");
  mid_fill_kernel<<<(unsigned)fem_div_up(M->n_occ, 256), 256, 0, (cudaStream_t)stream>>>(*M, coord, coord_mid, elem_mid, edge_el, surf);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_midpoints_p2_destroy(void* handle, fem_stream stream) {
  if (!handle) return FEM_OK;
  FEM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  mid_free((Midpoints*)handle);
  return FEM_OK;
}
