// SpMV row loop on the 2x2 node-block pattern, shared by the single-GPU kernels (solver.cu) and the peer-memory PCG
// (peer_pcg.cu).  GROUP lanes cooperate on one node (two rows), U independent row pairs in flight per lane group.
#pragma once
#include "common.cuh"

// y = mask .* (K x) over all nodes assigned to this thread's warp (grid-stride); returns this thread's share of x'y
// when want_dot (only lanes with sub == 0 contribute).
template <int GROUP, int U>
__device__ __forceinline__ double spmv_rows(const int64_t n_n, const int32_t* __restrict__ nbr_ptr,
                                            const int32_t* __restrict__ nbr_idx, const double* __restrict__ vals,
                                            const double* __restrict__ x, double* __restrict__ y,
                                            const uint8_t* __restrict__ mask, const bool want_dot) {
  constexpr int GPW = 32 / GROUP;  // lane groups per warp
  const int lane = threadIdx.x & 31, sub = lane % GROUP, gi = lane / GROUP;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int NPW = GPW * U;  // nodes per warp per sweep: U independent row pairs in flight per lane group
  double dot = 0.0;
  for (int64_t nb = warp_global * NPW; nb < n_n; nb += n_warps * NPW) {
    int p0[U], deg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t a = nb + u * GPW + gi;
      p0[u] = 0;
      deg[u] = 0;
      if (a < n_n) {
        p0[u] = __ldg(nbr_ptr + a);
        deg[u] = __ldg(nbr_ptr + a + 1) - p0[u];
      }
    }
    int m[U];
    double2 v0[U], v1[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      m[u] = 0;
      v0[u] = v1[u] = make_double2(0.0, 0.0);
      if (sub < deg[u]) {
        const double2* row0 = reinterpret_cast<const double2*>(vals + 4 * (int64_t)p0[u]);
        m[u] = __ldg(nbr_idx + p0[u] + sub);
        v0[u] = __ldcs(row0 + sub);
        v1[u] = __ldcs(row0 + deg[u] + sub);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      xv[u] = make_double2(0.0, 0.0);
      if (sub < deg[u]) xv[u] = __ldg(reinterpret_cast<const double2*>(x) + m[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double acc0 = fma(v0[u].y, xv[u].y, v0[u].x * xv[u].x);
      double acc1 = fma(v1[u].y, xv[u].y, v1[u].x * xv[u].x);
      for (int j = sub + GROUP; j < deg[u]; j += GROUP) {  // rows longer than GROUP blocks
        const double2* row0 = reinterpret_cast<const double2*>(vals + 4 * (int64_t)p0[u]);
        const int mm = __ldg(nbr_idx + p0[u] + j);
        const double2 w0 = __ldcs(row0 + j), w1 = __ldcs(row0 + deg[u] + j);
        const double2 xx = __ldg(reinterpret_cast<const double2*>(x) + mm);
        acc0 = fma(w0.x, xx.x, acc0);
        acc0 = fma(w0.y, xx.y, acc0);
        acc1 = fma(w1.x, xx.x, acc1);
        acc1 = fma(w1.y, xx.y, acc1);
      }
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
      }
      const int64_t a = nb + u * GPW + gi;
      if (sub == 0 && a < n_n) {
        if (mask) {
          const uchar2 mk = reinterpret_cast<const uchar2*>(mask)[a];
          if (!mk.x) acc0 = 0.0;
          if (!mk.y) acc1 = 0.0;
        }
        reinterpret_cast<double2*>(y)[a] = make_double2(acc0, acc1);
        if (want_dot) {
          const double2 xa = __ldg(reinterpret_cast<const double2*>(x) + a);
          dot = fma(xa.x, acc0, dot);
          dot = fma(xa.y, acc1, dot);
        }
      }
    }
  }
  return dot;
}

// lanes per node by block-row length (P1: 7 blocks per row pair -> 2 per lane), nodes in flight, persistent grid size
struct SpmvShape { int group, unroll; unsigned blocks; };
static inline SpmvShape spmv_shape(const fem_plan* P) {
  SpmvShape s;
  s.group = P->max_degree <= 8 ? 4 : (P->max_degree <= 16 ? 8 : 16);
  if (g_fem_tuning.spmv_group == 4 || g_fem_tuning.spmv_group == 8 || g_fem_tuning.spmv_group == 16) s.group = g_fem_tuning.spmv_group;
  s.unroll = g_fem_tuning.spmv_unroll;
  if (s.unroll != 1 && s.unroll != 2 && s.unroll != 4) s.unroll = 2;
  int64_t blocks = (P->n_n * s.group / s.unroll + 255) / 256;
  const int64_t cap = (int64_t)P->sm_count * (g_fem_tuning.spmv_blocks_per_sm > 0 ? g_fem_tuning.spmv_blocks_per_sm : 8);  // persistent grid (measured best), few dot-product atomics
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  s.blocks = (unsigned)blocks;
  return s;
}
