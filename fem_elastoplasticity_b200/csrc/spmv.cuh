// SpMV row loop on the 2x2 node-block pattern, shared by the single-GPU kernels (solver.cu) and the peer-memory PCG
// (peer_pcg.cu).  GROUP lanes cooperate on one node (two rows), U independent row pairs in flight per lane group.
#pragma once
#include "common.cuh"

// y = mask .* (K x) over all nodes assigned to this thread's warp (grid-stride); returns this thread's share of x'y
// when want_dot (only lanes with sub == 0 contribute).
// x gather with an L2 evict_last policy: the matrix values stream through L2 (ld.global.cs) and would otherwise push the
// few node rows of x the active window needs out of it
__device__ __forceinline__ double2 ldg_x_keep(const double* x, int m, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(reinterpret_cast<const double2*>(x) + m), "l"(pol));
  return v;
}

// Matrix values are FP64 (the operator of the CG) or FP32 (the copy the multigrid smoother streams: a preconditioner
// only needs a fixed symmetric approximation, and the SpMV is bound by the bytes of the values).  Products and sums are
// FP64 either way.
template <class VT> struct SpmvPair;
template <> struct SpmvPair<double> { using type = double2; };
template <> struct SpmvPair<float> { using type = float2; };
template <class VT>
__device__ __forceinline__ double2 spmv_ld_pair(const typename SpmvPair<VT>::type* p) {
  const typename SpmvPair<VT>::type v = __ldcs(p);
  return make_double2((double)v.x, (double)v.y);
}

// What happens to the two row sums (K x)_{2a}, (K x)_{2a+1} of node a.  An epilogue has two parts:
//   prefetch(a, sub)   called by EVERY lane of the node's group right after the row extent is known, i.e. together with the
//                      loads of the matrix values: operands the epilogue needs (one double2 per lane) are then in flight
//                      during the gather instead of costing a second, dependent memory round trip per tile;
//   operator()(...)    called by every lane (warp-converged; `lead` marks the lane that holds the sums of a valid node) -
//                      the lead lane may collect the other lanes' prefetched operands with shuffles.
// The default stores mask .* (K x) and accumulates x'y; the multigrid smoother (mg.cu) fuses its vector updates here.
template <bool COHERENT>
struct SpmvStoreEpilogue {
  double* __restrict__ y;
  const uint8_t* __restrict__ mask;
  const double* x;
  bool want_dot;
  __device__ __forceinline__ double2 prefetch(const int64_t a, const int sub) const {
    if (want_dot && sub == 0) return COHERENT ? __ldcg(reinterpret_cast<const double2*>(x) + a) : __ldg(reinterpret_cast<const double2*>(x) + a);
    return make_double2(0.0, 0.0);
  }
  __device__ __forceinline__ void operator()(const int64_t a, const bool lead, double acc0, double acc1, double& dot, const double2 pf) const {
    if (!lead) return;
    if (mask) {
      const uchar2 mk = reinterpret_cast<const uchar2*>(mask)[a];
      if (!mk.x) acc0 = 0.0;
      if (!mk.y) acc1 = 0.0;
    }
    reinterpret_cast<double2*>(y)[a] = make_double2(acc0, acc1);
    if (want_dot) {
      dot = fma(pf.x, acc0, dot);
      dot = fma(pf.y, acc1, dot);
    }
  }
};

template <int GROUP, int U, class EPI, class VT = double>
__device__ __forceinline__ double spmv_rows_epi(const int64_t n_n, const int32_t* __restrict__ nbr_ptr,
                                            const int32_t* __restrict__ nbr_idx, const VT* __restrict__ vals,
                                            const double* __restrict__ x, const EPI& epi) {
  constexpr int GPW = 32 / GROUP;  // lane groups per warp
  const int lane = threadIdx.x & 31, sub = lane % GROUP, gi = lane / GROUP;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int NPW = GPW * U;  // nodes per warp per sweep: U independent row pairs in flight per lane group
  double dot = 0.0;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
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
        using P2 = typename SpmvPair<VT>::type;
        const P2* row0 = reinterpret_cast<const P2*>(vals + 4 * (int64_t)p0[u]);
        m[u] = __ldg(nbr_idx + p0[u] + sub);
        v0[u] = spmv_ld_pair<VT>(row0 + sub);
        v1[u] = spmv_ld_pair<VT>(row0 + deg[u] + sub);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      xv[u] = make_double2(0.0, 0.0);
      if (sub < deg[u]) xv[u] = ldg_x_keep(x, m[u], pol);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double acc0 = fma(v0[u].y, xv[u].y, v0[u].x * xv[u].x);
      double acc1 = fma(v1[u].y, xv[u].y, v1[u].x * xv[u].x);
      for (int j = sub + GROUP; j < deg[u]; j += GROUP) {  // rows longer than GROUP blocks
        using P2 = typename SpmvPair<VT>::type;
        const P2* row0 = reinterpret_cast<const P2*>(vals + 4 * (int64_t)p0[u]);
        const int mm = __ldg(nbr_idx + p0[u] + j);
        const double2 w0 = spmv_ld_pair<VT>(row0 + j), w1 = spmv_ld_pair<VT>(row0 + deg[u] + j);
        const double2 xx = ldg_x_keep(x, mm, pol);
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
      epi(a, sub == 0 && a < n_n, acc0, acc1, dot, (a < n_n) ? epi.prefetch(a, sub) : make_double2(0.0, 0.0));
    }
  }
  return dot;
}

template <int GROUP, int U>
__device__ __forceinline__ double spmv_rows(const int64_t n_n, const int32_t* __restrict__ nbr_ptr,
                                            const int32_t* __restrict__ nbr_idx, const double* __restrict__ vals,
                                            const double* __restrict__ x, double* __restrict__ y,
                                            const uint8_t* __restrict__ mask, const bool want_dot) {
  const SpmvStoreEpilogue<false> epi{y, mask, x, want_dot};
  return spmv_rows_epi<GROUP, U>(n_n, nbr_ptr, nbr_idx, vals, x, epi);
}


// ---- x staged in shared memory ------------------------------------------------------------------------------------
// A CTA walks tiles of FEM_SPMV_TILE consecutive nodes.  The columns a tile references are <= FEM_SPMV_MAXSEG contiguous
// node ranges of x (plan: tile_seg); one thread brings them into shared memory with bulk async copies (cp.async.bulk ->
// UBLKCP, completion on an mbarrier) while all threads already stream the tile's matrix values; the gather then reads
// shared memory through the 16-bit positions of nbr_loc.  x is read once per tile (three tiles share a node row and meet
// in L2) instead of once per block through L1/L2, and the index stream is 2 B instead of 4 B per block.  The copies of
// tile i+1 are issued before tile i is computed (two buffers).  Tiles whose columns do not fit gather from global memory.
__device__ __forceinline__ uint32_t spmv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct SpmvTileSmem {
  double2 xbuf[2][FEM_SPMV_CAP];
  uint64_t bar[2];
};

// thread 0: expect the bytes of the tile's ranges and issue the copies
__device__ __forceinline__ void spmv_issue_tile(const int32_t* __restrict__ tile_seg, int64_t tile, const double* x, double2* dst, uint64_t* bar) {
  const int32_t* d = tile_seg + tile * FEM_SPMV_DESC;
  const int nseg = d[0];
  if (nseg <= 0) return;
  const uint32_t b = spmv_smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)d[1] * 16u) : "memory");
  int off = 0;
  for (int k = 0; k < nseg; ++k) {
    const int start = d[2 + 2 * k], len = d[3 + 2 * k];
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(spmv_smem_u32(dst + off)),
                 "l"(reinterpret_cast<const double2*>(x) + start), "r"((uint32_t)len * 16u), "r"(b)
                 : "memory");
    off += len;
  }
}

__device__ __forceinline__ void spmv_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = spmv_smem_u32(bar);
  uint32_t ok = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(b), "r"(parity)
                 : "memory");
  } while (!ok);
}

template <bool COHERENT>
__device__ __forceinline__ double2 spmv_ldx(const double* x, int m) {
  // COHERENT: x may be stored by a peer GPU while the kernel is resident (fused multi-GPU PCG): no non-coherent loads
  return COHERENT ? __ldcg(reinterpret_cast<const double2*>(x) + m) : __ldg(reinterpret_cast<const double2*>(x) + m);
}

// blockDim.x == FEM_SPMV_THREADS.  Returns this thread's share of x'y when want_dot.
template <int GROUP, bool COHERENT, class EPI, class VT = double>
__device__ __forceinline__ double spmv_tiles_epi(const int64_t n_n, const int64_t n_tiles, const int32_t* __restrict__ nbr_ptr,
                                                 const int32_t* __restrict__ nbr_idx, const uint16_t* __restrict__ nbr_loc,
                                                 const int32_t* __restrict__ tile_seg, const VT* __restrict__ vals, const double* x,
                                                 const EPI& epi, SpmvTileSmem& sm) {
  constexpr int U = 2, B = 2;                  // row pairs in flight per lane group, blocks per lane loaded up front
  constexpr int GPC = FEM_SPMV_THREADS / GROUP;  // lane groups per CTA
  constexpr int NPS = GPC * U;                 // nodes per sweep
  constexpr int SWEEPS = FEM_SPMV_TILE / NPS;
  static_assert(FEM_SPMV_TILE % NPS == 0, "tile size");
  const int sub = threadIdx.x % GROUP, gi = threadIdx.x / GROUP;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spmv_smem_u32(&sm.bar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spmv_smem_u32(&sm.bar[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (COHERENT) asm volatile("fence.proxy.async;" ::: "memory");  // peer stores observed through the flags precede the async-proxy reads
    if ((int64_t)blockIdx.x < n_tiles) spmv_issue_tile(tile_seg, blockIdx.x, x, sm.xbuf[0], &sm.bar[0]);
  }
  __syncthreads();
  uint32_t phase[2] = {0u, 0u};
  double dot = 0.0;
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int cur = it & 1;
    const int64_t next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < n_tiles) spmv_issue_tile(tile_seg, next, x, sm.xbuf[cur ^ 1], &sm.bar[cur ^ 1]);
    const bool staged = tile_seg[tile * FEM_SPMV_DESC] > 0;  // CTA-uniform
    const double2* xs = sm.xbuf[cur];
    bool waited = false;
#pragma unroll 1
    for (int sw = 0; sw < SWEEPS; ++sw) {
      int p0[U], deg[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t a = tile * FEM_SPMV_TILE + sw * NPS + u * GPC + gi;
        p0[u] = 0;
        deg[u] = 0;
        if (a < n_n) {
          p0[u] = __ldg(nbr_ptr + a);
          deg[u] = __ldg(nbr_ptr + a + 1) - p0[u];
        }
      }
      int m[U][B];
      double2 v0[U][B], v1[U][B];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const int j = sub + b * GROUP;
          m[u][b] = 0;
          v0[u][b] = v1[u][b] = make_double2(0.0, 0.0);
          if (j < deg[u]) {
            using P2 = typename SpmvPair<VT>::type;
            const P2* row0 = reinterpret_cast<const P2*>(vals + 4 * (int64_t)p0[u]);
            m[u][b] = staged ? (int)__ldcs(nbr_loc + p0[u] + j) : __ldg(nbr_idx + p0[u] + j);
            v0[u][b] = spmv_ld_pair<VT>(row0 + j);
            v1[u][b] = spmv_ld_pair<VT>(row0 + deg[u] + j);
          }
        }
      if (staged && !waited) {  // the matrix values above are in flight while the x ranges arrive
        spmv_bar_wait(&sm.bar[cur], phase[cur]);
        waited = true;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const int j = sub + b * GROUP;
          if (j < deg[u]) {
            const double2 xv = staged ? xs[m[u][b]] : spmv_ldx<COHERENT>(x, m[u][b]);
            acc0 = fma(v0[u][b].x, xv.x, acc0);
            acc0 = fma(v0[u][b].y, xv.y, acc0);
            acc1 = fma(v1[u][b].x, xv.x, acc1);
            acc1 = fma(v1[u][b].y, xv.y, acc1);
          }
        }
        for (int j = sub + B * GROUP; j < deg[u]; j += GROUP) {  // rows longer than B*GROUP blocks
          using P2 = typename SpmvPair<VT>::type;
          const P2* row0 = reinterpret_cast<const P2*>(vals + 4 * (int64_t)p0[u]);
          const double2 w0 = spmv_ld_pair<VT>(row0 + j), w1 = spmv_ld_pair<VT>(row0 + deg[u] + j);
          const double2 xx = staged ? xs[__ldcs(nbr_loc + p0[u] + j)] : spmv_ldx<COHERENT>(x, __ldg(nbr_idx + p0[u] + j));
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
        const int64_t a = tile * FEM_SPMV_TILE + sw * NPS + u * GPC + gi;
        epi(a, sub == 0 && a < n_n, acc0, acc1, dot, (a < n_n) ? epi.prefetch(a, sub) : make_double2(0.0, 0.0));
      }
    }
    if (staged) {
      if (!waited) spmv_bar_wait(&sm.bar[cur], phase[cur]);
      phase[cur] ^= 1u;
    }
    __syncthreads();  // xbuf[cur] may be refilled (by the copies for tile it+2, issued at the top of the next iteration)
  }
  return dot;
}

template <int GROUP, bool COHERENT>
__device__ __forceinline__ double spmv_tiles(const int64_t n_n, const int64_t n_tiles, const int32_t* __restrict__ nbr_ptr,
                                             const int32_t* __restrict__ nbr_idx, const uint16_t* __restrict__ nbr_loc,
                                             const int32_t* __restrict__ tile_seg, const double* __restrict__ vals, const double* x,
                                             double* __restrict__ y, const uint8_t* __restrict__ mask, const bool want_dot,
                                             SpmvTileSmem& sm) {
  const SpmvStoreEpilogue<COHERENT> epi{y, mask, x, want_dot};
  return spmv_tiles_epi<GROUP, COHERENT>(n_n, n_tiles, nbr_ptr, nbr_idx, nbr_loc, tile_seg, vals, x, epi, sm);
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
static inline bool spmv_use_tiles(const fem_plan* P) { return P->tile_seg != nullptr && g_fem_tuning.spmv_staged != 1; }
static inline unsigned spmv_tile_blocks(const fem_plan* P) {
  int64_t blocks = P->n_tiles;
  const int64_t cap = (int64_t)P->sm_count * (g_fem_tuning.spmv_blocks_per_sm > 0 ? g_fem_tuning.spmv_blocks_per_sm : 65536 / (64 * FEM_SPMV_THREADS));  // one resident wave (64 registers)
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}
