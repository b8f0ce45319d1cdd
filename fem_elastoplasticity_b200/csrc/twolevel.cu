// Two-level additive preconditioner for the masked PCG:  M^-1 = D^-1 + P A_c^-1 P^T.
// P interpolates a coarse bilinear (Q1) grid of ncx x ncy cells laid over the mesh's bounding box to the fine nodes
// (both displacement components separately, rows of Dirichlet DOFs zeroed); A_c = P^T K P is the Galerkin coarse operator,
// inverted once (dense, n_c = 2 (ncx+1)(ncy+1) ~ 8 450) and applied by a dense GEMV.  Point-Jacobi PCG needs O(N) iterations
// on the reference's footing problem (57 500 at 16 M elements); with the coarse correction the count depends on H/h only.
// The reference solves this system with a dense LU (Plasticity2D_DP/pythonFEM.py:1062-1066); any SPD preconditioner yields
// the same solution, so parity is unaffected.
#include "common.cuh"

struct CoarseGrid {
  double x0, y0, inv_hx, inv_hy;
  int ncx, ncy;
};

// coarse cell, bilinear weights and the four coarse node ids of a fine node
__device__ __forceinline__ int coarse_of(const CoarseGrid& g, double x, double y, double (&w)[4], int (&id)[4]) {
  double fx = (x - g.x0) * g.inv_hx, fy = (y - g.y0) * g.inv_hy;
  int cx = (int)fx, cy = (int)fy;
  cx = cx < 0 ? 0 : (cx > g.ncx - 1 ? g.ncx - 1 : cx);
  cy = cy < 0 ? 0 : (cy > g.ncy - 1 ? g.ncy - 1 : cy);
  double xi = fx - cx, et = fy - cy;
  xi = xi < 0.0 ? 0.0 : (xi > 1.0 ? 1.0 : xi);
  et = et < 0.0 ? 0.0 : (et > 1.0 ? 1.0 : et);
  w[0] = (1.0 - xi) * (1.0 - et);
  w[1] = xi * (1.0 - et);
  w[2] = xi * et;
  w[3] = (1.0 - xi) * et;
  const int base = cx + cy * (g.ncx + 1);
  id[0] = base;
  id[1] = base + 1;
  id[2] = base + 1 + (g.ncx + 1);
  id[3] = base + (g.ncx + 1);
  return cx + cy * g.ncx;
}

// ---- A_c = P^T K P (once per matrix) ---------------------------------------------------------------------------
__global__ void coarse_galerkin_kernel(int64_t n_n, CoarseGrid g, const int32_t* __restrict__ nbr_ptr,
                                       const int32_t* __restrict__ nbr_idx, const double* __restrict__ vals,
                                       const uint8_t* __restrict__ mask, const uint8_t* __restrict__ cmask,
                                       const double* __restrict__ coord, double* Ac, int ncd) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_n; a += (int64_t)gridDim.x * blockDim.x) {
    double wi[4], wj[4];
    int ii[4], ij[4];
    coarse_of(g, coord[a], coord[n_n + a], wi, ii);
    const double mi0 = mask ? (double)(mask[2 * a] != 0) : 1.0, mi1 = mask ? (double)(mask[2 * a + 1] != 0) : 1.0;
    const int p0 = nbr_ptr[a], deg = nbr_ptr[a + 1] - p0;
    const double* row0 = vals + 4 * (int64_t)p0;
    const double* row1 = row0 + 2 * deg;
    for (int j = 0; j < deg; ++j) {
      const int b = nbr_idx[p0 + j];
      coarse_of(g, coord[b], coord[n_n + b], wj, ij);
      const double mj0 = cmask ? (double)(cmask[2 * b] != 0) : 1.0, mj1 = cmask ? (double)(cmask[2 * b + 1] != 0) : 1.0;
      const double a00 = row0[2 * j] * mi0 * mj0, a01 = row0[2 * j + 1] * mi0 * mj1;
      const double a10 = row1[2 * j] * mi1 * mj0, a11 = row1[2 * j + 1] * mi1 * mj1;
#pragma unroll
      for (int ki = 0; ki < 4; ++ki) {
        if (wi[ki] == 0.0) continue;
#pragma unroll
        for (int kj = 0; kj < 4; ++kj) {
          const double ww = wi[ki] * wj[kj];
          if (ww == 0.0) continue;
          double* dst = Ac + (int64_t)(2 * ii[ki]) * ncd + 2 * ij[kj];
          if (a00 != 0.0) atomicAdd(dst, ww * a00);
          if (a01 != 0.0) atomicAdd(dst + 1, ww * a01);
          if (a10 != 0.0) atomicAdd(dst + ncd, ww * a10);
          if (a11 != 0.0) atomicAdd(dst + ncd + 1, ww * a11);
        }
      }
    }
  }
}

// ---- warp-aggregated restriction r_c += P^T r of one node per lane ------------------------------------------------
__device__ __forceinline__ void restrict_warp(const CoarseGrid& g, bool valid, double x, double y, double rx, double ry, double* rc) {
  double w[4];
  int id[4];
  int cell = -1;
  if (valid) cell = coarse_of(g, x, y, w, id);
  const int lane = threadIdx.x & 31;
  unsigned remaining = __ballot_sync(0xffffffffu, cell >= 0);
  while (remaining) {  // warp-uniform: one round per distinct coarse cell among the 32 nodes (1-2 on ordered meshes)
    const int leader = __ffs(remaining) - 1;
    const int lc = __shfl_sync(0xffffffffu, cell, leader);
    const bool peer = (cell == lc);
    const unsigned peers = __ballot_sync(0xffffffffu, peer);
    double v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = peer ? w[k] * rx : 0.0;
      v[2 * k + 1] = peer ? w[k] * ry : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = warp_sum(v[q]);
    if (lane == leader) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (v[2 * k] != 0.0) atomicAdd(rc + 2 * id[k], v[2 * k]);
        if (v[2 * k + 1] != 0.0) atomicAdd(rc + 2 * id[k] + 1, v[2 * k + 1]);
      }
    }
    remaining &= ~peers;
  }
}

__device__ __forceinline__ double2 prolong(const CoarseGrid& g, double x, double y, const double* __restrict__ zc) {
  double w[4];
  int id[4];
  coarse_of(g, x, y, w, id);
  double2 z = make_double2(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double2 c = *reinterpret_cast<const double2*>(zc + 2 * id[k]);
    z.x = fma(w[k], c.x, z.x);
    z.y = fma(w[k], c.y, z.y);
  }
  return z;
}

// scal layout as in solver.cu: [0] r'z even, [1] r'r, [2] r'z odd, [3] p'Kp, [4] |b|^2
__device__ __forceinline__ int rz_old_slot2(int it) { return (it & 1) ? 2 : 0; }
__device__ __forceinline__ int rz_new_slot2(int it) { return (it & 1) ? 0 : 2; }

// r = mask (b - K x0);  r_c += P^T r;  |b|^2, r'r
__global__ void __launch_bounds__(256) tl_init_kernel(int64_t n_n, CoarseGrid g, const double2* __restrict__ rhs, const double2* __restrict__ Kx0,
                                                      const uint8_t* __restrict__ mask, const double2* __restrict__ minv,
                                                      const double* __restrict__ coord, double2* __restrict__ r, double* rc, double* scal) {
  __shared__ double red[32];
  double rr = 0.0, bb = 0.0, rdr = 0.0;
  const int64_t n_pad = (n_n + 31) & ~(int64_t)31;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_pad; a += (int64_t)gridDim.x * blockDim.x) {
    const bool valid = a < n_n;
    double2 ri = make_double2(0.0, 0.0);
    double x = 0.0, y = 0.0;
    if (valid) {
      const bool f0 = mask ? mask[2 * a] != 0 : true, f1 = mask ? mask[2 * a + 1] != 0 : true;
      double2 b = rhs[a];
      b.x = f0 ? b.x : 0.0;
      b.y = f1 ? b.y : 0.0;
      ri = b;
      if (Kx0) {
        const double2 k = Kx0[a];
        ri.x = f0 ? b.x - k.x : 0.0;
        ri.y = f1 ? b.y - k.y : 0.0;
      }
      r[a] = ri;
      rr = fma(ri.x, ri.x, fma(ri.y, ri.y, rr));
      bb = fma(b.x, b.x, fma(b.y, b.y, bb));
      const double2 mi = minv[a];
      rdr = fma(ri.x * mi.x, ri.x, fma(ri.y * mi.y, ri.y, rdr));
      x = coord[a];
      y = coord[n_n + a];
    }
    restrict_warp(g, valid, x, y, ri.x, ri.y, rc);
  }
  rr = block_sum(rr, red);
  bb = block_sum(bb, red);
  rdr = block_sum(rdr, red);
  if (threadIdx.x == 0) {
    atomicAdd(scal + 1, rr);
    atomicAdd(scal + 4, bb);
    atomicAdd(scal + 0, rdr);
  }
}

// x += alpha p; r -= alpha q; r'r; r_c += P^T r; and the Jacobi part of r'z:  r'z = r'D^-1 r + r'(P z_c), where
// r'(P z_c) = (P^T r)'z_c = r_c'z_c is added by the coarse GEMV - no separate pass over the vectors for r'z.
__global__ void __launch_bounds__(256) tl_update_xr_kernel(int64_t n_n, CoarseGrid g, const double2* __restrict__ p, const double2* __restrict__ q,
                                                           const double2* __restrict__ minv, const double* __restrict__ coord,
                                                           double2* __restrict__ x, double2* __restrict__ r, double* rc, double* scal, int it) {
  __shared__ double red[32];
  const double rz_old = scal[rz_old_slot2(it)], pq = scal[3];
  const double alpha = (pq != 0.0) ? rz_old / pq : 0.0;
  double rr = 0.0, rdr = 0.0;
  const int64_t n_pad = (n_n + 31) & ~(int64_t)31;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_pad; a += (int64_t)gridDim.x * blockDim.x) {
    const bool valid = a < n_n;
    double2 ri = make_double2(0.0, 0.0);
    double cx = 0.0, cy = 0.0;
    if (valid) {
      const double2 pi = p[a], qi = __ldcs(q + a);
      double2 xi = x[a];
      ri = r[a];
      xi.x = fma(alpha, pi.x, xi.x);
      xi.y = fma(alpha, pi.y, xi.y);
      ri.x = fma(-alpha, qi.x, ri.x);
      ri.y = fma(-alpha, qi.y, ri.y);
      x[a] = xi;
      r[a] = ri;
      rr = fma(ri.x, ri.x, fma(ri.y, ri.y, rr));
      const double2 mi = minv[a];
      rdr = fma(ri.x * mi.x, ri.x, fma(ri.y * mi.y, ri.y, rdr));
      cx = coord[a];
      cy = coord[n_n + a];
    }
    restrict_warp(g, valid, cx, cy, ri.x, ri.y, rc);
  }
  rr = block_sum(rr, red);
  rdr = block_sum(rdr, red);
  if (threadIdx.x == 0) {
    atomicAdd(scal + 1, rr);
    atomicAdd(scal + rz_new_slot2(it), rdr);
  }
}

// z = minv r + P z_c;  MODE 0: r'z into scal[slot];  MODE 1: p = z (first direction);  MODE 2: p = z + beta p
template <int MODE>
__global__ void __launch_bounds__(256) tl_z_kernel(int64_t n_n, CoarseGrid g, const double2* __restrict__ r, const double2* __restrict__ minv,
                                                   const uint8_t* __restrict__ mask, const double* __restrict__ coord, const double* __restrict__ zc,
                                                   double2* __restrict__ p, double* scal, int slot, int it) {
  __shared__ double red[32];
  double beta = 0.0;
  if (MODE == 2) {
    const double rz_old = scal[rz_old_slot2(it)], rz_new = scal[rz_new_slot2(it)];
    beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
  }
  double rz = 0.0;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_n; a += (int64_t)gridDim.x * blockDim.x) {
    const double2 ri = r[a], mi = minv[a];
    double2 z = prolong(g, coord[a], coord[n_n + a], zc);
    if (mask) {
      const uchar2 mk = reinterpret_cast<const uchar2*>(mask)[a];
      if (!mk.x) z.x = 0.0;
      if (!mk.y) z.y = 0.0;
    }
    z.x = fma(mi.x, ri.x, z.x);
    z.y = fma(mi.y, ri.y, z.y);
    if (MODE == 0) {
      rz = fma(ri.x, z.x, fma(ri.y, z.y, rz));
    } else if (MODE == 1) {
      p[a] = z;
    } else {
      double2 pi = p[a];
      pi.x = fma(beta, pi.x, z.x);
      pi.y = fma(beta, pi.y, z.y);
      p[a] = pi;
    }
  }
  if (MODE == 0) {
    rz = block_sum(rz, red);
    if (threadIdx.x == 0) atomicAdd(scal + slot, rz);
  }
  if (MODE == 2 && blockIdx.x == 0 && threadIdx.x == 0) scal[3] = 0.0;  // p'Kp is re-accumulated by the next SpMV
}

// y = A x, dense row-major n x n (the inverted coarse operator): one warp per row, coalesced double2 loads
__global__ void __launch_bounds__(256) dense_gemv_kernel(int n, const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y,
                                                         double* dot) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double* row = A + (int64_t)warp * n;
  double acc = 0.0;
  if ((n & 1) == 0) {
    const double2* r2 = reinterpret_cast<const double2*>(row);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    for (int j = lane; j < n / 2; j += 32) {
      const double2 a = __ldcs(r2 + j), b = x2[j];
      acc = fma(a.x, b.x, fma(a.y, b.y, acc));
    }
  } else {
    for (int j = lane; j < n; j += 32) acc = fma(__ldcs(row + j), x[j], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    y[warp] = acc;
    if (dot) atomicAdd(dot, x[warp] * acc);  // x'Ax: the coarse part r_c'z_c of r'z
  }
}

static unsigned tl_grid(int64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t b = (n + 255) / 256;
  if (b > (int64_t)sms * 8) b = (int64_t)sms * 8;
  return (unsigned)(b < 1 ? 1 : b);
}

static int make_grid(CoarseGrid& g, double x0, double y0, double hx, double hy, int ncx, int ncy) {
  FEM_REQUIRE(hx > 0.0 && hy > 0.0 && ncx >= 1 && ncy >= 1, "coarse grid");
  g.x0 = x0; g.y0 = y0; g.inv_hx = 1.0 / hx; g.inv_hy = 1.0 / hy; g.ncx = ncx; g.ncy = ncy;
  return FEM_OK;
}

extern "C" int fem_coarse_galerkin(const fem_plan* P, const double* K_vals, const uint8_t* row_mask, const uint8_t* col_mask,
                                   const double* coord, double x0, double y0, double hx, double hy, int ncx, int ncy, double* Ac,
                                   fem_stream stream) {
  FEM_REQUIRE(P && K_vals && coord && Ac, "null pointer");
  CoarseGrid g;
  if (int rc = make_grid(g, x0, y0, hx, hy, ncx, ncy)) return rc;
  const int ncd = 2 * (ncx + 1) * (ncy + 1);
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemsetAsync(Ac, 0, sizeof(double) * (size_t)ncd * ncd, st));
  coarse_galerkin_kernel<<<tl_grid(P->n_n) * 4, 256, 0, st>>>(P->n_n, g, P->nbr_ptr, P->nbr_idx, K_vals, row_mask,
                                                              col_mask ? col_mask : row_mask, coord, Ac, ncd);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_dense_gemv(int n, const double* A, const double* x, double* y, double* dot, fem_stream stream) {
  FEM_REQUIRE(n > 0 && A && x && y, "null pointer");
  dense_gemv_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, A, x, y, dot);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_tl_init(int64_t n_n, const double* rhs, const double* Kx0, const uint8_t* free_mask, const double* minv, const double* coord,
                           double x0, double y0, double hx, double hy, int ncx, int ncy, double* r, double* rc, double* scal,
                           fem_stream stream) {
  FEM_REQUIRE(rhs && minv && coord && r && rc && scal && n_n > 0, "null pointer");
  CoarseGrid g;
  if (int rcode = make_grid(g, x0, y0, hx, hy, ncx, ncy)) return rcode;
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemsetAsync(scal, 0, 8 * sizeof(double), st));
  FEM_CUDA_CHECK(cudaMemsetAsync(rc, 0, sizeof(double) * 2 * (ncx + 1) * (ncy + 1), st));
  tl_init_kernel<<<tl_grid(n_n), 256, 0, st>>>(n_n, g, reinterpret_cast<const double2*>(rhs), reinterpret_cast<const double2*>(Kx0), free_mask,
                                               reinterpret_cast<const double2*>(minv), coord, reinterpret_cast<double2*>(r), rc, scal);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_tl_update_xr(int64_t n_n, const double* p, const double* q, const double* minv, const double* coord, double x0, double y0,
                                double hx, double hy, int ncx, int ncy, double* x, double* r, double* rc, double* scal, int iter,
                                fem_stream stream) {
  FEM_REQUIRE(p && q && minv && coord && x && r && rc && scal && n_n > 0, "null pointer");
  CoarseGrid g;
  if (int rcode = make_grid(g, x0, y0, hx, hy, ncx, ncy)) return rcode;
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemsetAsync(rc, 0, sizeof(double) * 2 * (ncx + 1) * (ncy + 1), st));
  tl_update_xr_kernel<<<tl_grid(n_n), 256, 0, st>>>(n_n, g, reinterpret_cast<const double2*>(p), reinterpret_cast<const double2*>(q),
                                                    reinterpret_cast<const double2*>(minv), coord, reinterpret_cast<double2*>(x),
                                                    reinterpret_cast<double2*>(r), rc, scal, iter);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

// mode 0: scal[slot] += r'z;  mode 1: p = z;  mode 2: p = z + beta p  (z = minv r + P z_c)
extern "C" int fem_tl_apply(int64_t n_n, int mode, const double* r, const double* minv, const uint8_t* free_mask, const double* coord,
                            double x0, double y0, double hx, double hy, int ncx, int ncy, const double* zc, double* p, double* scal,
                            int slot, int iter, fem_stream stream) {
  FEM_REQUIRE(r && minv && coord && zc && scal && n_n > 0 && mode >= 0 && mode <= 2 && slot >= 0 && slot < 8, "argument");
  FEM_REQUIRE(mode == 0 || p != nullptr, "p");
  CoarseGrid g;
  if (int rcode = make_grid(g, x0, y0, hx, hy, ncx, ncy)) return rcode;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = tl_grid(n_n);
  const double2 *r2 = reinterpret_cast<const double2*>(r), *m2 = reinterpret_cast<const double2*>(minv);
  double2* p2 = reinterpret_cast<double2*>(p);
  if (mode == 0) tl_z_kernel<0><<<blocks, 256, 0, st>>>(n_n, g, r2, m2, free_mask, coord, zc, p2, scal, slot, iter);
  else if (mode == 1) tl_z_kernel<1><<<blocks, 256, 0, st>>>(n_n, g, r2, m2, free_mask, coord, zc, p2, scal, slot, iter);
  else tl_z_kernel<2><<<blocks, 256, 0, st>>>(n_n, g, r2, m2, free_mask, coord, zc, p2, scal, slot, iter);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}
