// K7-K9: CSR SpMV on the 2x2 node-block pattern, masked Jacobi-PCG with device-resident scalars.
// Replaces the dense solve of the reference's Newton loop (Plasticity2D_DP/pythonFEM.py:1062-1066;
// the boolean-mask extraction K[Q,Q] becomes a projected CG on the full vectors) and provides the
// energy products of its stopping criterion (:1072-1075).
//
// SpMV: GROUP lanes cooperate on one node (two rows).  The pattern is stored once per 2x2 block
// (nbr_idx, 4 B per 4 values) instead of once per value, so a sweep moves 8 B/nnz of values plus
// 1 B/nnz of indices instead of CSR's 12 B/nnz.  Row values are read as coalesced double2, x is
// gathered as one double2 per block through L1/L2.
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "spmv.cuh"
#include "spmv_stream.cuh"

// Partial-sum buffers of the order-deterministic reductions of the PCG vector kernels (common.cuh: ordered_accumulate), one per
// scalar array the caller passes (fem_pcg_*'s `scal`): created on first use - fem_pcg_init, always outside a stream capture -
// and kept for the life of the process (196 KB each; a solver object has one scalar array).
FemRedBuf fem_red_buffer_for(const void* key, cudaStream_t st) {
  static std::mutex mu;
  static std::unordered_map<uintptr_t, FemRedBuf> map;
  std::lock_guard<std::mutex> lock(mu);
  const auto it = map.find(reinterpret_cast<uintptr_t>(key));
  if (it != map.end()) return it->second;
  FemRedBuf rb{nullptr, nullptr};
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return rb;  // no allocation inside a capture: atomics
  double* p = nullptr;
  if (cudaMalloc(&p, (3 * FEM_RED_MAXB + 8) * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    return rb;
  }
  rb.partials = p;
  rb.ticket = reinterpret_cast<unsigned*>(p + 3 * FEM_RED_MAXB);
  cudaMemsetAsync(rb.ticket, 0, 8 * sizeof(double), st);
  map.emplace(reinterpret_cast<uintptr_t>(key), rb);
  return rb;
}

template <int GROUP, int U>
__global__ void __launch_bounds__(256) spmv_blocks_kernel(int64_t n_n, const int32_t* __restrict__ nbr_ptr,
                                                          const int32_t* __restrict__ nbr_idx,
                                                          const double* __restrict__ vals, const double* __restrict__ x,
                                                          double* __restrict__ y, const uint8_t* __restrict__ mask,
                                                          double* dot_out, double* zero_a, double* zero_b, const FemRedBuf rb) {
  __shared__ double red[32];
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // scalar housekeeping for the PCG (see fem_pcg_spmv_dot)
    if (zero_a) *zero_a = 0.0;
    if (zero_b) *zero_b = 0.0;
  }
  double dot = spmv_rows<GROUP, U>(n_n, nbr_ptr, nbr_idx, vals, x, y, mask, dot_out != nullptr);
  if (dot_out) {
    const double v[1] = {block_sum(dot, red)};
    double* const dst[1] = {dot_out};
    ordered_accumulate<1>(v, dst, rb);
  }
}

// x staged through shared memory by bulk async copies (spmv.cuh: spmv_tiles)
template <int GROUP>
__global__ void __launch_bounds__(FEM_SPMV_THREADS) spmv_tiles_kernel(int64_t n_n, int64_t n_tiles, const int32_t* __restrict__ nbr_ptr,
                                                         const int32_t* __restrict__ nbr_idx, const uint16_t* __restrict__ nbr_loc,
                                                         const int32_t* __restrict__ tile_seg, const double* __restrict__ vals,
                                                         const double* __restrict__ x, double* __restrict__ y,
                                                         const uint8_t* __restrict__ mask, double* dot_out, double* zero_a, double* zero_b,
                                                         const FemRedBuf rb) {
  __shared__ double red[32];
  __shared__ SpmvTileSmem sm;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (zero_a) *zero_a = 0.0;
    if (zero_b) *zero_b = 0.0;
  }
  double dot = spmv_tiles<GROUP, false>(n_n, n_tiles, nbr_ptr, nbr_idx, nbr_loc, tile_seg, vals, x, y, mask, dot_out != nullptr, sm);
  if (dot_out) {
    const double v[1] = {block_sum(dot, red)};
    double* const dst[1] = {dot_out};
    ordered_accumulate<1>(v, dst, rb);
  }
}

// every operand streamed through shared memory by bulk async copies (spmv_stream.cuh); one persistent CTA per SM
template <int GROUP>
__global__ void __launch_bounds__(FEM_STREAM_THREADS, 1) spmv_stream_kernel(const SpmvStreamArgs A, const double* __restrict__ vals, const double* __restrict__ x,
                                                                        const SpmvStreamStore epi, double* dot_out, double* zero_a, double* zero_b,
                                                                        const FemRedBuf rb) {
  extern __shared__ __align__(128) unsigned char stream_smem[];
  __shared__ double red[32];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (zero_a) *zero_a = 0.0;
    if (zero_b) *zero_b = 0.0;
  }
  double dot = spmv_stream<GROUP, false, double>(A, vals, x, epi, stream_smem);
  if (dot_out) {
    const double v[1] = {block_sum(dot, red)};
    double* const dst[1] = {dot_out};
    ordered_accumulate<1>(v, dst, rb);
  }
}

static int launch_spmv(const fem_plan* P, const double* K_vals, const double* x, double* y, const uint8_t* mask,
                       double* dot, double* zero_a, double* zero_b, cudaStream_t st) {
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(K_vals) & 15u) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(y) & 15u) == 0, "K_vals, x, y must be 16-byte aligned");
  const SpmvShape sh = spmv_shape(P);
  const int threads = 256, group = sh.group, unroll = sh.unroll;
  const FemRedBuf rb{P->red_partials, P->red_ticket};  // order-deterministic x'y (one SpMV with a dot product per plan at a time)
  // default: every operand streamed through shared memory by a producer warp (0.380 ms at 16M elements; the register-fed
  // kernel below: 0.455 ms; tuning key spmv_staged = 2 selects it, 1 the round-1 gather kernel)
  if (g_fem_tuning.spmv_staged == 0 && spmv_can_stream(P) && (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 15u) == 0)) {
    const SpmvStreamArgs A{P->n_n, P->n_tiles, P->nbr_ptr, P->nbr_idx, P->nbr_loc, P->tile_seg};
    const SpmvStreamStore epi{y, mask, x, dot != nullptr};
    constexpr int smem = SpmvStreamSmem<double>::TOTAL;
    const unsigned sb = spmv_stream_blocks(P);
#define SPMVS(G)                                                                                                  \
  do {                                                                                                            \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      FEM_CUDA_CHECK(cudaFuncSetAttribute(spmv_stream_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    spmv_stream_kernel<G><<<sb, FEM_STREAM_THREADS, smem, st>>>(A, K_vals, x, epi, dot, zero_a, zero_b, rb);        \
  } while (0)
    if (group == 4) SPMVS(4);
    else if (group == 8) SPMVS(8);
    else SPMVS(16);
#undef SPMVS
    FEM_CUDA_CHECK(cudaGetLastError());
    return FEM_OK;
  }
  if (spmv_use_tiles(P)) {
    const unsigned tb = spmv_tile_blocks(P);
#define SPMVT(G) spmv_tiles_kernel<G><<<tb, FEM_SPMV_THREADS, 0, st>>>(P->n_n, P->n_tiles, P->nbr_ptr, P->nbr_idx, P->nbr_loc, P->tile_seg, K_vals, x, y, mask, dot, zero_a, zero_b, rb)
    if (group == 4) SPMVT(4);
    else if (group == 8) SPMVT(8);
    else SPMVT(16);
#undef SPMVT
    FEM_CUDA_CHECK(cudaGetLastError());
    return FEM_OK;
  }
  const unsigned blocks = sh.blocks;
#define SPMV(G, UU) spmv_blocks_kernel<G, UU><<<blocks, threads, 0, st>>>(P->n_n, P->nbr_ptr, P->nbr_idx, K_vals, x, y, mask, dot, zero_a, zero_b, rb)
#define SPMV_U(G) do { if (unroll == 1) SPMV(G, 1); else if (unroll == 2) SPMV(G, 2); else SPMV(G, 4); } while (0)
  if (group == 4) SPMV_U(4);
  else if (group == 8) SPMV_U(8);
  else SPMV_U(16);
#undef SPMV_U
#undef SPMV
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_spmv(const fem_plan* P, const double* K_vals, const double* x, double* y, const uint8_t* free_mask,
                        double* dot, fem_stream stream) {
  FEM_REQUIRE(P && K_vals && x && y, "null pointer");
  return launch_spmv(P, K_vals, x, y, free_mask, dot, nullptr, nullptr, (cudaStream_t)stream);
}

// ---- Jacobi ------------------------------------------------------------------------------------
__global__ void jacobi_kernel(int64_t n_n, const int32_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr_idx,
                              const double* __restrict__ vals, const uint8_t* __restrict__ mask, double* __restrict__ minv) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_n; a += (int64_t)gridDim.x * blockDim.x) {
    const int p0 = nbr_ptr[a], deg = nbr_ptr[a + 1] - p0;
    double d0 = 0.0, d1 = 0.0;
    for (int j = 0; j < deg; ++j)
      if (nbr_idx[p0 + j] == a) {
        d0 = vals[4 * (int64_t)p0 + 2 * j];
        d1 = vals[4 * (int64_t)p0 + 2 * deg + 2 * j + 1];
        break;
      }
    const bool f0 = mask ? mask[2 * a] != 0 : true, f1 = mask ? mask[2 * a + 1] != 0 : true;
    minv[2 * a] = (f0 && d0 != 0.0) ? 1.0 / d0 : 0.0;
    minv[2 * a + 1] = (f1 && d1 != 0.0) ? 1.0 / d1 : 0.0;
  }
}

extern "C" int fem_jacobi_setup(const fem_plan* P, const double* K_vals, const uint8_t* free_mask, double* minv,
                                fem_stream stream) {
  FEM_REQUIRE(P && K_vals && minv, "null pointer");
  const int threads = 256;
  jacobi_kernel<<<(unsigned)fem_div_up(P->n_n, threads), threads, 0, (cudaStream_t)stream>>>(P->n_n, P->nbr_ptr, P->nbr_idx,
                                                                                            K_vals, free_mask, minv);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

// ---- PCG vector kernels ------------------------------------------------------------------------
// scal[0] = r'z (even iterations), scal[1] = r'r, scal[2] = r'z (odd iterations), scal[3] = p'Kp, scal[4] = |b|^2.
// Iteration `it` reads rz_old = scal[(it&1)?2:0] and accumulates rz_new into scal[(it&1)?0:2]; the slot pair
// {rz_new, r'r} is contiguous either way, so a multi-GPU driver all-reduces two adjacent doubles.
__device__ __forceinline__ int rz_old_slot(int it) { return (it & 1) ? 2 : 0; }
__device__ __forceinline__ int rz_new_slot(int it) { return (it & 1) ? 0 : 2; }

__global__ void __launch_bounds__(256) pcg_init_kernel(int64_t n, const double* __restrict__ rhs, const double* __restrict__ Kx0,
                                                       const uint8_t* __restrict__ mask, const double* __restrict__ minv,
                                                       double* __restrict__ r, double* __restrict__ p, double* scal, const FemRedBuf rb) {
  __shared__ double red[32];
  double rz = 0.0, rr = 0.0, bb = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool fr = mask ? mask[i] != 0 : true;
    const double b = fr ? rhs[i] : 0.0;
    const double ri = fr ? (Kx0 ? b - Kx0[i] : b) : 0.0;
    const double z = minv[i] * ri;
    r[i] = ri;
    p[i] = z;
    rz = fma(ri, z, rz);
    rr = fma(ri, ri, rr);
    bb = fma(b, b, bb);
  }
  const double v[3] = {block_sum(rz, red), block_sum(rr, red), block_sum(bb, red)};
  double* const dst[3] = {scal + 0, scal + 1, scal + 4};
  ordered_accumulate<3>(v, dst, rb);
}

__global__ void __launch_bounds__(256) pcg_update_xr_kernel(int64_t n2, const double2* __restrict__ p, const double2* __restrict__ q,
                                                            const double2* __restrict__ minv, double2* __restrict__ x,
                                                            double2* __restrict__ r, double* scal, int it, const FemRedBuf rb) {
  __shared__ double red[32];
  const double rz_old = scal[rz_old_slot(it)], pq = scal[3];
  const double alpha = (pq != 0.0) ? rz_old / pq : 0.0;
  double rz = 0.0, rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 pi = p[i], qi = __ldcs(q + i), mi = minv[i];
    double2 xi = x[i], ri = r[i];
    xi.x = fma(alpha, pi.x, xi.x);
    xi.y = fma(alpha, pi.y, xi.y);
    ri.x = fma(-alpha, qi.x, ri.x);
    ri.y = fma(-alpha, qi.y, ri.y);
    x[i] = xi;
    r[i] = ri;
    rz = fma(ri.x * mi.x, ri.x, rz);
    rz = fma(ri.y * mi.y, ri.y, rz);
    rr = fma(ri.x, ri.x, rr);
    rr = fma(ri.y, ri.y, rr);
  }
  const double v[2] = {block_sum(rz, red), block_sum(rr, red)};
  double* const dst[2] = {scal + rz_new_slot(it), scal + 1};
  ordered_accumulate<2>(v, dst, rb);
}

__global__ void __launch_bounds__(256) pcg_update_p_kernel(int64_t n2, const double2* __restrict__ r, const double2* __restrict__ minv,
                                                           double2* __restrict__ p, double* scal, int it) {
  const double rz_old = scal[rz_old_slot(it)], rz_new = scal[rz_new_slot(it)];
  const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 ri = r[i], mi = minv[i];
    double2 pi = p[i];
    pi.x = fma(beta, pi.x, mi.x * ri.x);
    pi.y = fma(beta, pi.y, mi.y * ri.y);
    p[i] = pi;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[3] = 0.0;  // p'Kp is re-accumulated by the next SpMV
}

// p-update fused with the halo push: the rows other ranks need are stored to their ghost rows (NVLink peer memory)
// by the thread that computes them.  Offsets/counts are in doubles and even (whole nodes).
__global__ void __launch_bounds__(256) pcg_update_p_push_kernel(int64_t n2, const double2* __restrict__ r, const double2* __restrict__ minv,
                                                                double2* __restrict__ p, double* scal, int it, int64_t s0, int64_t c0,
                                                                double2* dst0, int64_t s1, int64_t c1, double2* dst1, int64_t own_lo) {
  const double rz_old = scal[rz_old_slot(it)], rz_new = scal[rz_new_slot(it)];
  const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
  // only the owned range [own_lo, n2) is updated: this rank's ghost rows are written by their owners' pushes, and a local
  // write there would race with them
  for (int64_t i = own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 ri = r[i], mi = minv[i];
    double2 pi = p[i];
    pi.x = fma(beta, pi.x, mi.x * ri.x);
    pi.y = fma(beta, pi.y, mi.y * ri.y);
    p[i] = pi;
    if (dst0 && i >= s0 && i < s0 + c0) dst0[i - s0] = pi;
    if (dst1 && i >= s1 && i < s1 + c1) dst1[i - s1] = pi;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[3] = 0.0;
}

__global__ void halo_push_kernel(const double2* __restrict__ v, int64_t s0, int64_t c0, double2* dst0, int64_t s1, int64_t c1,
                                 double2* dst1) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c0 + c1; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < c0) { if (dst0) dst0[i] = v[s0 + i]; }
    else if (dst1) dst1[i - c0] = v[s1 + (i - c0)];
  }
}

static unsigned vec_grid(const int64_t n_items, const int sm_count) {
  int64_t b = (n_items + 255) / 256;
  const int64_t cap = (int64_t)(sm_count > 0 ? sm_count : 148) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}
static int sm_count_now() {  // cached per device: one attribute query per device and process, not per launch
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  int sms = cache[dev];
  if (sms == 0) {
    sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = sms;
  }
  return sms;
}

extern "C" int fem_pcg_init(int64_t n, const double* rhs, const double* Kx0, const uint8_t* free_mask, const double* minv,
                            double* r, double* p, double* scal, fem_stream stream) {
  FEM_REQUIRE(rhs && minv && r && p && scal && n > 0, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemsetAsync(scal, 0, 8 * sizeof(double), st));
  pcg_init_kernel<<<vec_grid(n, sm_count_now()), 256, 0, st>>>(n, rhs, Kx0, free_mask, minv, r, p, scal, fem_red_buffer_for(scal, st));
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_pcg_spmv_dot(const fem_plan* P, const double* K_vals, const double* p, double* q,
                                const uint8_t* free_mask, double* scal, int iter, fem_stream stream) {
  FEM_REQUIRE(P && K_vals && p && q && scal, "null pointer");
  // zero the slots the following update_xr accumulates into: rz_new(iter) and r'r
  double* rz_new = scal + ((iter & 1) ? 0 : 2);
  return launch_spmv(P, K_vals, p, q, free_mask, scal + 3, rz_new, scal + 1, (cudaStream_t)stream);
}

extern "C" int fem_pcg_update_xr(int64_t n, const double* p, const double* q, const double* minv, double* x, double* r,
                                 double* scal, int iter, fem_stream stream) {
  FEM_REQUIRE(p && q && minv && x && r && scal && n > 0 && n % 2 == 0, "null pointer or odd n");
  pcg_update_xr_kernel<<<vec_grid(n / 2, sm_count_now()), 256, 0, (cudaStream_t)stream>>>(
      n / 2, reinterpret_cast<const double2*>(p), reinterpret_cast<const double2*>(q), reinterpret_cast<const double2*>(minv),
      reinterpret_cast<double2*>(x), reinterpret_cast<double2*>(r), scal, iter, fem_red_buffer_for(scal, (cudaStream_t)stream));
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_pcg_update_p(int64_t n, const double* r, const double* minv, double* p, double* scal, int iter,
                                fem_stream stream) {
  FEM_REQUIRE(r && minv && p && scal && n > 0 && n % 2 == 0, "null pointer or odd n");
  pcg_update_p_kernel<<<vec_grid(n / 2, sm_count_now()), 256, 0, (cudaStream_t)stream>>>(
      n / 2, reinterpret_cast<const double2*>(r), reinterpret_cast<const double2*>(minv), reinterpret_cast<double2*>(p), scal, iter);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_pcg(const fem_plan* P, const double* K_vals, const double* rhs, const uint8_t* free_mask, double rtol,
                       int maxit, int check_every, double* x, double* work, int* h_iters, double* h_relres,
                       fem_stream stream) {
  FEM_REQUIRE(P && K_vals && rhs && x && work, "null pointer");
  FEM_REQUIRE(maxit >= 0 && rtol >= 0.0, "maxit/rtol");
  if (check_every < 1) check_every = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = P->n_dof;
  double *r = work, *p = work + n, *q = work + 2 * n, *minv = work + 3 * n;
  double* scal = P->dscratch;
  int rc;
  if ((rc = fem_jacobi_setup(P, K_vals, free_mask, minv, stream)) != FEM_OK) return rc;
  if ((rc = launch_spmv(P, K_vals, x, q, free_mask, nullptr, nullptr, nullptr, st)) != FEM_OK) return rc;  // K x0
  if ((rc = fem_pcg_init(n, rhs, q, free_mask, minv, r, p, scal, stream)) != FEM_OK) return rc;
  double h[8];
  FEM_CUDA_CHECK(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st));
  FEM_CUDA_CHECK(cudaStreamSynchronize(st));
  const double bnorm2 = h[4];
  const double target2 = rtol * rtol * bnorm2;
  int it = 0;
  double rr = h[1];
  int status = FEM_OK;
  if (bnorm2 == 0.0 || rr <= target2) {
    if (h_iters) *h_iters = 0;
    if (h_relres) *h_relres = bnorm2 > 0.0 ? sqrt(rr / bnorm2) : 0.0;
    return FEM_OK;
  }
  while (it < maxit) {
    int chunk = check_every < (maxit - it) ? check_every : (maxit - it);
    for (int k = 0; k < chunk; ++k, ++it) {
      if ((rc = fem_pcg_spmv_dot(P, K_vals, p, q, free_mask, scal, it, stream)) != FEM_OK) return rc;
      if ((rc = fem_pcg_update_xr(n, p, q, minv, x, r, scal, it, stream)) != FEM_OK) return rc;
      if ((rc = fem_pcg_update_p(n, r, minv, p, scal, it, stream)) != FEM_OK) return rc;
    }
    FEM_CUDA_CHECK(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st));
    FEM_CUDA_CHECK(cudaStreamSynchronize(st));
    rr = h[1];
    if (!(rr == rr) || isinf(rr)) {
      fem_set_error("PCG breakdown: residual norm is not finite after %d iterations", it);
      status = FEM_ERR_PCG_BREAKDOWN;
      break;
    }
    if (rr <= target2) break;
  }
  if (status == FEM_OK && rr > target2) {
    fem_set_error("PCG did not reach rtol=%g in %d iterations (relres=%g)", rtol, it, sqrt(rr / bnorm2));
    status = FEM_ERR_PCG_MAXIT;
  }
  if (h_iters) *h_iters = it;
  if (h_relres) *h_relres = sqrt(rr / bnorm2);
  return status;
}

extern "C" int fem_energy_norms(const fem_plan* P, const double* K_vals, const double* v0, const double* v1,
                                const double* v2, double* work, double* out, fem_stream stream) {
  FEM_REQUIRE(P && K_vals && v0 && v1 && v2 && work && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FEM_CUDA_CHECK(cudaMemsetAsync(out, 0, 3 * sizeof(double), st));
  const double* v[3] = {v0, v1, v2};
  for (int i = 0; i < 3; ++i) {
    const int rc = launch_spmv(P, K_vals, v[i], work, nullptr, out + i, nullptr, nullptr, st);
    if (rc != FEM_OK) return rc;
  }
  return FEM_OK;
}

__global__ void axpby_kernel(int64_t n, double a, const double* x, double b, const double* y, double* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * x[i] + b * y[i];
}

extern "C" int fem_vec_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out, fem_stream stream) {
  FEM_REQUIRE(x && y && out && n >= 0, "null pointer");
  if (n == 0) return FEM_OK;
  axpby_kernel<<<vec_grid(n, sm_count_now()) * 4, 256, 0, (cudaStream_t)stream>>>(n, a, x, b, y, out);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_halo_push(const double* v, int64_t src0, int64_t n0, double* dst0, int64_t src1, int64_t n1, double* dst1,
                             fem_stream stream) {
  FEM_REQUIRE(v && src0 >= 0 && src1 >= 0 && n0 >= 0 && n1 >= 0 && src0 % 2 == 0 && src1 % 2 == 0 && n0 % 2 == 0 && n1 % 2 == 0,
              "halo ranges must be whole nodes");
  if ((!dst0 || n0 == 0) && (!dst1 || n1 == 0)) return FEM_OK;
  const int64_t tot = (n0 + n1) / 2;
  halo_push_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(v), src0 / 2, n0 / 2, reinterpret_cast<double2*>(dst0), src1 / 2, n1 / 2,
      reinterpret_cast<double2*>(dst1));
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_pcg_update_p_push(int64_t own_lo, int64_t own_hi, const double* r, const double* minv, double* p, double* scal,
                                     int iter, int64_t src0, int64_t n0, double* dst0, int64_t src1, int64_t n1, double* dst1,
                                     fem_stream stream) {
  FEM_REQUIRE(r && minv && p && scal && own_lo >= 0 && own_hi > own_lo && own_lo % 2 == 0 && own_hi % 2 == 0, "null pointer or bad owned range");
  FEM_REQUIRE(src0 % 2 == 0 && src1 % 2 == 0 && n0 % 2 == 0 && n1 % 2 == 0, "halo ranges must be whole nodes");
  pcg_update_p_push_kernel<<<vec_grid((own_hi - own_lo) / 2, sm_count_now()), 256, 0, (cudaStream_t)stream>>>(
      own_hi / 2, reinterpret_cast<const double2*>(r), reinterpret_cast<const double2*>(minv), reinterpret_cast<double2*>(p), scal, iter,
      src0 / 2, n0 / 2, reinterpret_cast<double2*>(dst0), src1 / 2, n1 / 2, reinterpret_cast<double2*>(dst1), own_lo / 2);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}
