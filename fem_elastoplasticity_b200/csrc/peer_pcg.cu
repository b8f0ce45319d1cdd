// Multi-GPU Jacobi-PCG iteration with its exchange steps fused into the three kernels (one process per GPU, strip
// partition).  The reference has no distributed code; this is the solve of its Newton loop
// (Plasticity2D_DP/pythonFEM.py:1062-1066) sharded over GPUs.
//
// Every rank owns a small "communication block" in symmetric memory (CUDA IPC mappings, reachable by every peer through
// NVLink/NVSwitch).  A PCG iteration needs three exchanges, and each one leaves from the kernel that produces the data
// and is awaited by the kernel that consumes it - no collective library call, no host round trip, so the whole
// iteration is three plain kernel launches and can be replayed from a CUDA graph:
//
//   A  q = K p, p'q      waits: neighbours' ghost rows of p (halo flags)      publishes: partial p'q  -> every rank
//   B  x, r update       waits: p'q partials of all ranks                     publishes: partial {r'z, r'r} -> every rank
//   C  p = z + beta p    waits: {r'z, r'r} partials of all ranks              publishes: interface rows of p -> neighbours
//
// The scalar partials travel as self-validating 16-byte lines {value.lo, seq, value.hi, seq} (one vector store; the
// receiver polls its own copy until both sequence words match), so neither side needs a system-scope fence.  The
// interface rows of p are bulk data: plain peer stores, fence, then a release store of the sequence number to the
// neighbour's flag word - issued by the few blocks that own those rows, which run FIRST in kernel C, so the flag is
// long there when the neighbour's next SpMV starts.  Partials are summed by every rank in rank order, so all ranks
// compute bit-identical alpha and beta.  The sequence number is a device-resident iteration counter that only ever
// grows, so a graph replay needs no changing arguments.  Waits are bounded (10 s, "peer_timeout_ms"): on a time-out the kernel sets an
// error word and carries on, so a dead peer cannot hang the GPU.
//
// Why overwriting is safe: rank X publishes p'q(it) at the end of A(it), i.e. after it passed the wait of C(it-1), which
// needs every rank's B(it-1) to have ended - and B(it-1) was the last reader of the p'q(it-1) lines.  The same argument
// covers the {r'z, r'r} lines (double-buffered by iteration parity because C(it) still needs r'z(it-1)) and the ghost rows
// of p (written in C(it), last read by the neighbours' A(it), which ended before their B(it) published).
#include "common.cuh"
#include "spmv.cuh"
#include "peer.cuh"

#define FEM_PEER_MAX 16
// layout of a communication block, in 8-byte words (FEM_PPCG_WORD_* in the header mirror ERR and OUT)
enum {
  PW_LL_PQ = 0,       // [16][2]    line of rank r's p'q partial
  PW_LL_RZ = 32,      // [2][16][4] lines of rank r's {r'z, r'r} partials, indexed by iteration parity
  PW_HFLAG = 160,     // [2]  ghost rows of p current: [0] written by the lower neighbour, [1] by the upper one
  // words below are only touched by the owning rank
  PW_IT = 164,        // iteration counter (monotone over the life of the block)
  PW_TICKET = 165,    // last-block detection
  PW_TICKET_HALO = 166,  // ... among the blocks that push the interface rows
  PW_ERR = FEM_PPCG_WORD_ERR,  // sticky: a wait timed out
  PW_ACC = 168,       // [3] doubles: local p'q, r'z, r'r accumulators
  PW_OUT = FEM_PPCG_WORD_OUT,  // [2] doubles: global r'z and r'r of the last finished iteration (for the host's convergence check)
  PW_WORDS = FEM_PPCG_WORDS
};
static_assert(PW_ERR == 167 && PW_OUT == 172 && PW_WORDS >= 174, "communication block layout");

struct PeerView {
  uint64_t* local;
  uint64_t* peer[FEM_PEER_MAX];  // peer[r] = rank r's block (peer[rank] == local)
  int rank, world;
  int nowait;  // diagnostic: never spin (see FemTuning::peer_nowait)
  uint64_t timeout_ns;  // bound of every wait ("peer_timeout_ms" tuning key, default 10 s)
};

// ---- A: q = K p and the partial p'q ---------------------------------------------------------------------------------
// p is NOT const/__restrict__ and is never read through the non-coherent path: its ghost rows are stored by the
// neighbour GPUs while this kernel is already resident (it spins on the halo flags), so every load of p is a coherent
// one (bulk async copies after a proxy fence, or ld.global.cg in the gather fallback).
template <int GROUP>
__global__ void __launch_bounds__(FEM_SPMV_THREADS) ppcg_spmv_kernel(int64_t n_n, int64_t n_tiles, const int32_t* __restrict__ nbr_ptr,
                                                        const int32_t* __restrict__ nbr_idx, const uint16_t* __restrict__ nbr_loc,
                                                        const int32_t* __restrict__ tile_seg, const double* __restrict__ vals,
                                                        const double* p, double* __restrict__ q,
                                                        const uint8_t* __restrict__ mask, const PeerView pv) {
  __shared__ double red[32];
  __shared__ int sh_last;
  __shared__ SpmvTileSmem sm;
  uint64_t* L = pv.local;
  const uint64_t it = L[PW_IT];
  // ghost rows of p: stored by the neighbours' C(it-1); nothing of p is loaded before the flags are seen
  if (threadIdx.x == 0 && pv.rank > 0) wait_flag(L + PW_HFLAG + 0, it, L + PW_ERR, pv.nowait, pv.timeout_ns);
  if (threadIdx.x == 32 && pv.rank < pv.world - 1) wait_flag(L + PW_HFLAG + 1, it, L + PW_ERR, pv.nowait, pv.timeout_ns);  // another warp: both polls in flight
  __syncthreads();
  double dot = spmv_tiles<GROUP, true>(n_n, n_tiles, nbr_ptr, nbr_idx, nbr_loc, tile_seg, vals, p, q, mask, true, sm);
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) atomicAdd(reinterpret_cast<double*>(L + PW_ACC) + 0, dot);
  if (arrive_last(L + PW_TICKET, gridDim.x, &sh_last)) {
    if (threadIdx.x < pv.world) {
      const double tot = __ldcg(reinterpret_cast<const double*>(L + PW_ACC) + 0);
      line_store(pv.peer[threadIdx.x] + PW_LL_PQ + 2 * pv.rank, tot, (uint32_t)(it + 1));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      reinterpret_cast<double*>(L + PW_ACC)[0] = 0.0;
      *reinterpret_cast<unsigned*>(L + PW_TICKET) = 0u;
    }
  }
}

// ---- B: x += alpha p, r -= alpha q and the partials {r'z, r'r} -------------------------------------------------------
__global__ void __launch_bounds__(256) ppcg_update_xr_kernel(int64_t n2, const double2* __restrict__ p, const double2* __restrict__ q,
                                                             const double2* __restrict__ minv, double2* __restrict__ x,
                                                             double2* __restrict__ r, const PeerView pv) {
  __shared__ double red[32];
  __shared__ double sh_pq[FEM_PEER_MAX], sh_rz[FEM_PEER_MAX];
  __shared__ int sh_last;
  uint64_t* L = pv.local;
  const uint64_t it = L[PW_IT];
  {  // one warp per quantity, one lane per rank: all lines are polled concurrently
    const int w = threadIdx.x >> 5, k = threadIdx.x & 31;
    if (w == 0 && k < pv.world) sh_pq[k] = line_wait(L + PW_LL_PQ + 2 * k, (uint32_t)(it + 1), L + PW_ERR, pv.nowait, pv.timeout_ns);
    if (w == 1 && k < pv.world) sh_rz[k] = line_wait(L + PW_LL_RZ + ((it + 1) & 1) * 4 * FEM_PEER_MAX + 4 * k, (uint32_t)it, L + PW_ERR, pv.nowait, pv.timeout_ns);
  }
  __syncthreads();
  double pq = 0.0, rz_old = 0.0;
  for (int k = 0; k < pv.world; ++k) {  // rank order: the same sum on every rank
    pq += sh_pq[k];
    rz_old += sh_rz[k];
  }
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
  rz = block_sum(rz, red);
  rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    atomicAdd(reinterpret_cast<double*>(L + PW_ACC) + 1, rz);
    atomicAdd(reinterpret_cast<double*>(L + PW_ACC) + 2, rr);
  }
  if (arrive_last(L + PW_TICKET, gridDim.x, &sh_last)) {
    if (threadIdx.x < 2 * pv.world) {  // thread 2r: r'z to rank r, thread 2r+1: r'r to rank r
      const int dst_rank = threadIdx.x >> 1, which = threadIdx.x & 1;
      const double tot = __ldcg(reinterpret_cast<const double*>(L + PW_ACC) + 1 + which);
      line_store(pv.peer[dst_rank] + PW_LL_RZ + (it & 1) * 4 * FEM_PEER_MAX + 4 * pv.rank + 2 * which, tot, (uint32_t)(it + 1));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      reinterpret_cast<double*>(L + PW_ACC)[1] = 0.0;
      reinterpret_cast<double*>(L + PW_ACC)[2] = 0.0;
      *reinterpret_cast<unsigned*>(L + PW_TICKET) = 0u;
    }
  }
}

// ---- C: p = z + beta p on the owned rows, interface rows stored straight into the neighbours' ghost rows -------------
// The first n_halo_blocks blocks do the interface rows only (one item per thread): their flag leaves early and their
// system fence overlaps with the other blocks' work.
__global__ void __launch_bounds__(256) ppcg_update_p_kernel(int64_t own_lo, int64_t own_hi, const double2* __restrict__ r,
                                                            const double2* __restrict__ minv, double2* __restrict__ p,
                                                            int64_t s_up, int64_t c_up, double2* dst_up, int64_t s_lo, int64_t c_lo,
                                                            double2* dst_lo, unsigned n_halo_blocks, const PeerView pv) {
  __shared__ double sh_new[FEM_PEER_MAX], sh_rr[FEM_PEER_MAX], sh_old[FEM_PEER_MAX];
  __shared__ int sh_last;
  uint64_t* L = pv.local;
  const uint64_t it = L[PW_IT];
  {
    const int w = threadIdx.x >> 5, k = threadIdx.x & 31;
    const uint64_t* nw = L + PW_LL_RZ + (it & 1) * 4 * FEM_PEER_MAX + 4 * k;
    if (w == 0 && k < pv.world) sh_new[k] = line_wait(nw, (uint32_t)(it + 1), L + PW_ERR, pv.nowait, pv.timeout_ns);
    if (w == 1 && k < pv.world) sh_rr[k] = line_wait(nw + 2, (uint32_t)(it + 1), L + PW_ERR, pv.nowait, pv.timeout_ns);
    if (w == 2 && k < pv.world) sh_old[k] = line_wait(L + PW_LL_RZ + ((it + 1) & 1) * 4 * FEM_PEER_MAX + 4 * k, (uint32_t)it, L + PW_ERR, pv.nowait, pv.timeout_ns);
  }
  __syncthreads();
  double rz_new = 0.0, rz_old = 0.0, rr = 0.0;
  for (int k = 0; k < pv.world; ++k) {
    rz_new += sh_new[k];
    rz_old += sh_old[k];
    rr += sh_rr[k];
  }
  const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
  if (blockIdx.x < n_halo_blocks) {  // interface rows: item j < c_up goes up, the rest goes down
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < c_up + c_lo) {
      const bool up = j < c_up;
      const int64_t i = up ? s_up + j : s_lo + (j - c_up);
      const double2 ri = r[i], mi = minv[i];
      double2 pi = p[i];
      pi.x = fma(beta, pi.x, mi.x * ri.x);
      pi.y = fma(beta, pi.y, mi.y * ri.y);
      p[i] = pi;
      if (up) dst_up[j] = pi;
      else dst_lo[j - c_up] = pi;
      __threadfence_system();
    }
    if (arrive_last(L + PW_TICKET_HALO, n_halo_blocks, &sh_last)) {
      if (threadIdx.x == 0 && c_up > 0) st_release_sys(pv.peer[pv.rank + 1] + PW_HFLAG + 0, it + 1);
      if (threadIdx.x == 1 && c_lo > 0) st_release_sys(pv.peer[pv.rank - 1] + PW_HFLAG + 1, it + 1);
      if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(L + PW_TICKET_HALO) = 0u;
    }
  }
  // the rest of the owned range, on the other blocks (the interface blocks are left to their fence and flag); ghost
  // rows of p are never written locally (their owners store them)
  const int64_t n_rest = (int64_t)gridDim.x - n_halo_blocks;
  if (blockIdx.x >= n_halo_blocks)
    for (int64_t i = own_lo + ((int64_t)blockIdx.x - n_halo_blocks) * blockDim.x + threadIdx.x; i < own_hi; i += n_rest * blockDim.x) {
      if ((i >= s_up && i < s_up + c_up) || (i >= s_lo && i < s_lo + c_lo)) continue;
      const double2 ri = r[i], mi = minv[i];
      double2 pi = p[i];
      pi.x = fma(beta, pi.x, mi.x * ri.x);
      pi.y = fma(beta, pi.y, mi.y * ri.y);
      p[i] = pi;
    }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    reinterpret_cast<double*>(L + PW_OUT)[0] = rz_new;
    reinterpret_cast<double*>(L + PW_OUT)[1] = rr;
  }
  if (arrive_last(L + PW_TICKET, gridDim.x, &sh_last)) {  // every block has read the iteration counter: advance it
    if (threadIdx.x == 0) {
      L[PW_IT] = it + 1;
      *reinterpret_cast<unsigned*>(L + PW_TICKET) = 0u;
    }
  }
}

// Start of a solve: the all-reduced {r'z, r'r} of the initial residual (scal[0], scal[1] of fem_pcg_init) become the
// "previous iteration" slots; accumulators, ticket and error word are cleared.  Local stores only.
__global__ void ppcg_begin_kernel(uint64_t* L, const double* __restrict__ scal, int world) {
  if (threadIdx.x == 0) {
    const uint64_t it = L[PW_IT];
    uint64_t* lines = L + PW_LL_RZ + ((it + 1) & 1) * 4 * FEM_PEER_MAX;
    for (int k = 0; k < FEM_PEER_MAX; ++k) {
      line_store(lines + 4 * k, k == 0 ? scal[0] : 0.0, (uint32_t)it);
      line_store(lines + 4 * k + 2, k == 0 ? scal[1] : 0.0, (uint32_t)it);
    }
    for (int k = 0; k < 3; ++k) reinterpret_cast<double*>(L + PW_ACC)[k] = 0.0;
    reinterpret_cast<double*>(L + PW_OUT)[0] = scal[0];
    reinterpret_cast<double*>(L + PW_OUT)[1] = scal[1];
    L[PW_TICKET] = 0;
    L[PW_TICKET_HALO] = 0;
    L[PW_ERR] = 0;
    // the halo flags already hold `it` (stored by the neighbours' last C kernel; 0 in a fresh block): the first A kernel
    // passes at once, so the ghost rows of the initial p come from fem_halo_push followed by a host-level barrier
  }
}

static int make_view(PeerView* pv, void* comm, const void* const* peers, int rank, int world) {
  FEM_REQUIRE(comm && peers && world >= 1 && world <= FEM_PEER_MAX && rank >= 0 && rank < world, "peer view");
  memset(pv, 0, sizeof(*pv));
  pv->local = reinterpret_cast<uint64_t*>(comm);
  for (int r = 0; r < world; ++r) {
    FEM_REQUIRE(peers[r] != nullptr, "null peer block");
    pv->peer[r] = reinterpret_cast<uint64_t*>(const_cast<void*>(peers[r]));
  }
  pv->peer[rank] = pv->local;
  pv->rank = rank;
  pv->world = world;
  pv->nowait = g_fem_tuning.peer_nowait;
  pv->timeout_ns = (uint64_t)(g_fem_tuning.peer_timeout_ms > 0 ? g_fem_tuning.peer_timeout_ms : 10000) * 1000000ull;
  return FEM_OK;
}

static unsigned vec_grid(const int64_t n_items, const int sm_count) {
  int64_t b = (n_items + 255) / 256;
  const int64_t cap = (int64_t)(sm_count > 0 ? sm_count : 148) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

extern "C" int fem_ppcg_words(void) { return PW_WORDS; }

extern "C" int fem_ppcg_begin(void* comm, const double* scal, int world, fem_stream stream) {
  FEM_REQUIRE(comm && scal && world >= 1 && world <= FEM_PEER_MAX, "null pointer or world size");
  ppcg_begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint64_t*>(comm), scal, world);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_ppcg_spmv_dot(const fem_plan* P, const double* K_vals, const double* p, double* q, const uint8_t* free_mask,
                                 void* comm, const void* const* peers, int rank, int world, fem_stream stream) {
  FEM_REQUIRE(P && K_vals && p && q, "null pointer");
  FEM_REQUIRE((reinterpret_cast<uintptr_t>(K_vals) & 15u) == 0 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(q) & 15u) == 0, "K_vals, p, q must be 16-byte aligned");
  PeerView pv;
  int rc = make_view(&pv, comm, peers, rank, world);
  if (rc != FEM_OK) return rc;
  const SpmvShape sh = spmv_shape(P);
  cudaStream_t st = (cudaStream_t)stream;
  FEM_REQUIRE(P->tile_seg != nullptr, "plan without SpMV tiles");
  const unsigned tb = spmv_tile_blocks(P);
  const int32_t* tseg = P->tile_seg;
#define PSPMV(G) ppcg_spmv_kernel<G><<<tb, FEM_SPMV_THREADS, 0, st>>>(P->n_n, P->n_tiles, P->nbr_ptr, P->nbr_idx, P->nbr_loc, tseg, K_vals, p, q, free_mask, pv)
  if (sh.group == 4) PSPMV(4);
  else if (sh.group == 8) PSPMV(8);
  else PSPMV(16);
#undef PSPMV
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_ppcg_update_xr(const fem_plan* P, const double* p, const double* q, const double* minv, double* x, double* r,
                                  void* comm, const void* const* peers, int rank, int world, fem_stream stream) {
  FEM_REQUIRE(P && p && q && minv && x && r, "null pointer");
  PeerView pv;
  int rc = make_view(&pv, comm, peers, rank, world);
  if (rc != FEM_OK) return rc;
  const int64_t n2 = P->n_dof / 2;
  ppcg_update_xr_kernel<<<vec_grid(n2, P->sm_count), 256, 0, (cudaStream_t)stream>>>(
      n2, reinterpret_cast<const double2*>(p), reinterpret_cast<const double2*>(q), reinterpret_cast<const double2*>(minv),
      reinterpret_cast<double2*>(x), reinterpret_cast<double2*>(r), pv);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}

extern "C" int fem_ppcg_update_p(const fem_plan* P, int64_t own_lo, int64_t own_hi, const double* r, const double* minv, double* p,
                                 int64_t src_up, int64_t n_up, double* dst_up, int64_t src_lo, int64_t n_lo, double* dst_lo,
                                 void* comm, const void* const* peers, int rank, int world, fem_stream stream) {
  FEM_REQUIRE(P && r && minv && p && own_lo >= 0 && own_hi > own_lo && own_hi <= P->n_dof && own_lo % 2 == 0 && own_hi % 2 == 0,
              "null pointer or bad owned range");
  FEM_REQUIRE(src_up % 2 == 0 && src_lo % 2 == 0 && n_up % 2 == 0 && n_lo % 2 == 0, "halo ranges must be whole nodes");
  PeerView pv;
  int rc = make_view(&pv, comm, peers, rank, world);
  if (rc != FEM_OK) return rc;
  if (!dst_up) n_up = 0;
  if (!dst_lo) n_lo = 0;
  FEM_REQUIRE((n_up == 0 || rank < world - 1) && (n_lo == 0 || rank > 0), "halo destination without a neighbour on that side");
  const unsigned n_halo_blocks = (unsigned)((n_up / 2 + n_lo / 2 + 255) / 256);
  const unsigned grid = n_halo_blocks + vec_grid((own_hi - own_lo) / 2, P->sm_count);
  ppcg_update_p_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      own_lo / 2, own_hi / 2, reinterpret_cast<const double2*>(r), reinterpret_cast<const double2*>(minv), reinterpret_cast<double2*>(p),
      src_up / 2, n_up / 2, reinterpret_cast<double2*>(dst_up), src_lo / 2, n_lo / 2, reinterpret_cast<double2*>(dst_lo), n_halo_blocks, pv);
  FEM_CUDA_CHECK(cudaGetLastError());
  return FEM_OK;
}
