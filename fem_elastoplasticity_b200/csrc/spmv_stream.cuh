// SpMV on the 2x2 node-block pattern with EVERY operand streamed through shared memory by bulk async copies.
//
// The register-fed kernels of spmv.cuh keep the bytes in flight in registers: 1024 threads x ~72 B per SM - enough for
// ~75 % of the HBM bandwidth with FP64 values and only ~50 % with the FP32 copy the multigrid smoother streams (half the
// bytes per load: measured the same run time as FP64).  Here one persistent CTA per SM walks tiles of FEM_SPMV_TILE
// consecutive nodes with a PRODUCER WARP and 16 CONSUMER WARPS:
//   producer   per tile: waits until the consumers have released a shared-memory stage (mbarrier `empty`), stores the
//              tile's row pointers, and one lane issues cp.async.bulk copies (UBLKCP, completing on mbarrier `full`) of
//                - the tile's matrix values   (contiguous: rows of consecutive nodes are adjacent in the block-CSR layout),
//                - its 16-bit block positions (nbr_loc),
//                - the <= 4 contiguous ranges of x its columns touch (plan: tile_seg),
//                - the node vectors the epilogue needs (b, D^-1, d, x of the multigrid smoother; the DOF mask).
//              Descriptor and row pointers are fetched one tile ahead, so no global-memory latency sits on its path
//              (without that prefetch the producer was the bottleneck: 0.57 ms per sweep instead of 0.35).
//   consumers  wait for `full`, compute the tile out of shared memory - U x B independent load/gather/FMA chains per lane,
//              no global load at all - run the epilogue, and arrive on `empty`.  No CTA-wide barrier in the loop.
// Two tiles (FP32 values: three stages of 70 KB) or one (FP64: two stages of 102 KB) are in flight per SM whatever the
// occupancy or register use.  Measured at 16M elements (profiles/r2p): Chebyshev step on FP32 values 0.352 ms = 5.1 TB/s of
// DRAM traffic (register-fed: 0.569 ms); plain FP64 SpMV 0.380 ms (register-fed: 0.455 ms).
// Tiles whose x ranges do not fit (plan: nseg == 0) gather x through L1/L2; plans whose tiles hold more blocks than a
// stage (P2/Q2 meshes) use the kernels of spmv.cuh.
#pragma once
#include "common.cuh"
#include "spmv.cuh"

#define FEM_STREAM_THREADS (FEM_SPMV_THREADS + 32)  // 16 consumer warps + 1 producer warp
#define FEM_STREAM_VCAP 2048  // 2x2 blocks per stage (a P1 tile holds ~1 800)
#define FEM_STREAM_NIN 5      // node vectors (one double2 per node) an epilogue can have staged

// Epilogue concept:
//   static constexpr int N_IN          node vectors staged per tile (<= FEM_STREAM_NIN)
//   const double2* in(int k) const     their base pointers (indexed by node)
//   const uint8_t* mask_ptr() const    DOF mask staged per tile (2 bytes per node) or nullptr
//   void operator()(int64_t a, double acc0, double acc1, const double2 (&s)[N_IN or 1], uchar2 mk, double& dot) const
//                                      called by the lead lane of node a with the staged values of that node
template <class VT>
struct SpmvStreamSmem {
  static constexpr int V_BYTES = FEM_STREAM_VCAP * 4 * (int)sizeof(VT);
  static constexpr int L_BYTES = (FEM_STREAM_VCAP + 8) * 2 + 112;  // + alignment slack of the source, padded to 128
  static constexpr int X_BYTES = FEM_SPMV_CAP * 16;
  static constexpr int E_BYTES = FEM_SPMV_TILE * 16;
  static constexpr int M_BYTES = FEM_SPMV_TILE * 2;
  static constexpr int P_BYTES = (FEM_SPMV_TILE + 1 + 31) / 32 * 128;
  static constexpr int STAGE = V_BYTES + L_BYTES + X_BYTES + FEM_STREAM_NIN * E_BYTES + M_BYTES + P_BYTES;
  // three stages when they fit (FP32 values: 3 x 70 KB), else two (FP64 values: 2 x 102 KB)
  static constexpr int NSTAGE = (3 * STAGE + 256 <= 227 * 1024) ? 3 : 2;
  static constexpr int TOTAL = NSTAGE * STAGE + 256;  // + barriers and per-stage scalars
  static_assert(L_BYTES % 128 == 0 && STAGE % 128 == 0, "stage alignment");
};

struct SpmvStreamArgs {
  int64_t n_n, n_tiles;
  const int32_t* nbr_ptr;
  const int32_t* nbr_idx;
  const uint16_t* nbr_loc;
  const int32_t* tile_seg;
};

__device__ __forceinline__ void stream_copy(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(spmv_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// blockDim.x == FEM_STREAM_THREADS, dynamic shared memory of SpmvStreamSmem<VT>::TOTAL bytes (128-byte aligned).
template <int GROUP, bool COHERENT, class VT, class EPI>
__device__ __forceinline__ double spmv_stream(const SpmvStreamArgs A, const VT* __restrict__ vals, const double* x, const EPI& epi,
                                              unsigned char* smem) {
  using S = SpmvStreamSmem<VT>;
  using P2 = typename SpmvPair<VT>::type;
  constexpr int U = 2;
  constexpr int GPC = FEM_SPMV_THREADS / GROUP;
  constexpr int NPS = GPC * U;
  constexpr int SWEEPS = FEM_SPMV_TILE / NPS;
  static_assert(FEM_SPMV_TILE % NPS == 0, "tile size");
  constexpr int NIN = EPI::N_IN;
  static_assert(NIN <= FEM_STREAM_NIN, "too many staged epilogue vectors");
  const int tid = threadIdx.x, sub = tid % GROUP;
  // Lane groups -> nodes.  With four lanes per node a warp holds eight consecutive nodes and a 64-bit (128-bit) shared-memory
  // load is served half (quarter) warp by half warp: the rows of consecutive P1 nodes lie 112 (224) bytes apart, so
  // neighbouring nodes overlap in half of their banks and every load of matrix values took twice its wavefronts (ncu,
  // profiles/r2zz).  Nodes two apart do not collide: groups 0-3 of a warp take the even nodes of the eight, groups 4-7
  // the odd ones, so a quarter warp holds the nodes n and n ^ 2 (the pairing the plan's bank phases assume, plan.cu).
  const int g_raw = tid / GROUP;
  const int gi = GROUP == 4 ? ((g_raw & ~7) | ((g_raw & 3) << 1) | ((g_raw >> 2) & 1)) : g_raw;
  // stage layout
  auto stage = [&](int b) { return smem + (size_t)b * S::STAGE; };
  auto vbuf = [&](int b) { return reinterpret_cast<const P2*>(stage(b)); };
  auto lbuf = [&](int b) { return reinterpret_cast<const uint16_t*>(stage(b) + S::V_BYTES); };
  auto xbuf = [&](int b) { return reinterpret_cast<const double2*>(stage(b) + S::V_BYTES + S::L_BYTES); };
  auto ebuf = [&](int b, int k) { return reinterpret_cast<const double2*>(stage(b) + S::V_BYTES + S::L_BYTES + S::X_BYTES + k * S::E_BYTES); };
  auto mbuf = [&](int b) { return reinterpret_cast<const uchar2*>(stage(b) + S::V_BYTES + S::L_BYTES + S::X_BYTES + FEM_STREAM_NIN * S::E_BYTES); };
  auto pbuf = [&](int b) { return reinterpret_cast<int32_t*>(stage(b) + S::V_BYTES + S::L_BYTES + S::X_BYTES + FEM_STREAM_NIN * S::E_BYTES + S::M_BYTES); };
  constexpr int NST = S::NSTAGE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NST * (size_t)S::STAGE);  // full[NST], empty[NST]
  int32_t* meta = reinterpret_cast<int32_t*>(bar + 8);  // per stage: [0] staged x?, [1] first block of the tile, [2] offset of its first position in lbuf

  // one thread: all copies of a tile into stage b (desc = the tile's 12-word descriptor, already in registers)
  auto issue = [&](const int64_t tile, const int b, const int32_t (&d)[FEM_SPMV_DESC]) {
    const uint32_t br = spmv_smem_u32(&bar[b]);
    const int64_t a0 = tile * FEM_SPMV_TILE;
    const int nodes = (int)((A.n_n - a0) < FEM_SPMV_TILE ? (A.n_n - a0) : FEM_SPMV_TILE);
    const int s0 = d[10], nblk = d[11], nseg = d[0];
    const int s0a = s0 & ~7;                                             // 16-byte aligned start of the positions
    const uint32_t lbytes = (uint32_t)((((s0 - s0a) + nblk) * 2 + 15) & ~15);
    const uint32_t vbytes = (uint32_t)nblk * 4u * (uint32_t)sizeof(VT);
    uint32_t total = vbytes + (nseg > 0 ? lbytes + (uint32_t)d[1] * 16u : 0u) + (uint32_t)NIN * (uint32_t)nodes * 16u;
    if (epi.mask_ptr()) total += (uint32_t)((nodes * 2 + 15) & ~15);
    meta[4 * b + 0] = nseg > 0;
    meta[4 * b + 1] = s0;
    meta[4 * b + 2] = s0 - s0a;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(br), "r"(total) : "memory");
    if (vbytes) stream_copy(stage(b), vals + 4 * (int64_t)s0, vbytes, br);
    if (nseg > 0) {
      stream_copy(stage(b) + S::V_BYTES, A.nbr_loc + s0a, lbytes, br);
      int off = 0;
      for (int k = 0; k < nseg; ++k) {
        const int start = d[2 + 2 * k], len = d[3 + 2 * k];
        stream_copy(stage(b) + S::V_BYTES + S::L_BYTES + (size_t)off * 16, reinterpret_cast<const double2*>(x) + start, (uint32_t)len * 16u, br);
        off += len;
      }
    }
#pragma unroll
    for (int k = 0; k < NIN; ++k)
      stream_copy(stage(b) + S::V_BYTES + S::L_BYTES + S::X_BYTES + k * S::E_BYTES, epi.in(k) + a0, (uint32_t)nodes * 16u, br);
    if (epi.mask_ptr())
      stream_copy(stage(b) + S::V_BYTES + S::L_BYTES + S::X_BYTES + FEM_STREAM_NIN * S::E_BYTES, epi.mask_ptr() + 2 * a0, (uint32_t)((nodes * 2 + 15) & ~15), br);
  };
  auto load_desc = [&](const int64_t tile, int32_t (&d)[FEM_SPMV_DESC]) {
#pragma unroll
    for (int k = 0; k < FEM_SPMV_DESC; ++k) d[k] = (tile < A.n_tiles) ? __ldg(A.tile_seg + tile * FEM_SPMV_DESC + k) : 0;
  };
  // Warp specialisation: warp 16 is the producer (row pointers of the tile by ordinary loads, then one lane issues the bulk
  // copies), warps 0-15 consume.  full[s]: the copies of stage s have landed (transaction count); empty[s]: all 16 consumer
  // warps are done with stage s.  No CTA-wide barrier inside the loop: consumer warps drift up to NST-1 tiles apart.
  uint64_t* full = bar;
  uint64_t* empty = bar + NST;
  constexpr int NCW = FEM_SPMV_THREADS / 32;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < NST; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spmv_smem_u32(&full[b])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spmv_smem_u32(&empty[b])), "r"(NCW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double dot = 0.0;
  if (tid >= FEM_SPMV_THREADS) {  // ---- producer warp
    const int lane = tid - FEM_SPMV_THREADS;
    if (COHERENT && lane == 0) asm volatile("fence.proxy.async;" ::: "memory");  // peer stores seen through the flags precede the async-proxy reads
    // descriptor and row pointers of a tile are fetched one tile ahead: the producer's loop then consists of the wait for a
    // free stage, a few shared-memory stores and the copy instructions - no global-memory latency on its path
    constexpr int NRP = (FEM_SPMV_TILE + 1 + 31) / 32;
    int32_t d[FEM_SPMV_DESC], dn[FEM_SPMV_DESC];
    int32_t rp[NRP], rpn[NRP];
    auto fetch = [&](const int64_t tile, int32_t (&dd)[FEM_SPMV_DESC], int32_t (&rr)[NRP]) {
      if (lane == 0) load_desc(tile, dd);
#pragma unroll
      for (int q = 0; q < NRP; ++q) {
        const int t = q * 32 + lane;
        const int64_t a = tile * FEM_SPMV_TILE + t;
        rr[q] = (tile < A.n_tiles && t <= FEM_SPMV_TILE) ? __ldg(A.nbr_ptr + (a < A.n_n ? a : A.n_n)) : 0;
      }
    };
    fetch(first, d, rp);
    int k = 0;
    for (int64_t tile = first; tile < A.n_tiles; tile += stride, ++k) {
      const int b = k % NST, use = k / NST;
      fetch(tile + stride, dn, rpn);
      if (use > 0) spmv_bar_wait(&empty[b], (uint32_t)((use - 1) & 1));  // the consumers have released this stage
#pragma unroll
      for (int q = 0; q < NRP; ++q) {
        const int t = q * 32 + lane;
        if (t <= FEM_SPMV_TILE) pbuf(b)[t] = rp[q];
      }
      __syncwarp();
      if (lane == 0) issue(tile, b, d);  // meta stores + arrive.expect_tx (release) + copies
#pragma unroll
      for (int q = 0; q < NRP; ++q) rp[q] = rpn[q];
#pragma unroll
      for (int q = 0; q < FEM_SPMV_DESC; ++q) d[q] = dn[q];
    }
    return 0.0;
  }
  // ---- consumer warps
  int k = 0;
  for (int64_t tile = first; tile < A.n_tiles; tile += stride, ++k) {
    const int cur = k % NST;
    spmv_bar_wait(&full[cur], (uint32_t)((k / NST) & 1));
    const bool staged = meta[4 * cur + 0] != 0;
    const int s0 = meta[4 * cur + 1], lofs = meta[4 * cur + 2];
    const P2* vb = vbuf(cur);
    const uint16_t* lb = lbuf(cur) + lofs;
    const double2* xs = xbuf(cur);
    const int32_t* pb = pbuf(cur);
    const int64_t a0 = tile * FEM_SPMV_TILE;
#pragma unroll 1
    for (int sw = 0; sw < SWEEPS; ++sw) {
      // U nodes x B blocks per lane with static indices and predicates: 2 U B independent shared-memory loads, then U B
      // independent x gathers, then the products - the warp has U B dependency chains in flight instead of one (a loop
      // over the blocks of one node at a time ran at 18 % issue utilisation: 16 warps per SM cannot hide serial LDS chains)
      constexpr int B = 2;
      int nl[U], p0[U], deg[U], rel[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        nl[u] = sw * NPS + u * GPC + gi;  // node within the tile
        p0[u] = 0;
        deg[u] = 0;
        if (a0 + nl[u] < A.n_n) {
          p0[u] = pb[nl[u]];
          deg[u] = pb[nl[u] + 1] - p0[u];
        }
        rel[u] = p0[u] - s0;
      }
      P2 w0[U][B], w1[U][B];
      int loc[U][B];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const int j = sub + b * GROUP;
          w0[u][b].x = w0[u][b].y = w1[u][b].x = w1[u][b].y = 0;
          loc[u][b] = 0;
          if (j < deg[u]) {
            w0[u][b] = vb[2 * rel[u] + j];
            w1[u][b] = vb[2 * rel[u] + deg[u] + j];
            loc[u][b] = staged ? (int)lb[rel[u] + j] : __ldg(A.nbr_idx + p0[u] + j);
          }
        }
      double2 xv[U][B];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int b = 0; b < B; ++b) {
          xv[u][b] = make_double2(0.0, 0.0);
          if (sub + b * GROUP < deg[u]) xv[u][b] = staged ? xs[loc[u][b]] : spmv_ldx<COHERENT>(x, loc[u][b]);
        }
      double acc0[U], acc1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc0[u] = fma((double)w0[u][1].y, xv[u][1].y, fma((double)w0[u][1].x, xv[u][1].x, fma((double)w0[u][0].y, xv[u][0].y, (double)w0[u][0].x * xv[u][0].x)));
        acc1[u] = fma((double)w1[u][1].y, xv[u][1].y, fma((double)w1[u][1].x, xv[u][1].x, fma((double)w1[u][0].y, xv[u][0].y, (double)w1[u][0].x * xv[u][0].x)));
        for (int j = sub + B * GROUP; j < deg[u]; j += GROUP) {  // rows longer than B * GROUP blocks
          const P2 v0 = vb[2 * rel[u] + j], v1 = vb[2 * rel[u] + deg[u] + j];
          const double2 xx = staged ? xs[lb[rel[u] + j]] : spmv_ldx<COHERENT>(x, __ldg(A.nbr_idx + p0[u] + j));
          acc0[u] = fma((double)v0.x, xx.x, acc0[u]);
          acc0[u] = fma((double)v0.y, xx.y, acc0[u]);
          acc1[u] = fma((double)v1.x, xx.x, acc1[u]);
          acc1[u] = fma((double)v1.y, xx.y, acc1[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) {
          acc0[u] += __shfl_xor_sync(0xffffffffu, acc0[u], o);
          acc1[u] += __shfl_xor_sync(0xffffffffu, acc1[u], o);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t a = a0 + nl[u];
        if (sub == 0 && a < A.n_n) {
          double2 sv[NIN > 0 ? NIN : 1];
#pragma unroll
          for (int k = 0; k < NIN; ++k) sv[k] = ebuf(cur, k)[nl[u]];
          const uchar2 mk = epi.mask_ptr() ? mbuf(cur)[nl[u]] : make_uchar2(1, 1);
          epi(a, acc0[u], acc1[u], sv, mk, dot);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(spmv_smem_u32(&empty[cur])) : "memory");
  }
  return dot;
}

// default epilogue: y = mask .* (K x), x'y
struct SpmvStreamStore {
  static constexpr int N_IN = 1;
  double* __restrict__ y;
  const uint8_t* mask;
  const double* x;
  bool want_dot;
  __device__ __forceinline__ const double2* in(int) const { return reinterpret_cast<const double2*>(x); }
  __device__ __forceinline__ const uint8_t* mask_ptr() const { return mask; }
  __device__ __forceinline__ void operator()(const int64_t a, double acc0, double acc1, const double2 (&s)[1], const uchar2 mk, double& dot) const {
    if (!mk.x) acc0 = 0.0;
    if (!mk.y) acc1 = 0.0;
    reinterpret_cast<double2*>(y)[a] = make_double2(acc0, acc1);
    if (want_dot) {
      dot = fma(s[0].x, acc0, dot);
      dot = fma(s[0].y, acc1, dot);
    }
  }
};

// the streaming kernels need every tile's blocks to fit one stage and 16-byte aligned operands
static inline bool spmv_can_stream(const fem_plan* P) {
  return P->tile_seg != nullptr && P->tile_max_blocks > 0 && P->tile_max_blocks <= FEM_STREAM_VCAP;
}
static inline unsigned spmv_stream_blocks(const fem_plan* P) {
  int64_t b = P->n_tiles < P->sm_count ? P->n_tiles : P->sm_count;  // persistent: one CTA per SM
  return (unsigned)(b < 1 ? 1 : b);
}
