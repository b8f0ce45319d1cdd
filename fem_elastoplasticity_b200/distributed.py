"""Multi-GPU layer: element-block (strip) partition by coordinate bisection along y, one process per GPU.

Assembly, strain, return map and internal force shard with NO communication: every rank also holds the
one layer of ghost elements that touches its owned nodes, so its owned matrix rows are complete
(owner-computes).  Only the PCG communicates: before each SpMV the interface DOFs of the
search direction are exchanged with the two neighbouring strips (one contiguous node row each way,
2(nx+1) doubles), and the dot products are all-reduced as device-resident scalars (one scalar after the SpMV,
one adjacent pair {r'z, r'r} after the vector update).  Collectives go through torch.distributed
(NCCL on GPUs; gloo in the CPU tests of the host logic).

The reference has no distributed code at all; the single-rank path below is the same
sequence of kernels with the collectives skipped."""
import ctypes as C

import torch
import torch.distributed as dist


class PCGNotConverged(ArithmeticError):
    """Raised by the distributed / two-level solves when ``maxit`` iterations did not reach ``rtol`` (the single-GPU
    fem_pcg returns FEM_ERR_PCG_MAXIT in the same situation).  ``x`` holds the last iterate."""

    def __init__(self, what, iters, relres, rtol):
        super().__init__(f"{what}: relative residual {relres:.3e} > rtol {rtol:.1e} after {iters} iterations")
        self.iters, self.relres, self.rtol = iters, relres, rtol


class StripPartition:
    """Rank r owns cell rows [r*ny_loc, (r+1)*ny_loc) of an nx x ny_global uniform P1 mesh and the node rows
    (r*ny_loc, (r+1)*ny_loc] (rank 0 also owns node row 0): interface nodes belong to the lower rank."""

    def __init__(self, nx, ny_global, rank, world, size_x=10.0, size_y=10.0):
        if ny_global % world:
            raise ValueError("ny_global must be divisible by the number of ranks")
        self.nx, self.ny_global, self.rank, self.world = nx, ny_global, rank, world
        self.size_x, self.size_y = size_x, size_y
        self.ny_loc = ny_global // world
        self.iy0 = rank * self.ny_loc                                   # first local cell/node row (global index)
        self.has_lower = rank > 0
        self.has_upper = rank < world - 1
        self.n_cell_rows = self.ny_loc + (1 if self.has_upper else 0)   # + ghost cell row above
        self.n_node_rows = self.n_cell_rows + 1
        self.row_dofs = 2 * (nx + 1)
        self.first_owned_row = 1 if self.has_lower else 0               # local node-row indices
        self.last_owned_row = self.ny_loc
        self.n_e_owned = 2 * nx * self.ny_loc
        self.n_n_local = (nx + 1) * self.n_node_rows

    def local_mesh(self, device):
        from . import meshgen
        return meshgen.square_mesh_p1(self.nx, self.n_cell_rows, self.size_x, self.size_y, device=device, iy0=self.iy0,
                                      n_y_global=self.ny_global, size_y_global=self.size_y)

    def owned_dof_range(self):
        return self.first_owned_row * self.row_dofs, (self.last_owned_row + 1) * self.row_dofs

    def owned_mask(self, device):
        m = torch.zeros(self.n_n_local * 2, dtype=torch.uint8, device=device)
        lo, hi = self.owned_dof_range()
        m[lo:hi] = 1
        return m

    def free_owned_mask(self, plan, mesh):
        """uint8 DOF mask: free (not Dirichlet) AND owned by this rank."""
        return plan.mask_u8(mesh["Q"]) & self.owned_mask(plan.device)

    def row_slice(self, v, row):
        return v[row * self.row_dofs:(row + 1) * self.row_dofs]

    def halo_exchange(self, *vs):
        """Fill the ghost node rows of the DOF vectors ``vs`` from the neighbouring strips (no-op on one rank); all
        vectors travel in one batched send/recv group."""
        if self.world == 1:
            return
        ops = []
        for v in vs:
            if self.has_upper:   # my top owned row -> upper rank's bottom ghost row; its first owned row -> my top ghost row
                ops.append(dist.P2POp(dist.isend, self.row_slice(v, self.last_owned_row), self.rank + 1))
                ops.append(dist.P2POp(dist.irecv, self.row_slice(v, self.last_owned_row + 1), self.rank + 1))
            if self.has_lower:
                ops.append(dist.P2POp(dist.isend, self.row_slice(v, self.first_owned_row), self.rank - 1))
                ops.append(dist.P2POp(dist.irecv, self.row_slice(v, 0), self.rank - 1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def all_reduce(self, t):
        if self.world > 1:
            dist.all_reduce(t)


class PeerHalo:
    """Halo exchange of the PCG search direction over NVLink peer memory (torch symmetric memory = CUDA IPC mappings).

    ``p`` lives in a symmetric allocation; ``fem_pcg_update_p_push`` stores this rank's interface rows into the neighbours'
    ghost rows from the kernel that computes them, and stream-ordered signals (put_signal / wait_signal) tell the neighbour
    that its ghosts are current.  Overwriting a neighbour's ghost rows is safe because the all-reduce of {r'z, r'r} sits
    between every SpMV (the reader) and the next p update (the writer).  With ``make_comm`` the same object also carries
    the communication block of the fused iteration, where flags polled inside the kernels replace signals and all-reduces."""
    CHANNEL = 0

    def __init__(self, part, n_dof_local, device):
        import torch.distributed._symmetric_memory as symm
        self.part = part
        n_max = torch.tensor([n_dof_local], dtype=torch.int64, device=device)
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX)           # symmetric allocations have one size on every rank
        self.buf = symm.empty(int(n_max.item()), dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.p = self.buf[:n_dof_local]
        rd = part.row_dofs
        n_sym = int(n_max.item())
        # peer views of the same symmetric tensor (get_buffer accounts for the tensor's offset inside the allocation block)
        self.peers = {r: self.hdl.get_buffer(r, (n_sym,), torch.float64) for r in (part.rank - 1, part.rank + 1) if 0 <= r < part.world}
        # (source offset, count, peer destination pointer) towards the upper and the lower neighbour
        self.up = (part.last_owned_row * rd, rd, self.peers[part.rank + 1].data_ptr()) if part.has_upper else (0, 0, 0)
        self.lo = ((part.first_owned_row * rd, rd, self.peers[part.rank - 1].data_ptr() + (part.ny_loc + 1) * rd * 8)
                   if part.has_lower else (0, 0, 0))
        self.hdl.barrier(channel=1)
        self.comm = None

    def make_comm(self, n_words):
        """Communication block of the fused iteration (csrc/peer_pcg.cu): ``n_words`` zeroed 8-byte words in symmetric
        memory on every rank + the table of all ranks' blocks as mapped into this process."""
        import torch.distributed._symmetric_memory as symm
        part = self.part
        self.comm = symm.empty(n_words, dtype=torch.int64, device=self.buf.device)
        self.comm.zero_()
        self.comm_hdl = symm.rendezvous(self.comm, dist.group.WORLD)
        self._comm_views = [self.comm if r == part.rank else self.comm_hdl.get_buffer(r, (n_words,), torch.int64) for r in range(part.world)]
        self.comm_table = (C.c_void_p * part.world)(*[v.data_ptr() for v in self._comm_views])
        torch.cuda.synchronize()
        self.comm_hdl.barrier(channel=0)                 # every block is zeroed before any peer may store into it
        torch.cuda.synchronize()

    def barrier(self):
        self.hdl.barrier(channel=1)

    def push_args(self):
        return (self.up[0], self.up[1], C.c_void_p(self.up[2]), self.lo[0], self.lo[1], C.c_void_p(self.lo[2]))

    def signal(self):
        """After a push: tell both neighbours their ghost rows are written, then wait for theirs."""
        part = self.part
        if part.has_upper:
            self.hdl.put_signal(part.rank + 1, self.CHANNEL)
        if part.has_lower:
            self.hdl.put_signal(part.rank - 1, self.CHANNEL)
        if part.has_upper:
            self.hdl.wait_signal(part.rank + 1, self.CHANNEL)
        if part.has_lower:
            self.hdl.wait_signal(part.rank - 1, self.CHANNEL)


class CudaOps:
    """The PCG step kernels of the C ABI (include/fem_b200.h) on one plan."""

    def __init__(self, plan):
        from .plan import _ptr, _stream
        from ._lib import call
        self.plan, self._ptr, self._stream, self._call = plan, _ptr, _stream, call
        self.n = plan.n_dof

    def new_vec(self, n=None):
        return torch.zeros(self.n if n is None else n, dtype=torch.float64, device=self.plan.device)

    def jacobi(self, k, mask, out):
        self.plan.jacobi(k, mask, out=out)

    def pcg_init(self, rhs, kx0, mask, minv, r, p, scal):
        self._call("fem_pcg_init", self.n, self._ptr(rhs), self._ptr(kx0), self._ptr(mask), self._ptr(minv), self._ptr(r),
                   self._ptr(p), self._ptr(scal), self._stream())

    def spmv_dot(self, k, p, q, mask, scal, it):
        self._call("fem_pcg_spmv_dot", self.plan._h, self._ptr(k), self._ptr(p), self._ptr(q), self._ptr(mask), self._ptr(scal),
                   int(it), self._stream())

    def update_xr(self, p, q, minv, x, r, scal, it):
        self._call("fem_pcg_update_xr", self.n, self._ptr(p), self._ptr(q), self._ptr(minv), self._ptr(x), self._ptr(r),
                   self._ptr(scal), int(it), self._stream())

    def update_p(self, r, minv, p, scal, it):
        self._call("fem_pcg_update_p", self.n, self._ptr(r), self._ptr(minv), self._ptr(p), self._ptr(scal), int(it), self._stream())

    def update_p_push(self, own, r, minv, p, scal, it, push):
        self._call("fem_pcg_update_p_push", int(own[0]), int(own[1]), self._ptr(r), self._ptr(minv), self._ptr(p), self._ptr(scal), int(it), *push,
                   self._stream())

    def halo_push(self, v, push):
        self._call("fem_halo_push", self._ptr(v), *push, self._stream())

    def spmv(self, k, x, y, mask, dot):
        self.plan.spmv(k, x, mask=mask, out=y, dot=dot)

    # fused multi-GPU iteration: exchanges inside the kernels (peer stores + flags), see csrc/peer_pcg.cu
    def ppcg_words(self):
        from ._lib import load
        return int(load().fem_ppcg_words())

    def ppcg_begin(self, peer, scal):
        self._call("fem_ppcg_begin", self._ptr(peer.comm), self._ptr(scal), peer.part.world, self._stream())

    def ppcg_iteration(self, peer, k, p, q, mask, minv, x, r, own):
        part = peer.part
        tail = (self._ptr(peer.comm), peer.comm_table, part.rank, part.world, self._stream())
        self._call("fem_ppcg_spmv_dot", self.plan._h, self._ptr(k), self._ptr(p), self._ptr(q), self._ptr(mask), *tail)
        self._call("fem_ppcg_update_xr", self.plan._h, self._ptr(p), self._ptr(q), self._ptr(minv), self._ptr(x), self._ptr(r), *tail)
        self._call("fem_ppcg_update_p", self.plan._h, int(own[0]), int(own[1]), self._ptr(r), self._ptr(minv), self._ptr(p),
                   *peer.push_args(), *tail)


class DistributedPCG:
    """Jacobi-PCG over a strip partition.  ``ops`` defaults to the CUDA kernels; the CPU tests inject a
    NumPy implementation of the same five steps to exercise partition + halo + reduction logic under gloo."""

    def __init__(self, plan, part, mask, ops=None, peer=False, use_graph=True):
        self.part, self.mask = part, mask
        self.ops = ops if ops is not None else CudaOps(plan)
        o = self.ops
        self.r, self.p, self.q, self.x, self.minv = (o.new_vec() for _ in range(5))
        # CUDA graph of an (even, odd) iteration pair: removes the per-launch host gaps.  Single rank only: with NCCL
        # send/recv + all-reduces captured inside the graph a full-size 2-GPU run hung (round-1 finding), so multi-rank
        # runs launch eagerly.
        self.use_graph = bool(use_graph) and ops is None and part.world == 1
        self.use_graph_fused = bool(use_graph)           # the fused iteration has no library call inside: capturable on any world size
        self._graph, self._graph_key = None, None
        # exchanges of the iteration: peer=False NCCL send/recv + all-reduces; True: halo as NVLink peer stores fused into the
        # p-update kernel, NCCL all-reduces; "fused": halo and both reductions inside the three kernels (csrc/peer_pcg.cu,
        # fastest: 0.122 vs 0.212 ms/iteration at 8 GPUs x 2M DOFs); "auto": fused, NCCL if symmetric memory is unavailable
        self.peer = None
        self.fused = False
        if not hasattr(part, "row_dofs"):                 # general (RCB) partition: index-list halos, NCCL exchanges
            if peer in (True, "fused"):
                raise ValueError("the peer-memory exchanges are built for the strip partition")
            peer = False
        if ops is None and part.world > 1 and peer in ("auto", True, "fused"):
            try:
                self.peer = PeerHalo(part, self.ops.n, self.r.device)
                self.p = self.peer.p
                if peer in ("fused", "auto"):
                    self.peer.make_comm(self.ops.ppcg_words())
                    self.fused = True
            except Exception as e:                       # symmetric memory unavailable: keep the NCCL path
                if peer in (True, "fused"):
                    raise
                self.peer_error = repr(e)
                self.peer, self.fused = None, False
                self.p = o.new_vec()
        self.scal = o.new_vec(8)
        self.en = o.new_vec(3)
        self.owned = part.owned_mask(self.r.device)
        self.launches_last = 0

    def solve(self, k_vals, rhs, iters=None, rtol=1e-10, maxit=100000, check_every=50):
        """Fixed ``iters`` iterations (benchmarks) or until |r| <= rtol |b|.  Returns (x, iterations)."""
        o, part, scal = self.ops, self.part, self.scal
        o.jacobi(k_vals, self.mask, self.minv)
        self.x.zero_()
        o.pcg_init(rhs, None, self.mask, self.minv, self.r, self.p, scal)
        part.all_reduce(scal[0:5])
        n_it = iters if iters is not None else maxit
        self.relres = float("nan")
        if iters is None:                                 # already converged (zero right-hand side, exact initial guess)?
            h = scal.cpu()
            self.relres = float((h[1] / h[4]).sqrt()) if h[4] > 0 else 0.0
            if self.relres <= rtol:
                part.halo_exchange(self.x)
                return self.x, 0
        if self.fused:
            return self._solve_fused(k_vals, n_it, iters is None, rtol, check_every)
        it = 0
        self.launches_last = 3
        peer = self.peer
        if peer is not None:                              # ghosts of the initial search direction
            o.halo_push(self.p, peer.push_args())
            peer.signal()
            self.launches_last += 1

        def iteration(i):
            if peer is None:
                part.halo_exchange(self.p)
            o.spmv_dot(k_vals, self.p, self.q, self.mask, scal, i)
            part.all_reduce(scal[3:4])
            o.update_xr(self.p, self.q, self.minv, self.x, self.r, scal, i)
            part.all_reduce(scal[1:3] if i % 2 == 0 else scal[0:2])
            if peer is None:
                o.update_p(self.r, self.minv, self.p, scal, i)
            else:                                         # p update + halo push in one kernel, then stream-ordered signals
                o.update_p_push(part.owned_dof_range(), self.r, self.minv, self.p, scal, i, peer.push_args())
                peer.signal()

        graph = self._pair_graph(k_vals, iteration) if (self.use_graph and n_it >= 4) else None
        while it < n_it:
            if graph is not None and it % 2 == 0 and it + 2 <= n_it:
                graph.replay()                            # iterations it (even) and it+1 (odd)
                it += 2
                self.launches_last += 6
            else:
                iteration(it)
                it += 1
                self.launches_last += 3
            if iters is None and (it % check_every == 0 or it == n_it):
                h = scal.cpu()
                if not torch.isfinite(h[1]):
                    raise ArithmeticError("PCG breakdown: residual is not finite")
                self.relres = float((h[1] / h[4]).sqrt()) if h[4] > 0 else 0.0
                if h[1] <= rtol * rtol * h[4]:
                    break
        part.halo_exchange(self.x)
        if iters is None and self.relres > rtol:
            raise PCGNotConverged("distributed PCG", it, self.relres, rtol)
        return self.x, it

    GRAPH_CHUNK = 10
    WORD_ERR, WORD_OUT = 167, 172                     # FEM_PPCG_WORD_ERR / FEM_PPCG_WORD_OUT of include/fem_b200.h

    def _solve_fused(self, k_vals, n_it, check, rtol, check_every):
        """Iterations with the exchanges inside the kernels (csrc/peer_pcg.cu): no collective call per iteration.  After
        one eagerly launched chunk the same launches are captured (capturing does not execute) and replayed from a CUDA
        graph; the iteration index lives on the device, so one graph serves every iteration."""
        o, part, peer = self.ops, self.part, self.peer
        own = part.owned_dof_range()
        o.ppcg_begin(peer, self.scal)
        o.halo_push(self.p, peer.push_args())            # ghosts of the initial search direction
        peer.barrier()
        self.launches_last = 5
        out = peer.comm.view(torch.float64)[self.WORD_OUT:self.WORD_OUT + 2]     # global r'z, r'r of the last finished iteration
        bb = None

        def chunk(n):
            for _ in range(n):
                o.ppcg_iteration(peer, k_vals, self.p, self.q, self.mask, self.minv, self.x, self.r, own)

        g, it, c = None, 0, self.GRAPH_CHUNK
        key = (k_vals.data_ptr(), self.p.data_ptr())
        if self._graph is not None and self._graph_key == key:
            g = self._graph
        while it < n_it:
            n = min(c, n_it - it)
            if check:
                n = min(n, check_every - it % check_every)
            if g is not None and n == c:
                g.replay()
            else:
                chunk(n)
                if self.use_graph_fused and g is None and n == c and it + 2 * c <= n_it:
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            chunk(c)
                        self._graph, self._graph_key = g, key
                    except Exception as e:               # capture unsupported: eager launches
                        self.graph_error, self.use_graph_fused, g = repr(e), False, None
            it += n
            self.launches_last += 3 * n
            if check and (it % check_every == 0 or it == n_it):
                if bb is None:
                    bb = float(self.scal[4].item())
                self._check_peer_error()                 # all ranks decide together: a timed-out rank's r'r is garbage
                rr = float(out[1].item())
                if rr != rr or rr == float("inf"):
                    raise ArithmeticError("PCG breakdown: residual is not finite")
                self.relres = (rr / bb) ** 0.5 if bb > 0 else 0.0
                if rr <= rtol * rtol * bb:
                    break
        self._check_peer_error()
        part.halo_exchange(self.x)
        if check and self.relres > rtol:
            raise PCGNotConverged("fused distributed PCG", it, self.relres, rtol)
        return self.x, it

    def _check_peer_error(self):
        """The sticky time-out word of the fused kernels, MAX-reduced over the ranks so that every rank raises together
        (a rank that carried on alone would hang in the next collective)."""
        err = self.peer.comm[self.WORD_ERR:self.WORD_ERR + 1].clone()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if int(err.item()) != 0:
            raise RuntimeError("fused PCG: a peer did not publish within the time-out (tuning key peer_timeout_ms); detected on rank %d" % self.part.rank)

    def _pair_graph(self, k_vals, iteration):
        """CUDA graph of iterations (0, 1) - the kernels only depend on the parity of the iteration index.  Captured after
        one eager pair (NCCL communicators / lazy initialisation must not happen under capture); cached per matrix buffer."""
        key = (k_vals.data_ptr(), self.p.data_ptr())
        if self._graph is not None and self._graph_key == key:
            return self._graph
        try:
            saved = [t.clone() for t in (self.x, self.r, self.p, self.q, self.scal)]
            iteration(0)
            iteration(1)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                iteration(0)
                iteration(1)
            for t, sv in zip((self.x, self.r, self.p, self.q, self.scal), saved):
                t.copy_(sv)                               # the warm-up and the capture must not advance the solve
            torch.cuda.synchronize()
            self._graph, self._graph_key = g, key
        except Exception as e:                            # capture unsupported in this configuration: eager launches
            self.graph_error = repr(e)
            self.use_graph = False
            self._graph = None
        return self._graph

    def energy_norms(self, k_vals, v0, v1, v2):
        """v_i' K v_i over the whole (distributed) DOF set; returns a device tensor of 3 doubles."""
        self.en.zero_()
        self.part.halo_exchange(v0, v1, v2)
        for i, v in enumerate((v0, v1, v2)):
            self.ops.spmv(k_vals, v, self.q, self.owned, self.en[i:i + 1])
        self.part.all_reduce(self.en)
        return self.en
