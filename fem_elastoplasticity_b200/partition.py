"""General element-block partition by recursive coordinate bisection (BASELINE north_star: "the mesh is
partitioned by element blocks (METIS-free coordinate bisection)") for meshes the strip partition of distributed.py cannot
describe: unstructured meshes (tsx-tunnel/pythonFEM.py:1687-1688), tiles of a uniform mesh, any number of neighbours.

* Elements are split by recursive bisection of their centroids along the longer side of the current bounding box, the
  part sizes proportional to the number of ranks on each side (so any rank count works, not only powers of two).
* A node belongs to the lowest rank among the elements around it; the rows of K follow the node.
* Every rank keeps its own elements plus the ghost elements that touch its owned nodes, so assembly, strain, return map
  and internal force need no communication and the owned rows are complete (owner computes), exactly as on strips.
* Local numbering: owned nodes first (ascending global id), ghost nodes after; own elements first, ghost elements after.
  The owned DOFs are therefore the contiguous range [0, 2 n_owned).
* Only the solver communicates: ``halo_exchange`` sends the owned values the neighbours hold as ghosts (index lists in
  ascending global id on both sides) in one batched send/recv group; scalars are all-reduced.

``GeneralPartition`` offers what ``DistributedPCG`` / ``NewtonSolver`` use of ``StripPartition``.  Every rank builds the
partition of the whole (host-resident) mesh redundantly - it is O(n log n) pre-processing next to the once-per-mesh plan
build; the exchanges themselves run on the device (NCCL; gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def rcb(centroids, n_parts):
    """Recursive coordinate bisection: (2, n_e) centroids -> (n_e,) part ids in [0, n_parts).  Deterministic (stable
    argsort; ties broken by element id)."""
    c = np.asarray(centroids, dtype=np.float64)
    part = np.zeros(c.shape[1], dtype=np.int64)

    def split(idx, first, n):
        if n == 1:
            part[idx] = first
            return
        n_lo = n // 2
        ext = c[:, idx].max(axis=1) - c[:, idx].min(axis=1) if idx.size else np.zeros(2)
        axis = 0 if ext[0] >= ext[1] else 1
        order = idx[np.argsort(c[axis, idx], kind="stable")]
        cut = (idx.size * n_lo) // n
        split(order[:cut], first, n_lo)
        split(order[cut:], first + n_lo, n - n_lo)

    split(np.arange(c.shape[1]), 0, n_parts)
    return part


class GeneralPartition:
    def __init__(self, elements, coordinates, rank, world, part_of_elem=None):
        el = np.asarray(elements.cpu() if isinstance(elements, torch.Tensor) else elements).astype(np.int64)
        co = np.asarray(coordinates.cpu() if isinstance(coordinates, torch.Tensor) else coordinates, dtype=np.float64)
        self.rank, self.world = rank, world
        n_p, n_e = el.shape
        n_n = co.shape[1]
        self.n_n_global, self.n_e_global = n_n, n_e
        if part_of_elem is None:
            part_of_elem = rcb(co[:, el].mean(axis=1), world)
        self.part_of_elem = pe = np.asarray(part_of_elem, dtype=np.int64)
        owner = np.full(n_n, world, dtype=np.int64)
        for p in range(n_p):
            np.minimum.at(owner, el[p], pe)
        assert owner.max() < world, "mesh has nodes that belong to no element"
        self.node_owner = owner
        # local sets of every rank (needed to derive the send lists without communication)
        self._local = [self._local_sets(el, pe, owner, r) for r in range(world)]
        own_nodes, ghost_nodes, own_elems, ghost_elems = self._local[rank]
        self.nodes = np.concatenate([own_nodes, ghost_nodes])           # local -> global node id
        self.elems = np.concatenate([own_elems, ghost_elems])           # local -> global element id
        self.n_owned, self.n_ghost = own_nodes.size, ghost_nodes.size
        self.n_n_local, self.n_e_owned, self.n_e_local = self.nodes.size, own_elems.size, self.elems.size
        g2l = np.full(n_n, -1, dtype=np.int64)
        g2l[self.nodes] = np.arange(self.nodes.size)
        self.elements_local = g2l[el[:, self.elems]]
        assert self.elements_local.min() >= 0
        # exchange lists: recv[s] = my ghost positions owned by s; send[s] = my owned positions that s holds as ghosts
        self.recv, self.send = {}, {}
        for s in range(world):
            if s == rank:
                continue
            mine_from_s = ghost_nodes[owner[ghost_nodes] == s]
            if mine_from_s.size:
                self.recv[s] = g2l[mine_from_s]
            theirs = self._local[s][1]
            to_s = theirs[owner[theirs] == rank]
            if to_s.size:
                self.send[s] = g2l[to_s]
        self.neighbours = sorted(set(self.recv) | set(self.send))
        self._dev_lists = {}

    @staticmethod
    def _local_sets(el, pe, owner, r):
        own_nodes = np.flatnonzero(owner == r)
        touches = (owner[el] == r).any(axis=0)                           # elements around r's owned nodes
        own_elems = np.flatnonzero(pe == r)
        ghost_elems = np.flatnonzero(touches & (pe != r))
        nodes = np.unique(el[:, np.concatenate([own_elems, ghost_elems])])
        ghost_nodes = nodes[owner[nodes] != r]
        return own_nodes, ghost_nodes, own_elems, ghost_elems

    # -- what the solver layer uses --------------------------------------------------------------------------------------
    def local_mesh(self, mesh, device):
        """This rank's part of a global mesh dict (coordinates, elements, and any (2, n_n) nodal arrays such as Q and
        dirichlet_nodes): owned + ghost nodes in local numbering."""
        out = {"elements": torch.as_tensor(self.elements_local.astype(np.int32)).to(device)}
        for k, v in mesh.items():
            if k == "elements":
                continue
            a = np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v)
            if a.ndim == 2 and a.shape[1] == self.n_n_global:
                out[k] = torch.as_tensor(np.ascontiguousarray(a[:, self.nodes])).to(device)
        return out

    def owned_dof_range(self):
        return 0, 2 * self.n_owned

    def owned_mask(self, device):
        m = torch.zeros(2 * self.n_n_local, dtype=torch.uint8, device=device)
        m[:2 * self.n_owned] = 1
        return m

    def free_owned_mask(self, plan, mesh):
        return plan.mask_u8(mesh["Q"]) & self.owned_mask(plan.device)

    def _lists(self, device):
        key = str(device)
        if key not in self._dev_lists:
            self._dev_lists[key] = ({s: torch.as_tensor(v).to(device) for s, v in self.send.items()},
                                    {s: torch.as_tensor(v).to(device) for s, v in self.recv.items()})
        return self._dev_lists[key]

    def halo_exchange(self, *vs):
        """Ghost entries of the DOF-interleaved vectors ``vs`` <- their owners' values; one batched send/recv group."""
        if self.world == 1 or not vs:
            return
        send, recv = self._lists(vs[0].device)
        ops, bufs = [], []
        for s in self.neighbours:
            if s in send:
                out = torch.stack([v.view(-1, 2)[send[s]] for v in vs]).contiguous()
                ops.append(dist.P2POp(dist.isend, out, s))
            if s in recv:
                buf = torch.empty((len(vs), recv[s].numel(), 2), dtype=vs[0].dtype, device=vs[0].device)
                ops.append(dist.P2POp(dist.irecv, buf, s))
                bufs.append((s, buf))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for s, buf in bufs:
            for i, v in enumerate(vs):
                v.view(-1, 2)[recv[s]] = buf[i]

    def all_reduce(self, t):
        if self.world > 1:
            dist.all_reduce(t)

    def gather_nodal(self, local, n_comp=2):
        """Owned part of a local nodal array (n_comp, n_n_local) -> (global node ids, values) for assembling a global
        field on the host."""
        return self.nodes[:self.n_owned], np.asarray(local)[:, :self.n_owned]
