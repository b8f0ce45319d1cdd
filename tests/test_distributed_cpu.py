"""Host-side logic of the multi-GPU path on CPU: strip partition, ghost layers, halo exchange and the
placement of the scalar all-reduces in the PCG, run with world_size 2 and 3 over gloo.  The CUDA step
kernels are replaced HERE (tests only) by a NumPy statement of the same five steps; the result must equal
a single-domain sparse direct solve of K[Q,Q] x = b."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fem_oracle as fo


class NumpyOps:
    """Same contract as fem_elastoplasticity_b200.distributed.CudaOps (include/fem_b200.h: fem_pcg_*)."""

    def __init__(self, K):
        self.K, self.n = K.tocsr(), K.shape[0]

    def new_vec(self, n=None):
        return torch.zeros(self.n if n is None else n, dtype=torch.float64)

    def jacobi(self, k, mask, out):
        d = self.K.diagonal()
        m = mask.numpy().astype(bool)
        out.copy_(torch.as_tensor(np.where(m & (d != 0), 1.0 / np.where(d != 0, d, 1.0), 0.0)))

    def pcg_init(self, rhs, kx0, mask, minv, r, p, scal):
        scal.zero_()
        b = rhs * mask
        r.copy_(b if kx0 is None else (b - kx0 * mask))
        p.copy_(minv * r)
        scal[0], scal[1], scal[4] = torch.dot(r, p), torch.dot(r, r), torch.dot(b, b)

    def spmv_dot(self, k, p, q, mask, scal, it):
        scal[0 if it & 1 else 2] = 0.0
        scal[1] = 0.0
        q.copy_(torch.as_tensor(self.K @ p.numpy()) * mask)
        scal[3] += torch.dot(p, q)

    def update_xr(self, p, q, minv, x, r, scal, it):
        old, new = (2, 0) if it & 1 else (0, 2)
        alpha = scal[old] / scal[3] if scal[3] != 0 else 0.0
        x += alpha * p
        r -= alpha * q
        scal[new] += torch.dot(r * minv, r)
        scal[1] += torch.dot(r, r)

    def update_p(self, r, minv, p, scal, it):
        old, new = (2, 0) if it & 1 else (0, 2)
        beta = scal[new] / scal[old] if scal[old] != 0 else 0.0
        p.copy_(minv * r + beta * p)
        scal[3] = 0.0

    def spmv(self, k, x, y, mask, dot):
        y.copy_(torch.as_tensor(self.K @ x.numpy()) * mask)
        dot += torch.dot(x, y)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, ny, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fem_elastoplasticity_b200.distributed import DistributedPCG, StripPartition
    part = StripPartition(nx, ny, rank, world, size_x=10.0, size_y=10.0 * world)
    mesh = part.local_mesh("cpu")
    et = fo.ElementType.P1
    xi, wf = fo.quadrature_volume(et)
    _, d1, d2 = fo.local_basis_volume(et, xi)
    elem, coord = mesh["elements"].numpy().astype(np.int64), mesh["coordinates"].numpy()
    n_e = elem.shape[1]
    G0, K0 = fo.footing_constants()[:2]
    K = fo.elastic_stiffness(elem, coord, G0 * np.ones(n_e), K0 * np.ones(n_e), d1, d2, wf)[0]
    q_local = mesh["Q"].t().reshape(-1).to(torch.uint8)
    mask = q_local & part.owned_mask("cpu")
    rng = np.random.default_rng(5)
    b_global = rng.standard_normal(2 * (nx + 1) * (ny + 1))
    lo = part.iy0 * part.row_dofs
    rhs = torch.as_tensor(b_global[lo:lo + 2 * part.n_n_local].copy())
    pcg = DistributedPCG(None, part, mask, ops=NumpyOps(K))
    x, its = pcg.solve(None, rhs, rtol=1e-13, maxit=5000, check_every=10)
    en = pcg.energy_norms(None, x.clone(), rhs.clone(), x.clone())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), x=x.numpy(), lo=lo, its=its, en=en.numpy(),
             own=np.array(part.owned_dof_range()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_strip_partitioned_pcg_equals_single_domain(tmp_path, world):
    nx, ny = 10, 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nx, ny, str(tmp_path)), nprocs=world, join=True)
    et = fo.ElementType.P1
    xi, wf = fo.quadrature_volume(et)
    _, d1, d2 = fo.local_basis_volume(et, xi)
    m = fo.square_mesh_p1(nx, ny, 10.0, 10.0 * world)
    n_e = m["elements"].shape[1]
    G0, K0 = fo.footing_constants()[:2]
    K = fo.elastic_stiffness(m["elements"], m["coordinates"], G0 * np.ones(n_e), K0 * np.ones(n_e), d1, d2, wf)[0].tocsr()
    qf = m["Q"].flatten(order="F")
    b = np.random.default_rng(5).standard_normal(K.shape[0])
    ref = np.zeros_like(b)
    ref[qf] = spla.spsolve(K[qf][:, qf].tocsc(), b[qf])
    got = np.full_like(b, np.nan)
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        lo, (a, e) = int(d["lo"]), d["own"]
        got[lo + a:lo + e] = d["x"][a:e]
        # ghost rows hold the neighbour's converged values after the final halo exchange
        np.testing.assert_allclose(d["x"], ref[lo:lo + d["x"].size], rtol=1e-8, atol=1e-10 * np.abs(ref).max())
        assert d["its"] > 0
        np.testing.assert_allclose(d["en"][1], b @ (K @ b), rtol=1e-11)         # distributed energy product
    assert not np.isnan(got).any()                                                # every DOF owned exactly once
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())


def test_partition_bookkeeping():
    from fem_elastoplasticity_b200.distributed import StripPartition
    nx, ny, world = 7, 12, 4
    owned = np.zeros(2 * (nx + 1) * (ny + 1), dtype=int)
    n_e = 0
    for r in range(world):
        p = StripPartition(nx, ny, r, world)
        lo = p.iy0 * p.row_dofs
        a, e = p.owned_dof_range()
        owned[lo + a:lo + e] += 1
        n_e += p.n_e_owned
        assert p.n_node_rows == p.ny_loc + 1 + (1 if p.has_upper else 0)
    assert (owned == 1).all() and n_e == 2 * nx * ny
    with pytest.raises(ValueError):
        StripPartition(4, 10, 0, 3)


# ---- two-level PCG: host logic of fem_elastoplasticity_b200.twolevel.TwoLevelPCG under gloo -------------------------
def _bilinear_prolongation(coord, grid):
    x0, y0, hx, hy, ncx, ncy = grid
    fx, fy = (coord[0] - x0) / hx, (coord[1] - y0) / hy
    cx, cy = np.clip(fx.astype(int), 0, ncx - 1), np.clip(fy.astype(int), 0, ncy - 1)
    xi, et = np.clip(fx - cx, 0, 1), np.clip(fy - cy, 0, 1)
    w = [(1 - xi) * (1 - et), xi * (1 - et), xi * et, (1 - xi) * et]
    base = cx + cy * (ncx + 1)
    ids = [base, base + 1, base + 1 + (ncx + 1), base + (ncx + 1)]
    n_n = coord.shape[1]
    rows = np.concatenate([2 * np.arange(n_n) + c for c in range(2) for _ in range(4)])
    cols = np.concatenate([2 * ids[k] + c for c in range(2) for k in range(4)])
    vals = np.concatenate([w[k] for _ in range(2) for k in range(4)])
    import scipy.sparse as sp
    return sp.csr_matrix((vals, (rows, cols)), shape=(2 * n_n, 2 * (ncx + 1) * (ncy + 1)))


class NumpyTLOps:
    """Same contract as fem_elastoplasticity_b200.twolevel.CudaTLOps (csrc/twolevel.cu), stated with SciPy on local arrays."""

    def __init__(self, K, coord):
        self.K, self.n_dof, self.n_n = K.tocsr(), K.shape[0], coord.shape[1]
        self.coord, self.device, self._np_coord, self._P = torch.as_tensor(coord), torch.device("cpu"), coord, {}

    def P(self, grid):
        if grid not in self._P:
            self._P[grid] = _bilinear_prolongation(self._np_coord, grid)
        return self._P[grid]

    def sync(self):
        pass

    jacobi = NumpyOps.jacobi

    def galerkin(self, k, row_mask, col_mask, grid, Ac):
        import scipy.sparse as sp
        P = self.P(grid)
        cm = row_mask if col_mask is None else col_mask
        A = (sp.diags(row_mask.numpy().astype(float)) @ P).T @ self.K @ (sp.diags(cm.numpy().astype(float)) @ P)
        Ac.copy_(torch.as_tensor(A.toarray()))

    def tl_init(self, rhs, mask, minv, grid, r, rc, scal):
        scal.zero_()
        b = rhs * mask
        r.copy_(b)
        rc.copy_(torch.as_tensor(self.P(grid).T @ r.numpy()))
        scal[0], scal[1], scal[4] = torch.dot(r * minv, r), torch.dot(r, r), torch.dot(b, b)

    def gemv(self, n, A, x, y, dot):
        y.copy_(A @ x)
        if dot is not None:
            dot += torch.dot(x, y)

    def tl_apply(self, mode, r, minv, mask, grid, zc, p, scal, it):
        z = torch.as_tensor(self.P(grid) @ zc.numpy()) * mask + minv * r
        if mode == 1:
            p.copy_(z)
        else:
            old, new = (2, 0) if it & 1 else (0, 2)
            beta = scal[new] / scal[old] if scal[old] != 0 else 0.0
            p.copy_(z + beta * p)
            scal[3] = 0.0

    spmv_dot = NumpyOps.spmv_dot

    def tl_update_xr(self, p, q, minv, grid, x, r, rc, scal, it):
        old, new = (2, 0) if it & 1 else (0, 2)
        alpha = scal[old] / scal[3] if scal[3] != 0 else 0.0
        x += alpha * p
        r -= alpha * q
        scal[1] += torch.dot(r, r)
        scal[new] += torch.dot(r * minv, r)
        rc.copy_(torch.as_tensor(self.P(grid).T @ r.numpy()))


def _local_problem(part, nx, ny):
    mesh = part.local_mesh("cpu")
    et = fo.ElementType.P1
    xi, wf = fo.quadrature_volume(et)
    _, d1, d2 = fo.local_basis_volume(et, xi)
    elem, coord = mesh["elements"].numpy().astype(np.int64), mesh["coordinates"].numpy()
    G0, K0 = fo.footing_constants()[:2]
    n_e = elem.shape[1]
    K = fo.elastic_stiffness(elem, coord, G0 * np.ones(n_e), K0 * np.ones(n_e), d1, d2, wf)[0]
    free = mesh["Q"].t().reshape(-1).to(torch.uint8)
    b_global = np.random.default_rng(5).standard_normal(2 * (nx + 1) * (ny + 1))
    lo = part.iy0 * part.row_dofs
    return K, coord, free, torch.as_tensor(b_global[lo:lo + 2 * part.n_n_local].copy()), lo


def _tl_worker(rank, world, port, nx, ny, nc, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fem_elastoplasticity_b200.distributed import StripPartition
    from fem_elastoplasticity_b200.twolevel import TwoLevelPCG
    part = StripPartition(nx, ny, rank, world, size_x=10.0, size_y=27.0)
    K, coord, free, rhs, lo = _local_problem(part, nx, ny)
    tl = TwoLevelPCG(None, free & part.owned_mask("cpu"), nc=nc, part=part, free_mask=free, ops=NumpyTLOps(K, coord))
    x, its, rel = tl.solve(None, rhs, rtol=1e-12, maxit=3000, check_every=1)
    np.savez(os.path.join(out_dir, f"t{rank}.npz"), x=x.numpy(), lo=lo, its=its, rel=rel, own=np.array(part.owned_dof_range()),
             grid=np.array(tl.grid))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_strip_partitioned_two_level_pcg_equals_single_domain(tmp_path, world):
    """TwoLevelPCG's host logic (partitioned Galerkin product with owned rows x free columns, replicated coarse solve,
    r_c'z_c added by rank 0, placement of the all-reduces) with ranks that have ghost rows on BOTH sides (world 4), on a
    coarse grid with hx != hy that does not nest in the fine mesh: same iterates as the single-domain run."""
    from fem_elastoplasticity_b200.distributed import StripPartition
    from fem_elastoplasticity_b200.twolevel import TwoLevelPCG
    nx, ny, nc = 10, 24, 3                                     # domain 10 x 27: coarse grid 3 x 8, hx = 3.33, hy = 3.375
    mp.spawn(_tl_worker, args=(world, _free_port(), nx, ny, nc, str(tmp_path)), nprocs=world, join=True)
    one = StripPartition(nx, ny, 0, 1, size_x=10.0, size_y=27.0)
    K, coord, free, rhs, _ = _local_problem(one, nx, ny)
    tl = TwoLevelPCG(None, free, nc=nc, ops=NumpyTLOps(K, coord))
    ref, ref_its, ref_rel = tl.solve(None, rhs, rtol=1e-12, maxit=3000, check_every=1)
    ref = ref.numpy()
    assert ref_rel <= 1e-12 and tl.grid[2] != tl.grid[3]
    jac = DistributedPCGJacobiCount(K, free, rhs)
    assert ref_its < jac                                       # the coarse correction pays even on this small mesh
    got = np.full_like(ref, np.nan)
    for r in range(world):
        d = np.load(tmp_path / f"t{r}.npz")
        assert np.allclose(d["grid"], np.array(tl.grid))       # every rank lays the same coarse grid over the GLOBAL box
        assert abs(int(d["its"]) - ref_its) <= 1, (int(d["its"]), ref_its)
        lo, (a, e) = int(d["lo"]), d["own"]
        got[lo + a:lo + e] = d["x"][a:e]
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())


def DistributedPCGJacobiCount(K, free, rhs):
    from fem_elastoplasticity_b200.distributed import DistributedPCG, StripPartition
    n_rows = K.shape[0] // 2
    part = StripPartition(10, n_rows // 11 - 1, 0, 1)
    pcg = DistributedPCG(None, part, free, ops=NumpyOps(K))
    return pcg.solve(None, rhs, rtol=1e-12, maxit=20000, check_every=1)[1]


# ---- general partition by recursive coordinate bisection (partition.py) ------------------------------------------------
def _general_mesh(name):
    if name == "tsx":
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "assembly_tsx_p1.npz"))
        return g["elements"].astype(np.int64), g["coordinates"], fo.tsx_q_mask(g["coordinates"])
    m = fo.square_mesh_p1(14, 10, 10.0, 7.0)
    return m["elements"].astype(np.int64), m["coordinates"], m["Q"]


@pytest.mark.parametrize("name,world", [("tsx", 2), ("tsx", 3), ("square", 4), ("square", 8)])
def test_rcb_partition_bookkeeping(name, world):
    from fem_elastoplasticity_b200.partition import GeneralPartition, rcb
    el, co, _ = _general_mesh(name)
    pe = rcb(co[:, el].mean(axis=1), world)
    sizes = np.bincount(pe, minlength=world)
    assert sizes.min() >= el.shape[1] // world - 1 and sizes.max() <= -(-el.shape[1] // world) + 1      # balanced element blocks
    parts = [GeneralPartition(el, co, r, world) for r in range(world)]
    owned = np.concatenate([p.nodes[:p.n_owned] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(co.shape[1]))                                      # every node owned exactly once
    assert np.array_equal(np.sort(np.concatenate([p.elems[:p.n_e_owned] for p in parts])), np.arange(el.shape[1]))
    for p in parts:
        # owned rows are complete: every element around an owned node is local
        around = np.flatnonzero(np.isin(el, p.nodes[:p.n_owned]).any(axis=0))
        assert np.isin(around, p.elems).all()
        assert np.array_equal(p.nodes[p.elements_local], el[:, p.elems])
        for s in p.neighbours:                                                                       # send / recv lists pair up
            q = parts[s]
            if s in p.send:
                assert np.array_equal(p.nodes[p.send[s]], q.nodes[q.recv[p.rank]])
            if s in p.recv:
                assert (p.recv[s] >= p.n_owned).all() and np.array_equal(p.nodes[p.recv[s]], q.nodes[q.send[p.rank]])
        ghosts = np.sort(np.concatenate([p.recv[s] for s in p.recv])) if p.recv else np.array([], dtype=np.int64)
        assert np.array_equal(ghosts, np.arange(p.n_owned, p.n_n_local))                             # every ghost has exactly one source
    if name == "square" and world == 4:                                                              # 2 x 2 tiles: > 2 neighbours
        assert max(len(p.neighbours) for p in parts) == 3


def _general_worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fem_elastoplasticity_b200.distributed import DistributedPCG
    from fem_elastoplasticity_b200.partition import GeneralPartition
    el, co, Q = _general_mesh(name)
    part = GeneralPartition(el, co, rank, world)
    lm = part.local_mesh({"coordinates": co, "Q": Q}, "cpu")
    et = fo.ElementType.P1
    xi, wf = fo.quadrature_volume(et)
    _, d1, d2 = fo.local_basis_volume(et, xi)
    n_e = part.n_e_local
    K = fo.elastic_stiffness(part.elements_local, lm["coordinates"].numpy(), 3.0 * np.ones(n_e), 5.0 * np.ones(n_e), d1, d2, wf)[0]
    mask = lm["Q"].t().reshape(-1).to(torch.uint8) & part.owned_mask("cpu")
    b_global = np.random.default_rng(5).standard_normal(2 * co.shape[1]).reshape(-1, 2)
    rhs = torch.as_tensor(b_global[part.nodes].reshape(-1).copy())
    pcg = DistributedPCG(None, part, mask, ops=NumpyOps(K))
    x, its = pcg.solve(None, rhs, rtol=1e-13, maxit=5000, check_every=10)
    en = pcg.energy_norms(None, x.clone(), rhs.clone(), x.clone())
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), x=x.numpy(), nodes=part.nodes, n_owned=part.n_owned, its=its, en=en.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("tsx", 2), ("tsx", 3), ("square", 4)])
def test_rcb_partitioned_pcg_equals_single_domain(tmp_path, name, world):
    """Jacobi-PCG over the RCB partition (index-list halos with up to 3 neighbours per rank, unstructured tsx mesh) against
    the single-domain sparse direct solve; ghost entries hold their owners' converged values after the final exchange."""
    mp.spawn(_general_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    el, co, Q = _general_mesh(name)
    et = fo.ElementType.P1
    xi, wf = fo.quadrature_volume(et)
    _, d1, d2 = fo.local_basis_volume(et, xi)
    n_e = el.shape[1]
    K = fo.elastic_stiffness(el, co, 3.0 * np.ones(n_e), 5.0 * np.ones(n_e), d1, d2, wf)[0].tocsr()
    qf = np.asarray(Q, dtype=bool).flatten(order="F")
    b = np.random.default_rng(5).standard_normal(K.shape[0])
    ref = np.zeros_like(b)
    ref[qf] = spla.spsolve(K[qf][:, qf].tocsc(), b[qf])
    ref2 = ref.reshape(-1, 2)
    got = np.full_like(ref2, np.nan)
    for r in range(world):
        d = np.load(tmp_path / f"g{r}.npz")
        x2, nodes, n_own = d["x"].reshape(-1, 2), d["nodes"], int(d["n_owned"])
        got[nodes[:n_own]] = x2[:n_own]
        np.testing.assert_allclose(x2, ref2[nodes], rtol=1e-8, atol=1e-10 * np.abs(ref).max())      # ghosts included
        np.testing.assert_allclose(d["en"][1], b @ (K @ b), rtol=1e-11)
    assert not np.isnan(got).any()
