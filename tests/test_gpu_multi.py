"""Two-GPU test of the distributed PCG (NCCL halo, NVLink peer-memory halo, and the fused iteration whose three exchanges
live inside the kernels): all must reproduce the single-GPU solve.  Skipped when fewer than two CUDA devices are visible (the driver's `pytest -m gpu` box has one)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, ny, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fem_elastoplasticity_b200 import meshgen
    from fem_elastoplasticity_b200 import pythonFEM as api
    from fem_elastoplasticity_b200.distributed import DistributedPCG, StripPartition
    from fem_elastoplasticity_b200.plan import FemPlan
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    part = StripPartition(nx, ny, rank, world, size_x=10.0, size_y=10.0 * world)
    mesh = part.local_mesh(dev)
    P = FemPlan(mesh["elements"], mesh["coordinates"], d1, d2, wf, device=dev)
    G, Kb, _, _ = meshgen.footing_materials(P.n_int, dev)
    k = P.assemble_elastic(G, Kb)
    mask = part.free_owned_mask(P, mesh)
    b_global = np.random.default_rng(9).standard_normal(2 * (nx + 1) * (ny + 1))
    lo = part.iy0 * part.row_dofs
    rhs = torch.as_tensor(b_global[lo:lo + P.n_dof].copy()).to(dev)
    res = {}
    for name, peer, graph in (("nccl", False, False), ("peer", True, False), ("fused", "fused", True)):
        pcg = DistributedPCG(P, part, mask, peer=peer, use_graph=graph)
        x, its = pcg.solve(k, rhs.clone(), rtol=1e-12, maxit=20000, check_every=25)
        res[name] = x.cpu().numpy().copy()
        res[name + "_its"] = its
        res[name + "_is_peer"] = pcg.peer is not None
        if name in ("nccl", "fused"):                     # a second solve on the same object, fixed iteration count
            x, n2 = pcg.solve(k, 2.0 * rhs, iters=70)
            res[name + "_second"] = x.cpu().numpy().copy()
            assert n2 == 70
        if name == "fused":
            assert pcg.fused and pcg._graph is not None, getattr(pcg, "graph_error", "fused iteration was not captured")
    from fem_elastoplasticity_b200.twolevel import TwoLevelPCG
    tl = TwoLevelPCG(P, mask, nc=8, part=part, free_mask=P.mask_u8(mesh["Q"]))
    x, its, rel = tl.solve(k, rhs.clone(), rtol=1e-12, maxit=20000, check_every=5)
    res["twolevel"], res["twolevel_its"], res["twolevel_is_peer"] = x.cpu().numpy().copy(), its, False
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, own=np.array(part.owned_dof_range()), **res)
    dist.destroy_process_group()


def test_two_gpu_pcg_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from fem_elastoplasticity_b200 import meshgen
    from fem_elastoplasticity_b200 import pythonFEM as api
    from fem_elastoplasticity_b200.plan import FemPlan
    nx, ny, world = 96, 128, 2
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, str(tmp_path)), nprocs=world, join=True)
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    m = meshgen.square_mesh_p1(nx, ny, 10.0, 10.0 * world)
    P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
    G, Kb, _, _ = meshgen.footing_materials(P.n_int)
    k = P.assemble_elastic(G, Kb)
    b = np.random.default_rng(9).standard_normal(P.n_dof)
    ref, its, rel = P.pcg(k, b, P.mask_u8(m["Q"]), rtol=1e-12, maxit=20000, check_every=25)
    ref = ref.cpu().numpy()
    for name in ("nccl", "peer", "fused", "twolevel"):
        got = np.full_like(ref, np.nan)
        for r in range(world):
            d = np.load(tmp_path / f"r{r}.npz")
            lo, (a, e) = int(d["lo"]), d["own"]
            got[lo + a:lo + e] = d[name][a:e]
            assert d[name + "_its"] > 0
            if name in ("peer", "fused"):
                assert bool(d[name + "_is_peer"]), "symmetric-memory halo was not active"
            if name == "fused":                           # same iteration, exchanges inside the kernels: same iterates
                assert int(d["fused_its"]) == int(d["nccl_its"])
                sa, sb = d["fused_second"][a:e], d["nccl_second"][a:e]
                np.testing.assert_allclose(sa, sb, rtol=1e-8, atol=1e-10 * np.abs(sb).max())
        assert not np.isnan(got).any()
        np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
    tl_its = int(np.load(tmp_path / "r0.npz")["twolevel_its"])
    assert tl_its < int(np.load(tmp_path / "r0.npz")["nccl_its"]) // 2
    # the partitioned coarse operator (owned rows x free columns, summed over ranks) is the single-domain one: same count
    from fem_elastoplasticity_b200.twolevel import TwoLevelPCG
    _, one_its, _ = TwoLevelPCG(P, P.mask_u8(m["Q"]), nc=8).solve(k, torch.as_tensor(b).cuda(), rtol=1e-12, maxit=20000, check_every=5)
    print("two-level iterations: 2 GPUs", tl_its, "1 GPU", one_its)
    assert abs(tl_its - one_its) <= max(5, one_its // 10)


def _newton_worker(rank, world, port, nx, ny, steps, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fem_elastoplasticity_b200 import newton
    from fem_elastoplasticity_b200.distributed import StripPartition
    part = StripPartition(nx, ny, rank, world, size_x=10.0, size_y=10.0)
    out = newton.footing_driver(part.local_mesh(dev), max_steps=steps, pcg_rtol=1e-12, part=part)
    a, e = part.first_owned_row * (nx + 1), (part.last_owned_row + 1) * (nx + 1)        # owned local nodes
    np.savez(os.path.join(out_dir, f"n{rank}.npz"), U=out["U"][:, a:e], node0=part.iy0 * (nx + 1) + a, steps=out["steps"],
             trace=np.array(out["trace"], dtype=float), hist=np.array(out["hist"], dtype=float),
             ep=out["Ep"][:, :part.n_e_owned], elem0=2 * nx * part.iy0)
    dist.destroy_process_group()


def test_two_gpu_footing_newton_matches_single_gpu(tmp_path):
    """The whole load-stepping Newton loop sharded over two GPUs (strip partition, fused PCG exchanges, all-reduced
    criterion / plastic count / footing pressure) follows the single-GPU run: same branches, same displacements."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from fem_elastoplasticity_b200 import meshgen, newton
    nx, ny, world, steps = 40, 48, 2, 6
    mp.spawn(_newton_worker, args=(world, _free_port(), nx, ny, steps, str(tmp_path)), nprocs=world, join=True)
    ref = newton.footing_driver(meshgen.square_mesh_p1(nx, ny, 10.0, 10.0), max_steps=steps, pcg_rtol=1e-12)
    rt = np.array(ref["trace"], dtype=float)
    U = np.full_like(ref["U"], np.nan)
    ep = np.full_like(ref["Ep"], np.nan)
    for r in range(world):
        d = np.load(tmp_path / f"n{r}.npz")
        assert int(d["steps"]) == ref["steps"]
        t = d["trace"]
        assert t.shape == rt.shape, "same number of Newton iterations on every rank as on one GPU"
        assert np.array_equal(t[:, :3], rt[:, :3])                     # load factor, iteration index, plastic points
        np.testing.assert_allclose(t[:, 3], rt[:, 3], rtol=1e-4, atol=1e-11)
        np.testing.assert_allclose(d["hist"], np.array(ref["hist"], dtype=float), rtol=1e-8)
        n0, e0 = int(d["node0"]), int(d["elem0"])
        U[:, n0:n0 + d["U"].shape[1]] = d["U"]
        ep[:, e0:e0 + d["ep"].shape[1]] = d["ep"]
    assert not np.isnan(U).any() and not np.isnan(ep).any()
    print("plastic points per Newton iteration:", rt[:, 2].astype(int).tolist())
    np.testing.assert_allclose(U, ref["U"], rtol=1e-7, atol=1e-9 * np.abs(ref["U"]).max())
    np.testing.assert_allclose(ep, ref["Ep"], rtol=1e-6, atol=1e-9 * max(np.abs(ref["Ep"]).max(), 1e-30))


def test_single_rank_graph_pcg_matches_c_loop():
    """World size 1: the Python-sequenced PCG (CUDA-graph replay of iteration pairs, and eager) equals fem_pcg."""
    import torch
    from fem_elastoplasticity_b200 import meshgen
    from fem_elastoplasticity_b200 import pythonFEM as api
    from fem_elastoplasticity_b200.distributed import DistributedPCG, StripPartition
    from fem_elastoplasticity_b200.plan import FemPlan
    nx, ny = 64, 80
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    part = StripPartition(nx, ny, 0, 1)
    mesh = part.local_mesh("cuda")
    P = FemPlan(mesh["elements"], mesh["coordinates"], d1, d2, wf)
    G, Kb, _, _ = meshgen.footing_materials(P.n_int)
    k = P.assemble_elastic(G, Kb)
    mask = part.free_owned_mask(P, mesh)
    b = torch.as_tensor(np.random.default_rng(4).standard_normal(P.n_dof)).cuda()
    ref, its, rel = P.pcg(k, b, mask, rtol=0.0, maxit=200, check_every=200, raise_on_maxit=False)
    for graph in (True, False):
        pcg = DistributedPCG(P, part, mask, use_graph=graph)
        x, n_it = pcg.solve(k, b.clone(), iters=200)
        assert n_it == 200
        assert (pcg._graph is not None) == graph
        np.testing.assert_allclose(x.cpu().numpy(), ref.cpu().numpy(), rtol=1e-9, atol=1e-12 * float(ref.abs().max()))


def _mg_worker(rank, world, port, nx, ny, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fem_elastoplasticity_b200 import meshgen
    from fem_elastoplasticity_b200 import pythonFEM as api
    from fem_elastoplasticity_b200.distributed import StripPartition
    from fem_elastoplasticity_b200.mg import MultigridPCG
    from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    part = StripPartition(nx, ny, rank, world, size_x=10.0, size_y=10.0 * ny / nx)
    mesh = part.local_mesh(dev)
    P = FemPlan(mesh["elements"], mesh["coordinates"], d1, d2, wf, device=dev)
    G, Kb, eta, c = meshgen.footing_materials(P.n_int, dev)
    k_el = P.assemble_elastic(G, Kb)
    rm = dp_return_map(meshgen.synthetic_strain_global(P.n_int, 2 * nx * part.iy0, dev), None, G, Kb, eta, c)
    k_tan = P.assemble_tangent(rm["ds"])
    mask = part.free_owned_mask(P, mesh)
    b_global = np.random.default_rng(9).standard_normal(2 * (nx + 1) * (ny + 1))
    lo = part.iy0 * part.row_dofs
    rhs = torch.as_tensor(b_global[lo:lo + P.n_dof].copy()).to(dev)
    # replicate_below=0: keep the small levels of this test mesh distributed (ghost-row exchanges on every level but the last)
    M = MultigridPCG(P, mask, part=part, free_mask=P.mask_u8(mesh["Q"]), max_coarse_dofs=300, replicate_below=0).setup(k_el)
    res = {"levels": np.array([[lv["rep"], lv["first_rep"]] for lv in M.lv])}
    for tag, graph in (("graph", True), ("eager", False)):
        M.use_graph = graph
        x, its, rel = M.solve(k_tan, rhs.clone(), rtol=1e-11)
        res[tag], res[tag + "_its"], res[tag + "_rel"] = x.cpu().numpy().copy(), its, rel
    res["graph_captured"] = M._graph is not None
    np.savez(os.path.join(out_dir, f"m{rank}.npz"), lo=lo, own=np.array(part.owned_dof_range()), **res)
    dist.destroy_process_group()


def test_multi_gpu_multigrid_matches_single_gpu(tmp_path):
    """Geometric multigrid PCG on a strip partition (2 GPUs, 4 when the box has them): distributed levels with ghost rows
    pushed over NVLink peer memory inside the V-cycle, replicated coarse levels behind a gather, scalar all-reduces through
    peer memory, the whole iteration replayed from a CUDA graph - against the single-GPU solve of the same system."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from fem_elastoplasticity_b200 import meshgen
    from fem_elastoplasticity_b200 import pythonFEM as api
    from fem_elastoplasticity_b200.mg import MultigridPCG
    from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map
    world = 4 if torch.cuda.device_count() >= 4 else 2
    nx, ny = 90, 128
    mp.spawn(_mg_worker, args=(world, _free_port(), nx, ny, str(tmp_path)), nprocs=world, join=True)
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    m = meshgen.square_mesh_p1(nx, ny, 10.0, 10.0 * ny / nx)
    P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
    G, Kb, eta, c = meshgen.footing_materials(P.n_int)
    k_el = P.assemble_elastic(G, Kb)
    rm = dp_return_map(meshgen.synthetic_strain_global(P.n_int, 0), None, G, Kb, eta, c)
    k_tan = P.assemble_tangent(rm["ds"])
    mask = P.mask_u8(m["Q"])
    b = torch.as_tensor(np.random.default_rng(9).standard_normal(P.n_dof)).cuda()
    M1 = MultigridPCG(P, mask, max_coarse_dofs=300).setup(k_el)
    ref, its1, _ = M1.solve(k_tan, b, rtol=1e-11)
    ref = ref.cpu().numpy()
    d0 = np.load(tmp_path / "m0.npz")
    assert d0["levels"][:, 0].sum() < len(d0["levels"]) and d0["levels"][:, 1].sum() == 1, "expected distributed levels and one gather level"
    assert bool(d0["graph_captured"])
    for tag in ("graph", "eager"):
        got = np.full_like(ref, np.nan)
        for r in range(world):
            d = np.load(tmp_path / f"m{r}.npz")
            lo, (a, e) = int(d["lo"]), d["own"]
            got[lo + a:lo + e] = d[tag][a:e]
        assert not np.isnan(got).any()
        np.testing.assert_allclose(got, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
        assert abs(int(d0[tag + "_its"]) - its1) <= 2, (tag, int(d0[tag + "_its"]), its1)
    print(f"multigrid PCG iterations: {world} GPUs", int(d0["graph_its"]), "1 GPU", its1)


def _rcb_worker(rank, world, port, case, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fem_elastoplasticity_b200 import newton
    from fem_elastoplasticity_b200.partition import GeneralPartition
    if case == "tsx":
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "assembly_tsx_p1.npz"))
        part = GeneralPartition(g["elements"], g["coordinates"], rank, world)
        out = newton.tsx_driver(g["coordinates"], g["elements"], pcg_rtol=1e-13, part=part)
    else:
        from oracle import fem_oracle as fo
        m = fo.footing_mesh(1, fo.ElementType.P1)
        part = GeneralPartition(m["elements"], m["coordinates"], rank, world)
        lm = part.local_mesh({k: m[k] for k in ("coordinates", "Q", "dirichlet_nodes")}, dev)
        lm["elements"] = torch.as_tensor(part.elements_local.astype(np.int32)).to(dev)
        out = newton.footing_driver(lm, max_steps=5, pcg_rtol=1e-13, part=part)
    np.savez(os.path.join(out_dir, f"{case}{rank}.npz"), U=out["U"][:, :part.n_owned], nodes=part.nodes[:part.n_owned], steps=out["steps"],
             trace=np.array(out["trace"], dtype=float), neighbours=len(part.neighbours))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["tsx", "footing"])
def test_rcb_partitioned_newton_matches_single_gpu(tmp_path, case):
    """The load-stepping Newton loops on a recursive-coordinate-bisection partition (partition.GeneralPartition: index-list
    halos, any number of neighbours): the unstructured tsx-tunnel mesh and the footing mesh cut into element blocks follow
    the single-GPU runs - same Newton iterations and plastic-point counts per step, same displacements."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from fem_elastoplasticity_b200 import newton
    from oracle import fem_oracle as fo
    world = 4 if torch.cuda.device_count() >= 4 else 2
    mp.spawn(_rcb_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    if case == "tsx":
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "assembly_tsx_p1.npz"))
        ref = newton.tsx_driver(g["coordinates"], g["elements"], pcg_rtol=1e-13)
    else:
        m = fo.footing_mesh(1, fo.ElementType.P1)
        ref = newton.footing_driver({k: m[k] for k in ("coordinates", "elements", "Q", "dirichlet_nodes")}, max_steps=5, pcg_rtol=1e-13)
    U = np.full_like(ref["U"], np.nan)
    for r in range(world):
        d = np.load(tmp_path / f"{case}{r}.npz")
        U[:, d["nodes"]] = d["U"]
        assert int(d["steps"]) == ref["steps"]
        tr, rt = d["trace"], np.array(ref["trace"], dtype=float)
        assert tr.shape == rt.shape and np.array_equal(tr[:, 1:3], rt[:, 1:3])            # Newton iteration index, plastic points
        np.testing.assert_allclose(tr[:, 3], rt[:, 3], rtol=1e-4, atol=1e-11)             # criteria
    assert not np.isnan(U).any()
    err = np.abs(U - ref["U"]).max() / np.abs(ref["U"]).max()
    print(case, f"RCB partition on {world} GPUs: displacement difference to the single-GPU run", err)
    assert err <= 1e-9
