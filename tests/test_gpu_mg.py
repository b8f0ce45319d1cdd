"""Geometric multigrid preconditioner (csrc/mg.cu, mg.py): every set-up kernel and the V-cycle against a NumPy/SciPy
statement of the same algorithm (tests/mg_reference.py), the preconditioned CG against the Jacobi-PCG solution, and the
Newton drivers with it against the reference traces.  The reference's solve is a dense LU
(Plasticity2D_DP/pythonFEM.py:1062-1066); any SPD preconditioner yields its solution."""
import numpy as np
import pytest

from oracle import fem_oracle as fo
from mg_reference import ReferenceMG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fem():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from fem_elastoplasticity_b200 import meshgen, mg, plan
    return {"torch": torch, "plan": plan, "mg": mg, "meshgen": meshgen}


def p1_tables():
    xi, wf = fo.quadrature_volume(fo.ElementType.P1)
    _, d1, d2 = fo.local_basis_volume(fo.ElementType.P1, xi)
    return d1, d2, wf


def stencil_to_scipy(S, nxn, nrows):
    import scipy.sparse as sp
    n = nxn * nrows
    S = S.reshape(9, 4, nrows, nxn)
    rows, cols, vals = [], [], []
    jj, ii = np.meshgrid(np.arange(nrows), np.arange(nxn), indexing="ij")
    for s in range(9):
        dx, dy = s % 3 - 1, s // 3 - 1
        ok = (ii + dx >= 0) & (ii + dx < nxn) & (jj + dy >= 0) & (jj + dy < nrows)
        node, nb = (ii + jj * nxn)[ok], (ii + dx + (jj + dy) * nxn)[ok]
        for q in range(4):
            rows.append(2 * node + q // 2), cols.append(2 * nb + q % 2), vals.append(S[s, q][ok])
        assert not S[s][:, ~ok].any()                      # nothing stored towards neighbours that do not exist
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * n, 2 * n))


@pytest.mark.parametrize("nx,ny", [(37, 21), (64, 48)])
def test_hierarchy_and_vcycle_match_numpy_statement(fem, nx, ny):
    torch, mg = fem["torch"], fem["mg"]
    d1, d2, wf = p1_tables()
    m = fem["meshgen"].square_mesh_p1(nx, ny, 10.0, 10.0 * ny / nx)
    P = fem["plan"].FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
    G, Kb, eta, c = fem["meshgen"].footing_materials(P.n_int)
    k_el = P.assemble_elastic(G, Kb)
    from fem_elastoplasticity_b200.plan import dp_return_map
    r = dp_return_map(fem["meshgen"].synthetic_strain(P.n_int), None, G, Kb, eta, c)
    k_tan = P.assemble_tangent(r["ds"])
    mask = P.mask_u8(m["Q"])
    M = mg.MultigridPCG(P, mask, degree=3, ratio=8.0, max_coarse_dofs=120, smoother_f32=False).setup(k_el)
    assert M.n_levels >= 3
    q = mask.cpu().numpy().astype(bool)
    Kel, Ktan = P.to_scipy_csr(k_el), P.to_scipy_csr(k_tan)
    ref = ReferenceMG(Kel, q, nx + 1, ny + 1, M.n_levels, 3, 8.0)
    for li, lv in enumerate(M.lv):
        assert (lv["nxn"], lv["nrows"]) == ref.dims[li + 1]
        got = stencil_to_scipy(lv["S"].cpu().numpy(), lv["nxn"], lv["nrows"])
        want = ref.A[li + 1]
        assert abs(got - want).max() <= 1e-12 * abs(want).max(), li
    # inverse 2x2 diagonal blocks (block-Jacobi smoother) and eigenvalue bounds: the power iteration underestimates, never
    # above the true lambda_max by more than the 1.1 margin
    import scipy.sparse.linalg as spla
    ref.set_bounds([1.0] * M.n_levels)
    M.block_jacobi(k_el)
    for l, (lm, dv) in enumerate([(M.lmax0, M.minv)] + [(lv["lmax"], lv["dinv"]) for lv in M.lv[:-1]]):
        A, Di = ref.A[l], ref.dinv[l]
        n = A.shape[0]
        dv = dv.cpu().numpy()
        got = np.stack([dv[0:n:2], dv[1:n:2], dv[n:2 * n:2], dv[n + 1:2 * n:2]])          # i00, i01, i10, i11 per node
        want_b = np.stack([Di.diagonal()[0::2], Di.diagonal(1)[0::2], Di.diagonal(-1)[0::2], Di.diagonal()[1::2]])
        np.testing.assert_allclose(got, want_b, rtol=1e-11, atol=1e-13 * np.abs(want_b).max())
        true = spla.eigs(spla.LinearOperator(A.shape, matvec=lambda v: Di @ (A @ v)), k=1, which="LM", return_eigenvectors=False, tol=1e-6)[0].real
        assert 0.9 * true <= lm <= 1.12 * true, (l, lm, true)
    # one V-cycle on the tangent matrix (coarse operators of K_elast) against the NumPy statement
    ref.set_fine(Ktan)
    ref.set_bounds([M.lmax0] + [lv["lmax"] for lv in M.lv[:-1]])
    M.block_jacobi(k_tan)
    rng = np.random.default_rng(3)
    rv = rng.standard_normal(P.n_dof) * q
    z = torch.zeros(P.n_dof, dtype=torch.float64, device="cuda")
    dot = torch.zeros(1, dtype=torch.float64, device="cuda")
    M.vcycle(k_tan, torch.as_tensor(rv).cuda(), z, dot)
    zr = ref.vcycle(rv)
    np.testing.assert_allclose(z.cpu().numpy(), zr, rtol=0, atol=1e-11 * np.abs(zr).max())
    assert abs(float(dot.item()) - rv @ zr) <= 1e-11 * abs(rv @ zr)
    # the V-cycle is symmetric: u'M(v) == v'M(u)
    uv = rng.standard_normal(P.n_dof) * q
    assert abs(uv @ zr - rv @ ref.vcycle(uv)) <= 1e-10 * abs(uv @ zr)


def test_multigrid_pcg_solves_the_newton_system(fem):
    """CG + V-cycle on K_tangent[Q,Q] x = -F[Q] (synthetic plastic state, coarse operators of K_elast): same solution as
    Jacobi-PCG, h-independent iteration count, graph replay == eager launches."""
    torch, mg = fem["torch"], fem["mg"]
    from fem_elastoplasticity_b200.plan import dp_return_map
    d1, d2, wf = p1_tables()
    its = {}
    for nx in (96, 384):
        m = fem["meshgen"].square_mesh_p1(nx, nx)
        P = fem["plan"].FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
        G, Kb, eta, c = fem["meshgen"].footing_materials(P.n_int)
        k_el = P.assemble_elastic(G, Kb)
        r = dp_return_map(fem["meshgen"].synthetic_strain(P.n_int), None, G, Kb, eta, c)
        k_tan, F = P.assemble_tangent_force(r["ds"], r["s"])
        mask = P.mask_u8(m["Q"])
        M = mg.MultigridPCG(P, mask).setup(k_el)
        x, n_it, rel = M.solve(k_tan, -F, rtol=1e-10)
        x = x.clone()
        true = float(((-F - P.spmv(k_tan, x)) * mask).norm() / (F * mask).norm())
        assert rel <= 1e-10 and true <= 2e-10, (rel, true)
        xj, _, _ = P.pcg(k_tan, -F, mask, rtol=1e-12, maxit=200000)
        assert float((x - xj).abs().max() / xj.abs().max()) <= 1e-7
        its[nx] = n_it
        M.use_graph = False
        x2, n2, _ = M.solve(k_tan, -F, rtol=1e-10)
        assert n2 == n_it and float((x2 - x).abs().max()) <= 1e-9 * float(x.abs().max())
        # the default streams an FP32 copy of the matrix in the smoother: same solution, (almost) the same count as FP64
        M64 = mg.MultigridPCG(P, mask, smoother_f32=False).setup(k_el)
        x3, n3, _ = M64.solve(k_tan, -F, rtol=1e-10)
        assert abs(n3 - n_it) <= 2 and float((x3 - x).abs().max()) <= 1e-8 * float(x.abs().max()), (n3, n_it)
    print("multigrid PCG iterations:", its)
    assert max(its.values()) <= 60 and its[384] <= its[96] + 8


def test_newton_drivers_with_multigrid(golden):
    """Footing load stepping with the multigrid-preconditioned solve: the reference's level-1 trace (109 Newton iterations,
    16 steps) and the dense-LU displacements."""
    from fem_elastoplasticity_b200 import newton
    f = golden("assembly_footing_p1_l1.npz")
    mesh = {k: f[k] for k in ("coordinates", "elements", "Q", "dirichlet_nodes")}
    out = newton.footing_driver(mesh, pcg_rtol=1e-13, precond="multigrid")
    oref = fo.footing_driver(1)
    assert len(out["trace"]) == len(oref["trace"]) == 109
    assert [t[2] for t in out["trace"]] == [t[2] for t in oref["trace"]]
    err = np.abs(out["U"] - oref["U"]).max() / np.abs(oref["U"]).max()
    print("footing L1 with multigrid: displacement error vs dense-LU oracle", err, "PCG iterations", [t[4] for t in out["trace"]][:8])
    assert err <= 1e-9


def test_fine_step_code_paths_agree(fem):
    """The level-0 multigrid step through its three SpMV code paths - all operands streamed through shared memory by bulk
    copies (default), x staged + matrix through registers, everything gathered - with the FP64 matrix and its FP32 copy:
    same results to rounding, and the FP64 one equal to the plain formulas."""
    torch, mg = fem["torch"], fem["mg"]
    from fem_elastoplasticity_b200 import _lib
    d1, d2, wf = p1_tables()
    m = fem["meshgen"].square_mesh_p1(333, 211, 10.0, 6.0)
    P = fem["plan"].FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
    G, Kb, _, _ = fem["meshgen"].footing_materials(P.n_int)
    k = P.assemble_elastic(G, Kb)
    mask = P.mask_u8(m["Q"])
    g = torch.Generator(device="cuda").manual_seed(4)
    rnd = lambda: torch.randn(P.n_dof, dtype=torch.float64, device="cuda", generator=g) * mask  # noqa: E731
    b, x = rnd(), rnd()
    for f32 in (False, True):
        M = mg.MultigridPCG(P, mask, smoother_f32=f32).setup(k)
        M.block_jacobi(k)
        if f32:
            _lib.call("fem_mg_to_f32", P.nnz, fem["plan"]._ptr(k), fem["plan"]._ptr(M.k32), fem["plan"]._stream())
        d0 = rnd()
        outs = []
        try:
            for staged in (0, 2, 1):
                _lib.call("fem_set_tuning", b"spmv_staged", staged)
                M.v0["d"].copy_(d0)
                out, res, dot = torch.zeros_like(x), torch.zeros_like(x), torch.zeros(1, dtype=torch.float64, device="cuda")
                M.fine_step(k, b, x, out, mode=2, step=1, dot=dot)
                M.fine_step(k, b, x, res, mode=1)
                outs.append((out.clone(), M.v0["d"].clone(), res.clone(), float(dot.item())))
        finally:
            _lib.call("fem_set_tuning", b"spmv_staged", 0)
        for o in outs[1:]:
            for a, c in zip(outs[0][:3], o[:3]):
                assert float((a - c).abs().max()) <= 1e-12 * float(c.abs().max())
            assert abs(outs[0][3] - o[3]) <= 1e-11 * abs(o[3])
        r_ref = (b - P.spmv(k, x)) * mask
        d_ref = M.desc.c1[1] * d0 + M.desc.c2[1] * M.block_apply(M.minv, r_ref)
        tol = 1e-12 if not f32 else 1e-6
        assert float((outs[0][2] - r_ref).abs().max()) <= tol * float(r_ref.abs().max())
        assert float((outs[0][1] - d_ref).abs().max()) <= tol * float(d_ref.abs().max())
        assert float((outs[0][0] - (x + d_ref)).abs().max()) <= tol * float((x + d_ref).abs().max())


def test_multigrid_solve_is_bit_reproducible(fem):
    """No atomics on floating-point data anywhere in the multigrid CG: the level-1 Galerkin product is a gather, the transfers
    are gathers, and every dot product adds its block sums in a fixed order.  Two independent set-ups + solves (graph replay and
    eager launches) give the same hierarchy and the same solution bit for bit."""
    torch, mg = fem["torch"], fem["mg"]
    from fem_elastoplasticity_b200.plan import dp_return_map
    d1, d2, wf = p1_tables()
    m = fem["meshgen"].square_mesh_p1(257, 190, 10.0, 7.0)
    P = fem["plan"].FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
    G, Kb, eta, c = fem["meshgen"].footing_materials(P.n_int)
    k_el = P.assemble_elastic(G, Kb)
    r = dp_return_map(fem["meshgen"].synthetic_strain(P.n_int), None, G, Kb, eta, c)
    k_tan, F = P.assemble_tangent_force(r["ds"], r["s"])
    mask = P.mask_u8(m["Q"])
    runs = []
    for use_graph in (True, False, True):
        M = mg.MultigridPCG(P, mask, use_graph=use_graph).setup(k_el)
        x, n_it, rel = M.solve(k_tan, -F, rtol=1e-10)
        runs.append((x.clone(), n_it, rel, [lv["S"].clone() for lv in M.lv], M.lmax0))
    for other in runs[1:]:
        assert other[1] == runs[0][1] and other[2] == runs[0][2] and other[4] == runs[0][4]
        assert all(torch.equal(a, b) for a, b in zip(runs[0][3], other[3]))
        assert torch.equal(runs[0][0], other[0])
