"""Device-resident semismooth-Newton iteration and the two load-stepping drivers of the reference
(SURVEY.md 8(f)-1), all state in HBM:

    strain (K4) -> Drucker-Prager return map (K5) -> K_tangent + internal force (K3+K6, one pass)
      -> masked Jacobi-PCG on K_tangent[Q,Q] dU[Q] = -F[Q] (K7-K9) -> energy-norm criterion

Reference: Plasticity2D_DP/pythonFEM.py:1040-1087 (Newton), :986-1131 (footing load stepping);
tsx-tunnel/pythonFEM.py:1763-1830.  The dense LAPACK solve (:1066) is replaced by PCG driven to
``pcg_rtol``; everything else follows the reference statement by statement.  Host Python only sequences
kernel launches and reads back one scalar (the criterion) per iteration.

With a ``distributed.StripPartition`` the same loop runs one process per GPU: every rank keeps its strip of
the mesh plus one ghost cell row, so strain, return map, assembly and internal force need no communication and its owned
rows are complete; the linear solve is ``DistributedPCG`` (exchanges inside its kernels over NVLink peer memory) and the
criterion, plastic-point count and footing pressure are all-reduced, so every rank takes the same branches."""
import numpy as np
import torch

from .plan import FemPlan, axpby, dp_return_map


class NewtonSolver:
    def __init__(self, plan: FemPlan, shear, bulk, eta, c, q_mask, pcg_rtol=1e-13, pcg_maxit=200000, check_every=50,
                 tangent_mode="direct", precond="jacobi", coarse_cells=64, part=None, halo="auto", refine=0):
        self.plan = plan
        dev = plan.device
        f = plan._f64
        self.shear, self.bulk = f(shear, (plan.n_int,)), f(bulk, (plan.n_int,))
        self.eta, self.c = f(eta, (plan.n_int,)), f(c, (plan.n_int,))
        self.free = plan.mask_u8(q_mask)                   # free (non-Dirichlet) DOFs of the local vectors, ghosts included
        self.mask = self.free
        self.part = part if (part is not None and part.world > 1) else None
        self._dpcg = None
        if self.part is not None:                          # unknowns of this rank: free AND owned
            from .distributed import DistributedPCG
            self.mask = self.free & part.owned_mask(dev)
            self._dpcg = DistributedPCG(plan, part, self.mask, peer=halo)
        self.pcg_rtol, self.pcg_maxit, self.check_every = pcg_rtol, pcg_maxit, check_every
        self.refine = refine                               # iterative-refinement steps of the single-GPU Jacobi solve (FemPlan.pcg)
        self.tangent_mode = tangent_mode
        self.precond, self.coarse_cells, self._tl, self._mg = precond, coarse_cells, None, None   # "jacobi" | "twolevel" (twolevel.py) | "multigrid" (mg.py)
        self.k_elast = plan.assemble_elastic(self.shear, self.bulk)
        self.k_tan = plan.empty(plan.nnz)
        self.E = plan.empty(3, plan.n_int)
        self.F = plan.empty(plan.n_dof)
        self.rhs = plan.empty(plan.n_dof)
        self.work = plan.empty(4 * plan.n_dof)
        self.tmp = plan.empty(plan.n_dof)
        self.rm = {}
        self.zero_ep = torch.zeros((4, plan.n_int), dtype=torch.float64, device=dev)
        self.last = {}

    # -- building blocks -----------------------------------------------------------------------------
    def constitutive(self, u, ep_old, e0=None, apply=False):
        E = self.plan.strain(u, out=self.E)
        return dp_return_map(E, ep_old, self.shear, self.bulk, self.eta, self.c, apply_plastic_strain=apply, e0=e0,
                             want_ep=False, out=self.rm)

    def solve(self, k_vals, rhs, x0=None):
        if self.precond == "twolevel":
            if self._tl is None:                          # coarse operator of K_elast, kept for every tangent solve
                from .twolevel import TwoLevelPCG
                self._tl = TwoLevelPCG(self.plan, self.mask, nc=self.coarse_cells, part=self.part, free_mask=self.free).setup(self.k_elast)
            x, its, rel = self._tl.solve(k_vals, rhs, rtol=self.pcg_rtol, maxit=self.pcg_maxit, check_every=min(self.check_every, 10))
            return x.clone(), its, rel
        if self.precond == "multigrid":                   # V-cycle preconditioner; coarse operators of K_elast, kept (mg.py)
            if self._mg is None:
                from .mg import MultigridPCG
                self._mg = MultigridPCG(self.plan, self.mask, part=self.part, free_mask=self.free).setup(self.k_elast)
            x, its, rel = self._mg.solve(k_vals, rhs, rtol=self.pcg_rtol, maxit=min(self.pcg_maxit, 2000), check_every=2)
            return x.clone(), its, rel
        if self._dpcg is not None:                        # ghost rows of the returned vector are current
            x, its = self._dpcg.solve(k_vals, rhs, rtol=self.pcg_rtol, maxit=self.pcg_maxit, check_every=self.check_every)
            return x.clone(), its, self._dpcg.relres     # solve() raises PCGNotConverged when maxit is reached
        return self.plan.pcg(k_vals, rhs, self.mask, rtol=self.pcg_rtol, maxit=self.pcg_maxit, check_every=self.check_every,
                             x0=x0, work=self.work, refine=self.refine)

    def criterion(self, du, u_it, u_new):
        """q1/(q2+q3) with q = sqrt(v' K_elast v)   (Plasticity2D_DP/pythonFEM.py:1072-1075)"""
        if self._dpcg is not None:
            q = np.sqrt(self._dpcg.energy_norms(self.k_elast, du, u_it, u_new).cpu().numpy())
        else:
            q = np.sqrt(self.plan.energy_norms(self.k_elast, du, u_it, u_new, work=self.tmp).cpu().numpy())
        return float(q[0] / (q[1] + q[2]))

    def plastic_points(self, r):
        """Number of yielding integration points (the reference prints it per return-map call); owned elements only,
        summed over the ranks, when distributed."""
        if self.part is None:
            return int(r["counts"].sum().item())
        n_own = self.part.n_e_owned * self.plan.n_q
        n = (r["ind_p"][:n_own] != 0).sum().to(torch.int64).reshape(1)
        self.part.all_reduce(n)
        return int(n.item())

    def iteration(self, u_it, ep_old, e0=None):
        """One semismooth Newton iteration (:1043-1075). Returns (u_new, criterion, n_plastic, pcg_iterations)."""
        P = self.plan
        r = self.constitutive(u_it, ep_old, e0=e0)
        if self.tangent_mode == "reference":
            P.assemble_tangent_ref(r["ds"], self.shear, self.bulk, self.k_elast, out=self.k_tan)
            P.internal_force(r["s"], out=self.F)
        else:
            P.assemble_tangent_force(r["ds"], r["s"], out_k=self.k_tan, out_f=self.F)
        axpby(-1.0, self.F, 0.0, self.F, out=self.rhs)
        du, its, rel = self.solve(self.k_tan, self.rhs)
        u_new = axpby(1.0, u_it, 1.0, du)
        crit = self.criterion(du, u_it, u_new)
        n_plast = self.plastic_points(r)
        self.last = {"pcg_iters": its, "pcg_relres": rel, "n_plast": n_plast, "criterion": crit}
        return u_new, crit, n_plast, its


def footing_driver(mesh, plan=None, level=None, max_steps=1000, pcg_rtol=1e-13, tangent_mode="direct", log=None, precond="jacobi",
                   coarse_cells=64, part=None, halo="auto", refine=0):
    """Strip-footing load stepping of Plasticity2D_DP.elasticity_fem (:986-1131) on the device.
    ``mesh``: dict with coordinates (2,n_n), elements (3,n_e), Q, dirichlet_nodes (NumPy or CUDA tensors).
    ``part``: a distributed.StripPartition when run one process per GPU; ``mesh`` is then this rank's ``part.local_mesh``
    and the returned U / Ep are the local arrays (owned rows + ghosts)."""
    from . import pythonFEM as api
    from .meshgen import footing_materials
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    P = plan or FemPlan(mesh["elements"], mesh["coordinates"], d1, d2, wf)
    dev = P.device
    G, Kb, eta, c = footing_materials(P.n_int, dev)
    c0 = 450
    ns = NewtonSolver(P, G, Kb, eta, c, mesh["Q"], pcg_rtol=pcg_rtol, tangent_mode=tangent_mode, precond=precond, coarse_cells=coarse_cells,
                      part=part, halo=halo, refine=refine)
    dn = torch.as_tensor(np.asarray(mesh["dirichlet_nodes"].cpu() if isinstance(mesh["dirichlet_nodes"], torch.Tensor)
                                    else mesh["dirichlet_nodes"], dtype=np.float64)).to(dev)
    q_nd = dn[1] > 0
    if ns.part is not None:                                          # footing nodes this rank owns
        q_nd = q_nd & part.owned_mask(dev)[0::2].bool()
    dn_flat = dn.t().reshape(-1).contiguous()
    d_zeta = 1 / 1000
    d_zeta_min, d_zeta_old = d_zeta / 1300, d_zeta
    zeta_old, zeta_max = 0, 1
    ud = axpby(-d_zeta, dn_flat, 0.0, dn_flat)                       # :997
    f = P.spmv(ns.k_elast, ud)                                       # f = -K_elast Ud   (:998)
    axpby(-1.0, f, 0.0, f, out=f)
    sol, _, _ = ns.solve(ns.k_elast, f)
    mk = ns.free.bool()
    u_it = torch.where(mk, sol, ud)                                  # :1004 (free DOFs overwritten)
    U = torch.zeros_like(u_it)
    u_old = axpby(-1.0, u_it, 0.0, u_it)                             # :1009
    ep_old = torch.zeros((4, P.n_int), dtype=torch.float64, device=dev)
    pressure_old = 0.0
    trace, hist, step = [], [], 1
    while step <= max_steps:
        zeta = zeta_old + d_zeta
        criterion = np.inf
        for it in range(25):
            u_new, criterion, n_plast, pits = ns.iteration(u_it, ep_old)
            trace.append((zeta, it, n_plast, criterion, pits))
            if log:
                log(f"zeta={zeta:.6g} it={it} plastic={n_plast} criterion={criterion:.3e} pcg={pits}")
            if np.isnan(criterion):
                break
            u_it = u_new
            if criterion < 1e-12:
                break
        if criterion < 1e-10:                                         # :1091-1112
            u_old, U = U, u_it
            r = ns.constitutive(U, ep_old, apply=True)                # Ep_old updated in place
            zeta_old, d_zeta_old = zeta, d_zeta
            step += 1
            pa = P.transform(r["s"][1])
            if ns.part is None:
                pressure = float((-pa[q_nd].mean() / c0).item())
            else:                                                     # mean over the footing nodes of all ranks
                acc = torch.stack([pa[q_nd].sum(), q_nd.sum().to(torch.float64)])
                part.all_reduce(acc)
                pressure = float((-(acc[0] / acc[1]) / c0).item())
            hist.append((zeta, pressure))
            if pressure - pressure_old < 0.1 and criterion < 1e-12:
                d_zeta *= 2
            pressure_old = pressure
        else:
            d_zeta /= 2
        du_ = axpby(1.0, U, -1.0, u_old)
        u_it = axpby(d_zeta / d_zeta_old, du_, 1.0, U)                # :1120
        if zeta_old >= zeta_max or d_zeta < d_zeta_min:
            break
    return {"U": U.cpu().numpy().reshape((2, -1), order="F"), "steps": step, "trace": trace, "hist": hist,
            "Ep": ep_old.cpu().numpy()}


def tsx_driver(coords, elem, max_steps=100, pcg_rtol=1e-13, tangent_mode="direct", log=None, precond="jacobi", coarse_cells=8, refine=0,
               part=None):
    """tsx-tunnel load stepping (tsx-tunnel/pythonFEM.py:1661-1830) for P1 on the device.  ``part``: a
    partition.GeneralPartition of the (global) mesh when run one process per GPU; the returned U is then the local array
    (owned nodes first, ghosts after)."""
    from . import pythonFEM as api
    if part is not None and part.world > 1:
        lm = part.local_mesh({"coordinates": coords, "elements": elem}, "cpu")
        coords, elem = lm["coordinates"].numpy(), lm["elements"].numpy()
    else:
        part = None
    et = api.LagrangeElementType.P1
    xi, wf = api.get_quadrature_volume(et)
    _, d1, d2 = api.get_local_basis_volume(et, xi)
    young, poisson = 60000, 0.2                                       # :1663-1681
    shear0 = young / (2 * (1 + poisson))
    bulk0 = young / (3 * (1 - 2 * poisson))
    cohesion, phi = 18.7, 49 * np.pi / 180
    eta0 = 3 * np.tan(phi) / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    c0 = 3 * cohesion / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    s0 = np.array([-45.0, -11.0, 0.0, -60.0])
    tr = s0[0] + s0[1] + s0[3]
    e0 = np.array([-poisson * tr + (1 + poisson) * s0[0], -poisson * tr + (1 + poisson) * s0[1], 0,
                   -poisson * tr + (1 + poisson) * s0[3]], dtype=float) / young
    q = np.ones(coords.shape, dtype=bool)                             # :1695-1699
    q[0, coords[0] < -49.99] = 0
    q[0, coords[0] > 49.99] = 0
    q[1, coords[1] < -49.99] = 0
    q[1, coords[1] > 49.99] = 0
    P = FemPlan(elem, coords, d1, d2, wf)
    n_int = P.n_int
    ones = np.ones(n_int)
    ns = NewtonSolver(P, shear0 * ones, bulk0 * ones, eta0 * ones, c0 * ones, q, pcg_rtol=pcg_rtol, tangent_mode=tangent_mode,
                      precond=precond, coarse_cells=coarse_cells, refine=refine, part=part)
    s_init = torch.as_tensor(np.tile(s0.reshape(-1, 1), (1, n_int))).to(P.device)
    f0 = P.internal_force(s_init)                                     # :1737
    rhs = axpby(-1.0, f0, 0.0, f0)
    u_elast, _, _ = ns.solve(ns.k_elast, rhs)                         # :1748
    d_zeta = 1 / 17
    d_zeta_min, d_zeta_old = d_zeta / 10, d_zeta
    zeta_old, zeta_max = 0, 1
    u_it = axpby(d_zeta, u_elast, 0.0, u_elast)
    U = torch.zeros_like(u_it)
    u_old = axpby(-1.0, u_it, 0.0, u_it)
    ep_old = torch.zeros((4, n_int), dtype=torch.float64, device=P.device)
    trace, step = [], 0
    while step < max_steps:
        zeta = zeta_old + d_zeta
        e0z = zeta * e0
        criterion = np.inf
        for it in range(25):
            u_new, criterion, n_plast, pits = ns.iteration(u_it, ep_old, e0=e0z)
            trace.append((zeta, it, n_plast, criterion, pits))
            if log:
                log(f"zeta={zeta:.6g} it={it} plastic={n_plast} criterion={criterion:.3e} pcg={pits}")
            if np.isnan(criterion):
                break
            u_it = u_new
            if criterion < 1e-12:
                break
        if criterion < 1e-10:
            u_old, U = U, u_it
            ns.constitutive(U, ep_old, e0=e0z)                        # :1808: no plastic-strain update (SURVEY B-5)
            zeta_old, d_zeta_old = zeta, d_zeta
            step += 1
        else:
            d_zeta = d_zeta / 2
        du_ = axpby(1.0, U, -1.0, u_old)
        u_it = axpby(d_zeta / d_zeta_old, du_, 1.0, U)
        if zeta_old >= zeta_max or d_zeta < d_zeta_min:
            break
    return {"U": U.cpu().numpy().reshape((2, -1), order="F"), "steps": step, "trace": trace, "F0": f0.cpu().numpy(),
            "Q": q, "plan": P, "solver": ns}
