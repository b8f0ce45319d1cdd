"""Device driver of the reference's linear-elastic demo after mesh generation (SURVEY.md 8(f)-3;
Elasticity2D/pythonFEM.py:1100-1171): K = B^T D B, load vectors, Dirichlet lift f = f_t + f_V - K u_D, solve on the free
DOFs, stored energy 0.5 u'Ku - (f_t + f_V)'u.  Everything stays on the device; the reference's dense solve (:1157) is
replaced by the masked PCG (multigrid-preconditioned where the mesh is a uniform lattice, Jacobi otherwise)."""
import numpy as np
import torch

from . import loads
from . import pythonFEM as api
from .plan import FemPlan, axpby


def elasticity2d_driver(el_type, elements, coordinates, neumann_nodes, dirichlet_nodes, q_mask, shear, bulk,
                        volume_force=(0.0, -1.0), traction_force=(0.0, 450.0), rtol=1e-13, precond="jacobi", refine=1):
    """``elements`` (n_p, n_e) 0-based, ``neumann_nodes`` (n_p_s, n_e_s) 0-based surface elements, ``dirichlet_nodes`` /
    ``q_mask`` (2, n_n).  Returns dict: u (2, n_n) NumPy, energy, f_V, f_t (2, n_n) NumPy, iterations."""
    xi, wf = api.get_quadrature_volume(el_type)
    hatp, d1, d2 = api.get_local_basis_volume(el_type, xi)
    xs, ws = api.get_quadrature_surface(el_type)
    hs, ds = api.get_local_basis_surface(el_type, xs)
    P = FemPlan(np.asarray(elements).astype(np.int64), coordinates, d1, d2, wf)
    dev = P.device
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dt)  # noqa: E731
    k = P.assemble_elastic(shear * np.ones(P.n_int), bulk * np.ones(P.n_int))
    n_int_s = np.asarray(neumann_nodes).shape[1] * len(ws)
    f_v = loads.vector_volume_plan(P, np.dot(np.array([volume_force]).T, np.ones((1, P.n_int))), hatp)
    f_t = loads.vector_traction(t(neumann_nodes, torch.int64), P.coord, t(np.dot(np.array([traction_force]).T, np.ones((1, n_int_s)))),
                                t(hs), t(ds), t(ws))
    load = (f_t + f_v).t().reshape(-1).contiguous()                 # DOF-interleaved = flatten('F')
    ud = (0.5 * t(dirichlet_nodes)).t().reshape(-1).contiguous()
    f = axpby(1.0, load, -1.0, P.spmv(k, ud))                       # f = load - K u_D
    mask = P.mask_u8(q_mask)
    its = None
    if precond == "multigrid":
        from .mg import MultigridPCG
        x, its, _ = MultigridPCG(P, mask).setup(k).solve(k, f, rtol=rtol, maxit=2000)
    else:
        x, its, _ = P.pcg(k, f, mask, rtol=rtol, maxit=200000, refine=refine)
    u = torch.where(mask.bool(), x, ud)
    energy = 0.5 * float(torch.dot(u, P.spmv(k, u))) - float(torch.dot(load, u))
    return {"u": u.cpu().numpy().reshape((2, -1), order="F"), "energy": energy, "f_V": f_v.cpu().numpy(), "f_t": f_t.cpu().numpy(),
            "iterations": its, "plan": P}
