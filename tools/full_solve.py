"""One converged Jacobi-PCG solve at config-4 size: the footing's initial elastic solve K_elast[Q,Q] u = -K_elast Ud
(Plasticity2D_DP/pythonFEM.py:997-1004) on the 2828 x 2828 mesh, rtol 1e-10.  Reports iterations and wall time, i.e. what
'PCG-Newton s/step' costs when the inner solve is driven to convergence (Jacobi iterations grow ~ N)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_elastoplasticity_b200 import meshgen, pythonFEM as api
from fem_elastoplasticity_b200.plan import FemPlan, axpby
ap = argparse.ArgumentParser(); ap.add_argument("--nx", type=int, default=2828); ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--maxit", type=int, default=400000); ap.add_argument("--precond", default="jacobi", choices=["jacobi", "twolevel"])
ap.add_argument("--nc", type=int, default=64); a = ap.parse_args()
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et); _, d1, d2 = api.get_local_basis_volume(et, xi)
m = meshgen.square_mesh_p1(a.nx, a.nx)
P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
G, Kb, _, _ = meshgen.footing_materials(P.n_int)
k = P.assemble_elastic(G, Kb)
ud = (-1e-3 * m["dirichlet_nodes"]).t().reshape(-1).contiguous()
f = P.spmv(k, ud); axpby(-1.0, f, 0.0, f, out=f)
mask = P.mask_u8(m["Q"])
extra = {}
if a.precond == "twolevel":
    from fem_elastoplasticity_b200.twolevel import TwoLevelPCG
    tl = TwoLevelPCG(P, mask, nc=a.nc).setup(k)
    extra = {"coarse_dofs": tl.ncd, "coarse_grid": tl.grid[4:], "setup_seconds": tl.setup_seconds}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, its, rel = tl.solve(k, f, rtol=a.rtol, maxit=a.maxit, check_every=50)
else:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, its, rel = P.pcg(k, f, mask, rtol=a.rtol, maxit=a.maxit, check_every=500, raise_on_maxit=False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"precond": a.precond, **extra, "n_e": P.n_e, "n_dof": P.n_dof, "free_dof": int(mask.sum().item()), "rtol": a.rtol, "iterations": its, "relres": rel,
                  "seconds": dt, "ms_per_iteration": 1e3 * dt / max(its, 1)}))
