"""Geometric multigrid PCG on the bench's synthetic Newton system (config 4): set-up time, iterations to rtol 1e-10 and
time for a few smoother settings.  python tools/mg_probe.py [nx] [degree:ratio ...]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from fem_elastoplasticity_b200 import meshgen, mg, pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2828
combos = [tuple(float(v) for v in a.split(":")) for a in sys.argv[2:]] or [(3, 8.0), (2, 4.0), (2, 8.0), (4, 16.0)]
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et)
_, d1, d2 = api.get_local_basis_volume(et, xi)
m = meshgen.square_mesh_p1(nx, nx)
P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
G, Kb, eta, c = meshgen.footing_materials(P.n_int)
k_el = P.assemble_elastic(G, Kb)
r = dp_return_map(meshgen.synthetic_strain_global(P.n_int, 0), None, G, Kb, eta, c)
k_tan, F = P.assemble_tangent_force(r["ds"], r["s"])
mask = P.mask_u8(m["Q"])
rhs = -F
out = {"nx": nx, "n_dof": P.n_dof, "runs": []}
for deg, ratio in combos:
    M = mg.MultigridPCG(P, mask, degree=int(deg), ratio=ratio).setup(k_el)
    res = {"degree": int(deg), "ratio": ratio, "levels": M.n_levels, "setup_s": M.setup_seconds, "lmax0": M.lmax0,
           "lmax": [lv["lmax"] for lv in M.lv[:-1]]}
    for tag, k in (("tangent", k_tan), ("elastic", k_el)):
        M.solve(k, rhs, rtol=1e-10)                      # warm-up (graph capture)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, its, rel = M.solve(k, rhs, rtol=1e-10)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        true = float(((rhs - P.spmv(k, x)) * mask).norm() / (rhs * mask).norm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        M.solve(k, rhs, iters=20)
        e1.record()
        torch.cuda.synchronize()
        res[tag] = {"iterations": its, "relres": rel, "true_relres": true, "seconds": dt, "ms_per_iteration_fixed20": e0.elapsed_time(e1) / 20}
    out["runs"].append(res)
    del M
print(json.dumps(out, indent=1))
