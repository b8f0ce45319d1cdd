"""A/B of tuning settings on the multigrid-CG iteration of the bench's 16M-element system:
    python tools/mg_ab.py key=value [key=value ...]   ->  ms per iteration and iteration count to rtol 1e-10 per setting"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from fem_elastoplasticity_b200 import _lib, meshgen, mg, pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map  # noqa: E402

nx = 2828
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
M0 = mg.MultigridPCG(P, mask).setup(k_el)
for rnd in range(2):
    for setting in [a for a in sys.argv[1:] if "=" in a]:
        key, val = setting.split("=")
        if key.startswith("ctor:"):                      # constructor argument of MultigridPCG, e.g. ctor:max_coarse_dofs=4500
            M = mg.MultigridPCG(P, mask, **{key[5:]: int(val)}).setup(k_el)
            key = "mg_stencil_sym"
            val = 0
        else:
            M = M0
        _lib.call("fem_set_tuning", key.encode(), int(val))
        M._graph = None
        x, its, rel = M.solve(k_tan, -F, rtol=1e-10)
        M.solve(k_tan, -F, iters=40)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        M.solve(k_tan, -F, iters=40)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        _lib.call("fem_set_tuning", key.encode(), 0)
        print(json.dumps({"setting": setting, "round": rnd, "iterations_to_1e-10": its, "relres": rel, "ms_per_iteration": dt * 1e3 / 40}), flush=True)
