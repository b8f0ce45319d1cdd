"""A few multigrid-PCG iterations on the bench's 16M-element system, launched eagerly (no CUDA graph), for the ncu launch
list / full captures of the V-cycle kernels:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mg_launches.csv python tools/prof_mg.py"""
import sys

import torch

sys.path.insert(0, ".")
from fem_elastoplasticity_b200 import meshgen, mg, pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2828
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
for setting in [a for a in sys.argv[3:] if "=" in a]:          # tuning settings, e.g. mg_stencil_sym=2
    from fem_elastoplasticity_b200 import _lib
    _lib.call("fem_set_tuning", setting.split("=")[0].encode(), int(setting.split("=")[1]))
M = mg.MultigridPCG(P, mask, use_graph=False).setup(k_el)
torch.cuda.synchronize()
M.solve(k_tan, -F, iters=iters)
torch.cuda.synchronize()
print("ok", P.n_e, M.n_levels)
