"""Minimal driver for ncu: builds the 16M-element plan and launches each hot kernel a few times.
    python tools/prof_kernels.py [--nx 2828] [--only assemble|spmv|return_map|all]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_elastoplasticity_b200 import meshgen  # noqa: E402
from fem_elastoplasticity_b200 import pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=2828)
ap.add_argument("--only", default="all")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et)
_, d1, d2 = api.get_local_basis_volume(et, xi)
m = meshgen.square_mesh_p1(a.nx, a.nx)
P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
G, Kb, eta, c = meshgen.footing_materials(P.n_int)
Es = meshgen.synthetic_strain(P.n_int)
ep = torch.zeros((4, P.n_int), dtype=torch.float64, device="cuda")
rm = {}
k, F = P.empty(P.nnz), P.empty(P.n_dof)
u = torch.randn(P.n_dof, dtype=torch.float64, device="cuda")
y = P.empty(P.n_dof)
mask = P.mask_u8(m["Q"])
dot = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(a.reps):
    if a.only in ("all", "return_map"):
        dp_return_map(Es, ep, G, Kb, eta, c, want_ep=False, out=rm)
    else:
        dp_return_map(Es, ep, G, Kb, eta, c, want_ep=False, out=rm) if not rm else None
    if a.only in ("all", "assemble", "elastic"):
        P.assemble_elastic(G, Kb, out=k)
    if a.only in ("all", "assemble", "tangent_force", "tf_spmv"):
        P.assemble_tangent_force(rm["ds"], rm["s"], out_k=k, out_f=F)
    if a.only in ("all", "spmv", "tf_spmv"):
        P.spmv(k, u, mask=mask, out=y, dot=dot)
    if a.only in ("all", "strain"):
        P.strain(u)
torch.cuda.synchronize()
print("ok", P.n_e)
