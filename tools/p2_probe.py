"""P2 (6-node triangles, 7-point rule, 12 x 12 local matrices) at scale: enrich a uniform P1 mesh on the GPU
(meshgen.create_midpoints_p2), build the plan, time K_elast / K_tangent assembly, the return map and the SpMV, and report
them against the HBM roof with the general algorithmic-byte formula of SURVEY 8(d).  python tools/p2_probe.py [nx]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from fem_elastoplasticity_b200 import meshgen, pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 700
m = meshgen.square_mesh_p1(nx, nx)
torch.cuda.synchronize()
t0 = time.perf_counter()
d = meshgen.create_midpoints_p2(m["coordinates"], m["elements"].to(torch.int64))
torch.cuda.synchronize()
t_mid = time.perf_counter() - t0
et = api.LagrangeElementType.P2
xi, wf = api.get_quadrature_volume(et)
_, d1, d2 = api.get_local_basis_volume(et, xi)
t0 = time.perf_counter()
P = FemPlan(d["elem_ext"].to(torch.int32), d["coord_ext"], d1, d2, wf)
torch.cuda.synchronize()
t_plan = time.perf_counter() - t0
G, Kb, eta, c = meshgen.footing_materials(P.n_int)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


k = P.empty(P.nnz)
u = torch.randn(P.n_dof, dtype=torch.float64, device="cuda") * 1e-4
E = P.strain(u)
r = dp_return_map(E, None, G, Kb, eta, c)
peak = 6532.2
n_q, n_p = P.n_q, P.n_p
bytes_el = (4 * n_p + 16 * P.n_n / P.n_e + n_q * (16 + 8) + 8 * P.nnz / P.n_e) * P.n_e          # shear, bulk per point + weight out
bytes_tan = (4 * n_p + 16 * P.n_n / P.n_e + n_q * (72 + 8) + 8 * P.nnz / P.n_e) * P.n_e
res = {"nx": nx, "n_e": P.n_e, "n_n": P.n_n, "n_int": P.n_int, "nnz": P.nnz, "max_degree": P.max_degree, "midpoints_s": t_mid, "plan_s": t_plan}
for name, fn, nb in (("assemble_elastic", lambda: P.assemble_elastic(G, Kb, out=k), bytes_el),
                     ("assemble_tangent", lambda: P.assemble_tangent(r["ds"], out=k), bytes_tan),
                     ("return_map", lambda: dp_return_map(E, None, G, Kb, eta, c, want_ep=False, out=r), 193.0 * P.n_int),
                     ("strain", lambda: P.strain(u), (4 * n_p + 24 * n_q) * P.n_e + 8.0 * P.n_dof),
                     ("spmv", lambda: P.spmv(k, u), 12.0 * P.nnz + 20.0 * P.n_dof)):
    ms = timeit(fn)
    res[name] = {"ms": ms, "GBs": nb / ms / 1e6, "frac_of_measured_peak": nb / ms / 1e6 / peak, "per_unit": (P.n_e if "assemble" in name else P.n_int) / ms / 1e3}
print(json.dumps(res, indent=1))
