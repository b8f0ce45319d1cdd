"""SpMV grid-shape probe at 16M elements: persistent grid-stride vs launch-ordered one-pass grid, with/without the fused dot."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_elastoplasticity_b200 import _lib, meshgen, pythonFEM as api
from fem_elastoplasticity_b200.plan import FemPlan
def knob(k, v): _lib.call("fem_set_tuning", k.encode(), int(v))
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et); _, d1, d2 = api.get_local_basis_volume(et, xi)
m = meshgen.square_mesh_p1(2828, 2828)
P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
G, Kb, _, _ = meshgen.footing_materials(P.n_int)
k = P.assemble_elastic(G, Kb)
u = torch.randn(P.n_dof, dtype=torch.float64, device="cuda"); y = P.empty(P.n_dof)
mask = P.mask_u8(m["Q"]); dot = torch.zeros(1, dtype=torch.float64, device="cuda")
res = {}
for g in (4, 8):
    for un in (1, 2):
        for bps in (4, 16, 32, 128, 100000):
            knob("spmv_group", g); knob("spmv_unroll", un); knob("spmv_blocks_per_sm", bps)
            res[f"g{g}_u{un}_b{bps}_dot"] = timeit(lambda: P.spmv(k, u, mask=mask, out=y, dot=dot))
            res[f"g{g}_u{un}_b{bps}_nodot"] = timeit(lambda: P.spmv(k, u, mask=mask, out=y))
print(json.dumps(res, indent=1))
