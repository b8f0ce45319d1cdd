"""Isolated timings of the streaming level-0 kernels at 16M elements (FP64 SpMV + dot, multigrid Chebyshev step and residual
step on the FP32 copy): FEM_B200_LIB=<lib> python tools/stream_probe.py  ->  one JSON line (A/B of kernel builds)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from fem_elastoplasticity_b200 import _lib, meshgen, mg, pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2828
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et)
_, d1, d2 = api.get_local_basis_volume(et, xi)
m = meshgen.square_mesh_p1(nx, nx)
P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
G, Kb, _, _ = meshgen.footing_materials(P.n_int)
k = P.assemble_elastic(G, Kb)
mask = P.mask_u8(m["Q"])
M = mg.MultigridPCG(P, mask, max_coarse_dofs=2500).setup(k)
M.block_jacobi(k)
_lib.call("fem_mg_to_f32", P.nnz, k.data_ptr(), M.k32.data_ptr(), torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1)
b = torch.randn(P.n_dof, dtype=torch.float64, device="cuda", generator=g) * mask
x = torch.randn(P.n_dof, dtype=torch.float64, device="cuda", generator=g) * mask
out, y = torch.zeros_like(x), torch.zeros_like(x)
dot = torch.zeros(1, dtype=torch.float64, device="cuda")


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps


res = {"lib": os.path.basename(_lib.LIB_PATH)}
for rnd in range(2):
    res[f"cheb_ms_{rnd}"] = timeit(lambda: M.fine_step(k, b, x, out, mode=2, step=1))
    res[f"resid_ms_{rnd}"] = timeit(lambda: M.fine_step(k, b, x, out, mode=1))
    res[f"spmv_dot_ms_{rnd}"] = timeit(lambda: P.spmv(k, x, mask=mask, out=y, dot=dot))
ref = (b - P.spmv(k, x)) * mask
M.fine_step(k, b, x, out, mode=1)
res["resid_err"] = float((out - ref).abs().max() / ref.abs().max())
print(json.dumps(res))
