"""A/B of assembly kernel builds: FEM_B200_LIB=<lib> python tools/ab_assemble.py [nx] [key=value ...]  ->  one line per
setting (none: the defaults; each key=value is one fem_set_tuning setting, measured in turn, two rounds) with the isolated
times of the fused tangent+force assembly, tangent only and elastic at 16M elements, and a bit-equality check against variant B."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from fem_elastoplasticity_b200 import _lib, meshgen, pythonFEM as api  # noqa: E402
from fem_elastoplasticity_b200.plan import FemPlan, dp_return_map  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2828
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et)
_, d1, d2 = api.get_local_basis_volume(et, xi)
m = meshgen.square_mesh_p1(nx, nx)
P = FemPlan(m["elements"], m["coordinates"], d1, d2, wf)
G, Kb, eta, c = meshgen.footing_materials(P.n_int)
r = dp_return_map(meshgen.synthetic_strain_global(P.n_int, 0), None, G, Kb, eta, c)
k, F = P.empty(P.nnz), P.empty(P.n_dof)


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


settings = [a for a in sys.argv[2:] if "=" in a] or [None]
for rnd in range(2 if settings != [None] else 1):
    for setting in settings:
        if setting:
            key, val = setting.split("=")
            _lib.call("fem_set_tuning", key.encode(), int(val))
        out = {"lib": os.path.basename(_lib.LIB_PATH), "setting": setting, "round": rnd}
        out["tangent_force_ms"] = timeit(lambda: P.assemble_tangent_force(r["ds"], r["s"], out_k=k, out_f=F))
        kd, Fd = k.clone(), F.clone()
        out["tangent_ms"] = timeit(lambda: P.assemble_tangent(r["ds"], out=k))
        out["elastic_ms"] = timeit(lambda: P.assemble_elastic(G, Kb, out=k))
        kel = k.clone()
        k2 = P.empty(P.nnz)
        out["tangent_ref_ms"] = timeit(lambda: P.assemble_tangent_ref(r["ds"], G, Kb, kel, out=k2))
        kr = k2.clone()
        _lib.call("fem_set_tuning", b"assemble_variant", 2)
        kb, Fb = P.assemble_tangent_force(r["ds"], r["s"])
        krb = P.assemble_tangent_ref(r["ds"], G, Kb, kel)
        _lib.call("fem_set_tuning", b"assemble_variant", 0)
        out["equals_variant_B_bits"] = bool(torch.equal(kd, kb) and torch.equal(Fd, Fb) and torch.equal(kr, krb))
        if setting:
            _lib.call("fem_set_tuning", setting.split("=")[0].encode(), 0)
        print(json.dumps(out), flush=True)
