"""2-GPU diagnostic: distributed PCG (nccl / peer halo) vs single-GPU, true residuals."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_elastoplasticity_b200 import meshgen, pythonFEM as api
from fem_elastoplasticity_b200.distributed import DistributedPCG, StripPartition
from fem_elastoplasticity_b200.plan import FemPlan
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
nx, ny = 96, 128
et = api.LagrangeElementType.P1
xi, wf = api.get_quadrature_volume(et); _, d1, d2 = api.get_local_basis_volume(et, xi)
part = StripPartition(nx, ny, rank, world, 10.0, 10.0 * world)
mesh = part.local_mesh(dev)
P = FemPlan(mesh["elements"], mesh["coordinates"], d1, d2, wf, device=dev)
G, Kb, _, _ = meshgen.footing_materials(P.n_int, dev)
k = P.assemble_elastic(G, Kb)
mask = part.free_owned_mask(P, mesh)
b_global = np.random.default_rng(9).standard_normal(2 * (nx + 1) * (ny + 1))
lo = part.iy0 * part.row_dofs
rhs = torch.as_tensor(b_global[lo:lo + P.n_dof].copy()).to(dev)
# global single-GPU reference on every rank
m = meshgen.square_mesh_p1(nx, ny, 10.0, 10.0 * world, device=dev)
Pg = FemPlan(m["elements"], m["coordinates"], d1, d2, wf, device=dev)
Gg, Kg, _, _ = meshgen.footing_materials(Pg.n_int, dev)
kg = Pg.assemble_elastic(Gg, Kg)
mg = Pg.mask_u8(m["Q"])
bg = torch.as_tensor(b_global).to(dev)
ref, its, rel = Pg.pcg(kg, bg, mg, rtol=1e-12, maxit=20000, check_every=25)
print(rank, "single: its", its, "rel", rel, flush=True)
# owned rows of local K equal the global rows?
a, e = part.owned_dof_range()
xg = torch.randn(Pg.n_dof, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
yg = Pg.spmv(kg, xg)
yl = P.spmv(k, xg[lo:lo + P.n_dof].contiguous())
print(rank, "local-vs-global spmv on owned rows:", float((yl[a:e] - yg[lo + a:lo + e]).abs().max()), float(yg.abs().max()), flush=True)
print(rank, "mask equal on owned:", bool(torch.equal(mask[a:e], mg[lo + a:lo + e])), "mask sum", int(mask.sum()), flush=True)
for name, peer in (("nccl", False), ("peer", True)):
    pcg = DistributedPCG(P, part, mask, peer=peer)
    x, it2 = pcg.solve(k, rhs.clone(), rtol=1e-12, maxit=20000, check_every=25)
    h = pcg.scal.cpu().numpy()
    err = float((x[a:e] - ref[lo + a:lo + e]).abs().max() / ref.abs().max())
    print(rank, name, "peer_active", pcg.peer is not None, "its", it2, "scal", h[:5], "err vs single", err, flush=True)
dist.destroy_process_group()
