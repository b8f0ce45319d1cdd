"""Load vectors of the reference's linear-elastic demo (Elasticity2D/pythonFEM.py:246-364) as
device-agnostic torch code: they are O(n) pre-processing on either side of the hot path, evaluated once per run.

The reference builds COO triplets and lets SciPy sum the duplicates; SciPy adds the contributions of a node in input
order (ascending integration point, then local node).  ``_scatter_ordered`` reproduces exactly that order on any device -
the k-th contribution of every node is added in round k, so no two additions of a round hit the same node and no atomics
are involved - which makes the result bit-identical to the reference's and reproducible from run to run."""
import torch


def _scatter_ordered(vals, nodes, n_n):
    order = torch.argsort(nodes, stable=True)
    sn, sv = nodes[order], vals[order]
    counts = torch.bincount(sn, minlength=n_n)
    if vals.is_cuda:          # one thread per node adds its (stably sorted) contributions in order: fem_segment_sum_ordered
        from ._lib import call
        from .plan import _ptr, _stream
        ptr = torch.zeros(n_n + 1, dtype=torch.int64, device=vals.device)
        ptr[1:] = torch.cumsum(counts, 0)
        sv = sv.contiguous()
        out = torch.empty(n_n, dtype=torch.float64, device=vals.device)
        with torch.cuda.device(vals.device):
            call("fem_segment_sum_ordered", n_n, _ptr(ptr), _ptr(sv), _ptr(out), _stream())
        return out
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(sn.numel(), device=sn.device) - starts[sn]
    out = torch.zeros(n_n, dtype=vals.dtype, device=vals.device)
    for k in range(int(counts.max()) if sn.numel() else 0):
        sel = rank == k
        idx = sn[sel]
        out[idx] = out[idx] + sv[sel]
    return out


def vector_volume(elements, n_n, f_v_int, hatp, weight):
    """f_V (2, n_n): sum over integration points g = e*n_q + q and local nodes a of hatp[a, q] * (weight[g] * f[c, g])
    at node elements[a, e]  (:281-290).  elements (n_p, n_e) 0-based, f_v_int (2, n_int), hatp (n_p, n_q), weight (n_int,)."""
    elements = elements.to(torch.int64)
    n_p, n_e = elements.shape
    n_q = hatp.shape[1]
    hatphi = hatp.repeat(1, n_e)                                           # (n_p, n_int): column g holds hatp[:, g % n_q]
    nodes = elements.repeat_interleave(n_q, dim=1).t().reshape(-1)         # flatten('F'): g-major, local node minor
    out = []
    for c in range(2):
        v = (hatphi * (weight.reshape(1, -1) * f_v_int[c].reshape(1, -1))).t().reshape(-1)
        out.append(_scatter_ordered(v, nodes, n_n))
    return torch.stack(out)


def vector_traction(elements_s, coordinates, f_t_int, hatp_s, dhatp1_s, wf_s):
    """f_t (2, n_n) (:327-362).  Kept from the reference: the surface Jacobian is |sum_a x_a dhatp1_s[a, q]| (x coordinate
    only: the loaded side is horizontal) and the load is f_t_int[c, -1], the last integration point's value, everywhere."""
    elements_s = elements_s.to(torch.int64)
    n_n = coordinates.shape[1]
    n_p_s, n_e_s = elements_s.shape
    n_q_s = wf_s.numel()
    dhatphi1_s = dhatp1_s.repeat(1, n_e_s)                                  # (n_p_s, n_q_s) tiled over the surface elements
    hatphi_s = hatp_s.repeat(1, n_e_s)
    coord_int1 = coordinates[0][elements_s].repeat_interleave(n_q_s, dim=1)  # (n_p_s, n_int_s)
    prod = coord_int1 * dhatphi1_s
    j11 = prod[0].clone()
    for a in range(1, n_p_s):                                               # builtin sum(): rows added in order
        j11 = j11 + prod[a]
    weight_s = j11.abs() * wf_s.to(coordinates.dtype).repeat(n_e_s)
    nodes = elements_s.repeat_interleave(n_q_s, dim=1).t().reshape(-1)
    out = []
    for c in range(2):
        v = (hatphi_s * (weight_s * f_t_int[c, -1]).reshape(1, -1)).t().reshape(-1)
        out.append(_scatter_ordered(v, nodes, n_n))
    return torch.stack(out)


def vector_volume_plan(plan, f_v_int, hatp):
    """f_V (2, n_n) on the device through the plan's node -> element incidence lists (fem_vector_volume): the same
    ascending-integration-point order as ``vector_volume``, no sort, no atomics."""
    import ctypes as C

    import numpy as np

    from ._lib import call
    from .plan import _ptr, _stream
    f = plan._f64(f_v_int, (2, plan.n_int))
    h = np.ascontiguousarray(np.asarray(hatp, dtype=np.float64).reshape(plan.n_p, plan.n_q))
    out = torch.empty((2, plan.n_n), dtype=torch.float64, device=plan.device)
    with torch.cuda.device(plan.device):
        call("fem_vector_volume", plan._h, _ptr(f), h.ctypes.data_as(C.POINTER(C.c_double)), _ptr(out), _stream())
    return out
