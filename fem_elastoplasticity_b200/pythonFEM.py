"""Drop-in host API: the reference's ``pythonFEM.py`` names and signatures for the hot path, backed by
the CUDA kernels.  NumPy / SciPy objects in, NumPy / SciPy objects out (same shapes, layouts and
in-place side effects as the reference); every number is produced on the GPU.

Reference functions mirrored
  get_elastic_stiffness_matrix   Plasticity2D_DP/pythonFEM.py:491-601 (== tsx-tunnel:432-542); Elasticity2D:368-477
  construct_constitutive_problem Plasticity2D_DP/pythonFEM.py:604-757; tsx-tunnel/pythonFEM.py:990-1157
  get_quadrature_volume / get_local_basis_volume / LagrangeElementType   :55-60, :364-488 (host tables)
Entry points for the statements the reference writes inline in its Newton loop (:1043-1075)
  strain, assemble_tangent, internal_force, solve_increment (PCG instead of the dense solve), stopping_criterion
"""
import enum

import numpy as np
import scipy.sparse as ssp
import torch

from .plan import FemPlan, dp_return_map


class LagrangeElementType(enum.Enum):
    P1 = 1
    P2 = 2
    Q1 = 3
    Q2 = 4


def flatten_row(v):
    return np.reshape(v, (1, -1), order='F')


def flatten_col(v):
    return np.reshape(v, (np.size(v), 1), order='F')


def get_quadrature_volume(el_type):
    """(Xi (2,n_q), WF (1,n_q)) - same rules as the reference (:398-410)."""
    g = 1 / np.sqrt(3)
    third = 1 / 3
    if el_type == LagrangeElementType.P1:
        return np.array([[third], [third]]), np.array([[0.5]])
    if el_type == LagrangeElementType.P2:
        p, q, r, s = 0.1012865073235, 0.7974269853531, 0.4701420641051, 0.0597158717898
        return (np.array([[p, q, p, r, r, s, third], [p, p, q, s, r, r, third]]),
                0.5 * np.array([[0.1259391805448] * 3 + [0.1323941527885] * 3 + [0.225]]))
    if el_type == LagrangeElementType.Q1:
        return np.array([[-g, -g, g, g], [-g, g, -g, g]]), np.array([[1, 1, 1, 1]])
    if el_type == LagrangeElementType.Q2:
        return (np.array([[-g, g, g, -g, 0, g, 0, -g, 0], [-g, -g, g, g, -g, 0, g, 0, 0]]),
                np.array([[25 / 81] * 4 + [40 / 81] * 4 + [64 / 81]]))
    raise ValueError(f"unsupported element type {el_type}")


def get_local_basis_volume(el_type, xi):
    """(HatP, DHatP1, DHatP2), each (n_p, n_q) (:434-488)."""
    x, y = xi[0], xi[1]
    zero = np.zeros(np.size(x, 0))
    if el_type == LagrangeElementType.P1:
        return np.array([1 - x - y, x, y]), np.array([[-1], [1], [0]]), np.array([[-1], [0], [1]])
    if el_type == LagrangeElementType.P2:
        o = 1 - x - y
        return (np.array([o * (2 * o - 1), x * (2 * x - 1), y * (2 * y - 1), 4 * x * y, 4 * o * y, 4 * o * x]),
                np.array([-4 * o + 1, 4 * x - 1, zero, 4 * y, -4 * y, 4 * (o - x)]),
                np.array([-4 * o + 1, zero, 4 * y - 1, 4 * x, 4 * (o - y), -4 * x]))
    if el_type == LagrangeElementType.Q1:
        return (np.array([(1 - x) * (1 - y) / 4, (1 + x) * (1 - y) / 4, (1 + x) * (1 + y) / 4, (1 - x) * (1 + y) / 4]),
                np.array([-(1 - y) / 4, (1 - y) / 4, (1 + y) / 4, -(1 + y) / 4]),
                np.array([-(1 - x) / 4, -(1 + x) / 4, (1 + x) / 4, (1 - x) / 4]))
    if el_type == LagrangeElementType.Q2:
        xx, yy = pow(x, 2), pow(y, 2)
        return (np.array([(1 - x) * (1 - y) * (-1 - x - y) / 4, (1 + x) * (1 - y) * (-1 + x - y) / 4,
                          (1 + x) * (1 + y) * (-1 + x + y) / 4, (1 - x) * (1 + y) * (-1 - x + y) / 4,
                          (1 - xx) * (1 - y) / 2, (1 + x) * (1 - yy) / 2, (1 - xx) * (1 + y) / 2, (1 - x) * (1 - yy) / 2]),
                np.array([(1 - y) * (2 * x + y) / 4, (1 - y) * (2 * x - y) / 4, (1 + y) * (2 * x + y) / 4,
                          (1 + y) * (2 * x - y) / 4, -x * (1 - y), (1 - yy) / 2, -x * (1 + y), -(1 - yy) / 2]),
                np.array([(1 - x) * (x + 2 * y) / 4, (1 + x) * (-x + 2 * y) / 4, (1 + x) * (x + 2 * y) / 4,
                          (1 - x) * (-x + 2 * y) / 4, -(1 - xx) / 2, -(1 + x) * y, (1 - xx) / 2, -(1 - x) * y]))
    raise ValueError(f"unsupported element type {el_type}")


# ------------------------------------------------------------------------------------------------
_PIN_MIN_BYTES = 1 << 20


def _to_numpy(t):
    """Device tensor -> NumPy array.  Arrays of a megabyte or more are downloaded straight into page-locked memory from
    torch's caching host allocator and handed out as the NumPy array itself (it keeps the block alive and returns it to the
    cache when it is freed): the copy runs at the PCIe rate instead of the pageable-memory rate, and when the caller passes
    the array back into the next call of this module (E -> construct_constitutive_problem, ds -> assemble_tangent,
    s -> internal_force) the upload does too - the CUDA runtime recognises page-locked source memory by address."""
    if t.numel() * t.element_size() < _PIN_MIN_BYTES:
        return t.cpu().numpy()
    try:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:                       # no page-locked memory left on this host: ordinary download
        return t.cpu().numpy()
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def _plan_of(obj):
    if isinstance(obj, FemPlan):
        return obj
    plan = getattr(obj, "_fem_plan", None)
    if plan is None:
        raise TypeError("expected a FemPlan or a matrix returned by get_elastic_stiffness_matrix")
    return plan


def _host_B(plan, elements):
    """scipy CSR B (3 n_int x 2 n_n) with the reference's explicit zeros (:549-571); values from the GPU geometry."""
    n_p, n_q, n_int = plan.n_p, plan.n_q, plan.n_int
    d1 = plan.dphi1.cpu().numpy()
    d2 = plan.dphi2.cpu().numpy()
    node = np.repeat(np.asarray(elements, dtype=np.int64), n_q, axis=1)
    g = np.arange(n_int)
    z = np.zeros_like(d1)
    vals = np.stack([d1, z, d2, z, d2, d1], axis=1)
    rows = np.broadcast_to(3 * g + np.array([0, 1, 2, 0, 1, 2])[None, :, None], vals.shape)
    cols = 2 * node[:, None, :] + np.array([0, 0, 0, 1, 1, 1])[None, :, None]
    return ssp.csr_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(3 * n_int, plan.n_dof))


def _d_indices(n_int):
    aux = np.arange(3 * n_int).reshape((3, n_int), order='F') + 1
    return np.tile(aux, (3, 1)), np.repeat(aux, 3, axis=0)


def _host_K(plan, k_vals, prune=True):
    K = plan.to_scipy_csr(k_vals)
    if prune:
        K.eliminate_zeros()          # scipy's csr_matmat / binop drop exact zeros
    K = K.tocsc()
    K._fem_plan = plan
    K._fem_vals = k_vals
    return K


def get_elastic_stiffness_matrix(elements, coordinates, shear, bulk, dhatp1, dhatp2, wf, variant="plasticity",
                                 host_matrices=True, device=None):
    """K_elast = B^T D B.  ``variant='plasticity'`` (Plasticity2D_DP / tsx-tunnel): 0-based ``elements``, returns
    ``(K, B, weight, id, jd, D)``.  ``variant='elasticity2d'``: 1-based (possibly float) ``elements`` that are shifted
    IN PLACE like the reference (Elasticity2D/pythonFEM.py:389), returns ``(K, weight)``.
    ``K`` is a csc_matrix pruned of exact zeros; ``K._fem_plan`` / ``K._fem_vals`` keep the device plan and values."""
    if variant == "elasticity2d":
        elements -= 1
    plan = FemPlan(np.asarray(elements).astype(np.int64), coordinates, dhatp1, dhatp2, wf, device=device)
    k_vals = plan.assemble_elastic(shear, bulk)
    weight = plan.weight.cpu().numpy().reshape(1, -1).copy()
    if not host_matrices:                               # large meshes: values stay on the device (DeviceMatrix), no host B / D
        K = DeviceMatrix(plan, k_vals)
        return (K, weight) if variant == "elasticity2d" else (K, plan, weight, None, None, None)
    K = _host_K(plan, k_vals)
    if variant == "elasticity2d":
        return K, weight
    i_d, j_d = _d_indices(plan.n_int)
    B = _host_B(plan, elements)
    B._fem_plan = plan
    vd = plan.elastic_dmat(shear, bulk).cpu().numpy()
    D = ssp.csr_matrix((vd.ravel(), (i_d.ravel() - 1, j_d.ravel() - 1)))
    D._fem_plan = plan
    D._fem_shear, D._fem_bulk = np.asarray(shear, dtype=np.float64), np.asarray(bulk, dtype=np.float64)
    return K, B, weight, i_d, j_d, D


def construct_constitutive_problem(e, *args, apply_plastic_strain=False):
    """Both reference signatures:
        construct_constitutive_problem(e, ep_prev, shear, bulk, eta, c, apply_plastic_strain=False)       (Plasticity2D_DP)
        construct_constitutive_problem(e, e0, ep_prev, shear, bulk, eta, c, apply_plastic_strain=False)   (tsx-tunnel)
    Returns the reference's dict (s, ds, ind_p, lambda_final, ep) as NumPy arrays.  ``ep_prev`` is updated in place when
    ``apply_plastic_strain`` (:750-755).  ``lambda_final`` is None whenever an apex point exists, as in the reference
    (SURVEY B-3); the intended apex multipliers are returned under the extra key ``lambda_apex_intended``."""
    args = list(args)
    if len(args) in (6, 7) and isinstance(args[-1], (bool, np.bool_)):
        apply_plastic_strain = bool(args.pop())
    if len(args) == 5:
        e0 = None
        ep_prev, shear, bulk, eta, c = args
    elif len(args) == 6:
        e0, ep_prev, shear, bulk, eta, c = args
    else:
        raise TypeError("construct_constitutive_problem: wrong number of arguments")
    ep_dev = None
    if ep_prev is not None:
        ep_dev = torch.as_tensor(np.ascontiguousarray(ep_prev, dtype=np.float64)).cuda()
    r = dp_return_map(e, ep_dev, shear, bulk, eta, c, apply_plastic_strain=apply_plastic_strain, e0=e0, want_lambda=True)
    counts = r["counts"].cpu().numpy()
    ind_p = _to_numpy(r["ind_p"]).view(np.bool_)             # flags are 0/1 bytes
    lam = _to_numpy(r["lambda"]).reshape(1, -1)
    ep = _to_numpy(r["ep"])
    if e0 is not None and counts.sum() == 0:
        ep = np.zeros_like(ep)              # tsx early-out (tsx-tunnel/pythonFEM.py:1101-1103): ep_prev is ignored
    elif apply_plastic_strain and ep_prev is not None:
        ep_prev[...] = ep                   # the reference returns ep_prev itself, mutated
        ep = ep_prev
    out = {'s': _to_numpy(r["s"]), 'ds': _to_numpy(r["ds"]), 'ind_p': ind_p,
           'lambda_final': None if counts[1] > 0 else lam, 'ep': ep,
           'lambda_apex_intended': lam, 'n_smooth': int(counts[0]), 'n_apex': int(counts[1])}
    return out


# ---- the reference's inline Newton statements as functions ------------------------------------------------
def strain(plan_or_B, U):
    """E = reshape(B @ U(:), (3, n_int), 'F') (:1043); U is (2, n_n)."""
    plan = _plan_of(plan_or_B)
    u = np.ascontiguousarray(np.asarray(U, dtype=np.float64).reshape(-1, order='F'))
    return _to_numpy(plan.strain(u))


def assemble_tangent(plan_or_K, ds, mode="reference", D_elast=None, K_elast=None, host_matrix=True):
    """K_tangent (:1047-1050).  ``mode='reference'`` evaluates K_elast + B^T (D_p - D_elast) B in the reference's own
    order (bit-identical; needs ``K_elast`` from get_elastic_stiffness_matrix and its ``D``); ``mode='direct'`` sums
    B^T (w ds) B in one pass (fastest, equal within rounding)."""
    plan = _plan_of(plan_or_K)
    wrap = _host_K if host_matrix else DeviceMatrix     # host_matrix=False: values stay on the device (DeviceMatrix)
    if mode == "direct":
        return wrap(plan, plan.assemble_tangent(ds))
    K_elast = plan_or_K if K_elast is None else K_elast
    if D_elast is None or not hasattr(D_elast, "_fem_shear") or not hasattr(K_elast, "_fem_vals"):
        raise TypeError("mode='reference' needs K_elast and D returned by get_elastic_stiffness_matrix")
    vals = plan.assemble_tangent_ref(ds, D_elast._fem_shear, D_elast._fem_bulk, K_elast._fem_vals)
    return wrap(plan, vals)


def internal_force(plan_or_B, s):
    """F = B^T vec(w * s[0:3]) as an (n_dof, 1) column (:1058)."""
    plan = _plan_of(plan_or_B)
    return _to_numpy(plan.internal_force(np.asarray(s, dtype=np.float64)[0:3])).reshape(-1, 1)


class DeviceMatrix:
    """What ``assemble_tangent(..., host_matrix=False)`` returns: the values stay in HBM next to the plan's pattern, and the
    SciPy ``csc_matrix`` the reference would hold (pruned of exact zeros) is only built when ``.tocsc()`` / ``.host`` is
    asked for.  ``solve_increment`` and ``stopping_criterion`` accept it wherever they accept a matrix."""

    def __init__(self, plan, k_vals):
        self._fem_plan, self._fem_vals, self._host = plan, k_vals, None
        self.shape = (plan.n_dof, plan.n_dof)

    @property
    def host(self):
        if self._host is None:
            self._host = _host_K(self._fem_plan, self._fem_vals)
        return self._host

    def tocsc(self):
        return self.host

    def tocsr(self):
        return self.host.tocsr()


def solve_increment(K_tangent, F, Q, rtol=1e-13, maxit=200000, precond="auto", K_elast=None):
    """dU with dU[Q] = K_tangent[Q,Q]^-1 (-F[Q]) (:1062-1066), by preconditioned CG on the device; returns (2, n_n).
    ``precond``: "multigrid" (geometric V-cycle, mg.py: uniform-lattice meshes), "jacobi", or "auto" = multigrid where the
    mesh allows it.  The multigrid hierarchy is built once per (mesh, Q) from ``K_elast`` (default: the first matrix seen)."""
    plan = _plan_of(K_tangent)
    mask = plan.mask_u8(Q)
    rhs = -np.asarray(F, dtype=np.float64).reshape(-1)
    if precond in ("auto", "multigrid"):
        from .mg import MultigridPCG, MultigridUnsupported
        cache = plan.__dict__.setdefault("_mg_cache", {})
        key = hash(mask.cpu().numpy().tobytes())
        if key not in cache:
            try:
                cache[key] = MultigridPCG(plan, mask).setup((K_elast if K_elast is not None else K_tangent)._fem_vals)
            except MultigridUnsupported:
                if precond == "multigrid":
                    raise
                cache[key] = None
        if cache[key] is not None:
            x, its, rel = cache[key].solve(K_tangent._fem_vals, plan._f64(rhs), rtol=rtol, maxit=min(maxit, 2000))
            return _to_numpy(x).reshape((2, -1), order='F')
    x, its, rel = plan.pcg(K_tangent._fem_vals, rhs, mask, rtol=rtol, maxit=maxit)
    dU = _to_numpy(x).reshape((2, -1), order='F')
    return dU


def stopping_criterion(K_elast, dU, U_it, U_new):
    """sqrt(dU'K dU) / (sqrt(U_it'K U_it) + sqrt(U_new'K U_new)) (:1072-1075); inputs (2, n_n)."""
    plan = _plan_of(K_elast)
    v = [plan._f64(np.asarray(a, dtype=np.float64).reshape(-1, order='F')) for a in (dU, U_it, U_new)]
    q = torch.sqrt(plan.energy_norms(K_elast._fem_vals, *v)).cpu().numpy()
    return q[0] / (q[1] + q[2])


def create_midpoints_P2(coord, elem, device=None):
    """tsx-tunnel/pythonFEM.py:1508-1626: P1 -> P2 enrichment, same dict (keys, shapes, dtypes, midpoint numbering) as the
    reference; computed by ``meshgen.create_midpoints_p2`` (one sort over the edges instead of the reference's
    O(n_e^2) search) on ``device`` (default: CUDA when available)."""
    from .meshgen import create_midpoints_p2
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    d = create_midpoints_p2(torch.as_tensor(np.ascontiguousarray(coord, dtype=np.float64)).to(dev),
                            torch.as_tensor(np.ascontiguousarray(elem).astype(np.int64)).to(dev))
    out = {k: v.cpu().numpy() for k, v in d.items()}
    out["elem_ext"] = out["elem_ext"].astype(int)
    return out


def create_midpoints(elem_type, coord, elem, device=None):
    """tsx-tunnel/pythonFEM.py:1629-1633 (P2 only; P4 is out of scope)."""
    if elem_type == LagrangeElementType.P2:
        return create_midpoints_P2(coord, elem, device=device)
    raise NotImplementedError("only the P2 enrichment is provided")


# ---- load vectors of the linear-elastic demo (Elasticity2D/pythonFEM.py) ------------------------------
def get_quadrature_surface(el_type):
    """Elasticity2D/pythonFEM.py:112-132 -> (Xi_s (n_q_s,), WF_s (n_q_s,))."""
    pt = 1 / np.sqrt(3)
    if el_type in (LagrangeElementType.P1, LagrangeElementType.Q1):
        return np.array([0]), np.array([2])
    return np.array([-pt, pt]), np.array([1, 1])


def get_local_basis_surface(el_type, xi_s):
    """Elasticity2D/pythonFEM.py:212-243 -> (HatP_s (n_p_s, n_q_s), DHatP1_s)."""
    xi = np.asarray(xi_s)
    if el_type in (LagrangeElementType.P1, LagrangeElementType.Q1):
        return 0.5 * np.array([1 - xi, 1 + xi]), np.array([[-0.5], [0.5]])
    return (np.array([xi * (xi - 1) / 2, xi * (xi + 1) / 2, (xi + 1) * (1 - xi)]), np.array([xi - 0.5, xi + 0.5, -2 * xi]))


def _load_device(device):
    return torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")


def get_vector_volume(elements, coordinates, f_V_int, hatp, weight, device=None):
    """Elasticity2D/pythonFEM.py:246-292: vector of volume forces, csc_matrix (2, n_n) like the reference's."""
    from . import loads
    dev = _load_device(device)
    t = lambda a, dt=np.float64: torch.as_tensor(np.ascontiguousarray(np.asarray(a), dtype=dt)).to(dev)  # noqa: E731
    f = loads.vector_volume(t(elements, np.int64), np.shape(coordinates)[1], t(f_V_int), t(hatp), t(weight).reshape(-1))
    return ssp.csc_matrix(f.cpu().numpy())


def get_vector_traction(elements_s, coordinates, f_t_int, hatp_s, dhatp1_s, wf_s, device=None):
    """Elasticity2D/pythonFEM.py:295-364: vector of traction forces on the loaded (horizontal) side, csc_matrix (2, n_n)."""
    from . import loads
    dev = _load_device(device)
    t = lambda a, dt=np.float64: torch.as_tensor(np.ascontiguousarray(np.asarray(a), dtype=dt)).to(dev)  # noqa: E731
    f = loads.vector_traction(t(elements_s, np.int64), t(coordinates), t(f_t_int), t(hatp_s), t(dhatp1_s), t(wf_s))
    return ssp.csc_matrix(f.cpu().numpy())
