"""NumPy/SciPy restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The restatement keeps the reference's floating-point
operation ORDER, so on the same machine it is bit-identical to the reference
(checked in ``tests/test_oracle_vs_reference.py``); it is written from the
algorithm description in SURVEY.md Appendix A, not copied from the sources.

Conventions: DOF = 2*node + component, integration point
g = e*n_q + q, Voigt order (11, 22, 12[, 33]) with engineering shear, ``ds`` is
the column-major 3x3 tangent, masks are ``(2, n_n)`` booleans flattened F-order.
"""
from __future__ import annotations

import enum
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.sparse as sp


class ElementType(enum.Enum):
    """Plasticity2D_DP/pythonFEM.py:55-60 (tsx adds P4=5, out of scope)."""
    P1 = 1
    P2 = 2
    Q1 = 3
    Q2 = 4


# ----------------------------------------------------------------------------
# L1: reference-element tables (Plasticity2D_DP/pythonFEM.py:364-488)
# ----------------------------------------------------------------------------
def quadrature_volume(et: ElementType) -> Tuple[np.ndarray, np.ndarray]:
    """(Xi (2,n_q), WF (1,n_q)); follows Plasticity2D_DP/pythonFEM.py:398-410."""
    g = 1 / np.sqrt(3)
    if et == ElementType.P1:
        return np.array([[1 / 3], [1 / 3]]), np.array([[0.5]])
    if et == ElementType.P2:
        a, b, c_, d = 0.1012865073235, 0.7974269853531, 0.4701420641051, 0.0597158717898
        xi = np.array([[a, b, a, c_, c_, d, 1 / 3],
                       [a, a, b, d, c_, c_, 1 / 3]])
        w1, w2 = 0.1259391805448, 0.1323941527885
        return xi, 0.5 * np.array([[w1, w1, w1, w2, w2, w2, 0.225]])
    if et == ElementType.Q1:
        return np.array([[-g, -g, g, g], [-g, g, -g, g]]), np.array([[1, 1, 1, 1]])
    if et == ElementType.Q2:
        xi = np.array([[-g, g, g, -g, 0, g, 0, -g, 0],
                       [-g, -g, g, g, -g, 0, g, 0, 0]])
        wa, wb, wc = 25 / 81, 40 / 81, 64 / 81
        return xi, np.array([[wa, wa, wa, wa, wb, wb, wb, wb, wc]])
    raise ValueError(et)


def local_basis_volume(et: ElementType, xi: np.ndarray):
    """(HatP, DHatP1, DHatP2), each (n_p, n_q); Plasticity2D_DP/pythonFEM.py:434-488."""
    s, t = xi[0], xi[1]
    n_q = s.shape[0]
    z = np.zeros(n_q)
    if et == ElementType.P1:
        return (np.array([1 - s - t, s, t]),
                np.array([[-1], [1], [0]]), np.array([[-1], [0], [1]]))
    if et == ElementType.P2:
        r = 1 - s - t
        hat = np.array([r * (2 * r - 1), s * (2 * s - 1), t * (2 * t - 1), 4 * s * t, 4 * r * t, 4 * r * s])
        d1 = np.array([-4 * r + 1, 4 * s - 1, z, 4 * t, -4 * t, 4 * (r - s)])
        d2 = np.array([-4 * r + 1, z, 4 * t - 1, 4 * s, 4 * (r - t), -4 * s])
        return hat, d1, d2
    if et == ElementType.Q1:
        hat = np.array([(1 - s) * (1 - t) / 4, (1 + s) * (1 - t) / 4, (1 + s) * (1 + t) / 4, (1 - s) * (1 + t) / 4])
        d1 = np.array([-(1 - t) / 4, (1 - t) / 4, (1 + t) / 4, -(1 + t) / 4])
        d2 = np.array([-(1 - s) / 4, -(1 + s) / 4, (1 + s) / 4, (1 - s) / 4])
        return hat, d1, d2
    if et == ElementType.Q2:
        s2, t2 = pow(s, 2), pow(t, 2)
        hat = np.array([(1 - s) * (1 - t) * (-1 - s - t) / 4, (1 + s) * (1 - t) * (-1 + s - t) / 4,
                        (1 + s) * (1 + t) * (-1 + s + t) / 4, (1 - s) * (1 + t) * (-1 - s + t) / 4,
                        (1 - s2) * (1 - t) / 2, (1 + s) * (1 - t2) / 2, (1 - s2) * (1 + t) / 2,
                        (1 - s) * (1 - t2) / 2])
        d1 = np.array([(1 - t) * (2 * s + t) / 4, (1 - t) * (2 * s - t) / 4, (1 + t) * (2 * s + t) / 4,
                       (1 + t) * (2 * s - t) / 4, -s * (1 - t), (1 - t2) / 2, -s * (1 + t), -(1 - t2) / 2])
        d2 = np.array([(1 - s) * (s + 2 * t) / 4, (1 + s) * (-s + 2 * t) / 4, (1 + s) * (s + 2 * t) / 4,
                       (1 - s) * (-s + 2 * t) / 4, -(1 - s2) / 2, -(1 + s) * t, (1 - s2) / 2, -(1 - s) * t])
        return hat, d1, d2
    raise ValueError(et)


# ----------------------------------------------------------------------------
# L0: meshes
# ----------------------------------------------------------------------------
def square_mesh_p1(n_x: int, n_y: int, size_x: float, size_y: float) -> Dict[str, np.ndarray]:
    """Uniform P1 mesh, numbering of get_nodes_1 (Plasticity2D_DP/pythonFEM.py:73-122):
    node id = ix + iy*(n_x+1); cell (ix,iy) -> triangles (V1,V2,V4),(V2,V3,V4),
    cell-major with ix fastest.  Footing boundary data (:178-184) included."""
    cx = np.linspace(0, size_x, n_x + 1)
    cy = np.linspace(0, size_y, n_y + 1)
    coord = np.array([np.tile(cx, n_y + 1), np.repeat(cy, n_x + 1)])
    ix, iy = np.meshgrid(np.arange(n_x), np.arange(n_y), indexing='xy')
    v1 = (ix + iy * (n_x + 1)).ravel()
    v2, v4 = v1 + 1, v1 + (n_x + 1)
    v3 = v4 + 1
    elem = np.empty((3, 2 * n_x * n_y), dtype=np.int64)
    elem[:, 0::2] = (v1, v2, v4)
    elem[:, 1::2] = (v2, v3, v4)
    top_left = np.logical_and(coord[1] == size_y, coord[0] <= 1.0001)
    dirichlet = np.zeros(coord.shape)
    dirichlet[1, top_left] = 1
    q = coord > 0
    q[1, top_left] = 0
    q[0, coord[0] == size_x] = 0
    return {'coordinates': coord, 'elements': elem, 'dirichlet_nodes': dirichlet, 'Q': q}


def footing_mesh(level: int, et: ElementType, size_xy: int = 10) -> Dict[str, np.ndarray]:
    """Strip-footing mesh of Plasticity2D_DP (get_nodes_1, :63-186); P1 and Q1 only."""
    n = size_xy * 2 ** level
    m = square_mesh_p1(n, n, size_xy, size_xy)
    if et == ElementType.P1:
        return m
    if et == ElementType.Q1:
        ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing='xy')
        v1 = (ix + iy * (n + 1)).ravel()
        m['elements'] = np.array([v1, v1 + 1, v1 + n + 2, v1 + n + 1])
        return m
    raise NotImplementedError("oracle mesh generator covers P1/Q1 (get_nodes_1); P2/Q2 come from the reference")


# ----------------------------------------------------------------------------
# L2: elastic stiffness (Plasticity2D_DP/pythonFEM.py:491-601 == tsx-tunnel:432-542)
# ----------------------------------------------------------------------------
def _seq_sum(rows: np.ndarray):
    """Python ``sum`` over the first axis: 0 + r0 + r1 + ... (:530-533)."""
    acc = 0
    for r in rows:
        acc = acc + r
    return acc


def geometry(elements, coordinates, dhatp1, dhatp2, wf):
    """dphi1, dphi2 (n_p, n_int), det (n_int,), weight (1, n_int); :500-546, :585."""
    n_e = elements.shape[1]
    n_q = np.size(wf)
    idx = np.asarray(elements).astype(np.int64)
    xe = np.repeat(coordinates[0][idx], n_q, axis=1)        # :517-527
    ye = np.repeat(coordinates[1][idx], n_q, axis=1)
    h1 = np.tile(dhatp1, (1, n_e))                           # :510-511
    h2 = np.tile(dhatp2, (1, n_e))
    j11, j12 = _seq_sum(xe * h1), _seq_sum(ye * h1)          # :530-533
    j21, j22 = _seq_sum(xe * h2), _seq_sum(ye * h2)
    det = j11 * j22 - j12 * j21                              # :536
    i11, i12, i21, i22 = j22 / det, -j12 / det, -j21 / det, j11 / det   # :539-542
    dphi1 = i11 * h1 + i12 * h2                              # :545-546
    dphi2 = i21 * h1 + i22 * h2
    weight = np.abs(det) * np.tile(wf, (1, n_e))             # :585
    return dphi1, dphi2, det, weight


def elastic_dmat_coeffs():
    """The 9 (column-major) coefficients of 2*Dev and Vol used at :579-582."""
    iota = np.array([[1], [1], [0]])
    vol = iota * iota.T
    dev = np.diag([1, 1, 0.5]) - vol / 3
    return 2 * dev.reshape(-1, 1, order='F'), vol.reshape(-1, 1, order='F').astype(float)


def elastic_stiffness(elements, coordinates, shear, bulk, dhatp1, dhatp2, wf):
    """-> (K csc, B csr, weight (1,n_int), iD, jD (9,n_int) 1-based, D csr); :491-601."""
    n_n = coordinates.shape[1]
    n_p, n_e = elements.shape
    n_q = np.size(wf)
    n_int = n_e * n_q
    dphi1, dphi2, _, weight = geometry(elements, coordinates, dhatp1, dhatp2, wf)

    # B: rows 3g+(0,1,2); per node the 6 stored values [d1,0,d2 | 0,d2,d1] (:549-571)
    idx = np.asarray(elements).astype(np.int64)
    node = np.repeat(idx, n_q, axis=1)                       # (n_p, n_int)
    g = np.arange(n_int)
    zero = np.zeros_like(dphi1)
    vals = np.stack([dphi1, zero, dphi2, zero, dphi2, dphi1], axis=1)        # (n_p, 6, n_int)
    rows = np.broadcast_to(3 * g + np.array([0, 1, 2, 0, 1, 2])[None, :, None], vals.shape)
    cols = 2 * node[:, None, :] + np.array([0, 0, 0, 1, 1, 1])[None, :, None]
    B = sp.csr_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(3 * n_int, 2 * n_n))

    # D: block-diagonal, entry k of point g at (3g + k%3, 3g + k//3) (:579-592)
    dev2, vol = elastic_dmat_coeffs()
    elast = dev2 * shear + vol * bulk                        # (9, n_int)
    aux = np.arange(3 * n_int).reshape((3, n_int), order='F') + 1
    i_d = np.tile(aux, (3, 1))
    j_d = np.repeat(aux, 3, axis=0)
    vd = elast * (np.ones((9, 1)) * weight)
    D = sp.csr_matrix((vd.ravel(), (i_d.ravel() - 1, j_d.ravel() - 1)))
    K = B.transpose() @ D @ B                                # :595
    return K, B, weight, i_d, j_d, D


# ----------------------------------------------------------------------------
# L3: Drucker-Prager return map (Plasticity2D_DP:604-757, tsx-tunnel:990-1157)
# ----------------------------------------------------------------------------
def constitutive_problem(e, ep_prev, shear, bulk, eta, c, apply_plastic_strain=False,
                         e0: Optional[np.ndarray] = None, tsx_variant: bool = False):
    """dict(s (4,n), ds (9,n), ind_p (n,), lambda_final, ep (4,n)).

    ``e0``/``tsx_variant`` select the tsx-tunnel signature (adds ``e0`` to the
    strain at :1052 and skips the plastic branch when no point yields, :1103).
    ``ep_prev`` is updated in place when ``apply_plastic_strain`` (:750-755).
    The reference's ``lambda_a`` outer-product bug (:714, SURVEY B-3) is not
    reproduced: ``lambda_final`` holds the smooth-branch multipliers and NaN at
    apex points, where the reference would return ``None`` for the whole array.
    """
    n_int = len(shear)
    iota = np.array([1, 1, 0, 1])
    vol = np.outer(iota, iota)
    dev = np.diag([1, 1, 1 / 2, 1]) - vol / 3
    dev3, vol3 = dev[0:3, 0:3], vol[0:3, 0:3]

    e4 = np.concatenate([e, np.zeros((1, n_int))], axis=0)
    if e0 is not None:
        e4 = e4 + e0                                          # tsx :1052
    e_tr = e4
    if ep_prev is not None:
        e_tr -= ep_prev                                       # :666-668 (e_tr IS e4)
    dev_e = dev @ e_tr                                        # :673
    s_tr = 2 * np.tile(shear, (4, 1)) * (dev @ e_tr) + np.tile(bulk, (4, 1)) * (vol @ e_tr)   # :670
    sq = _seq_sum(e_tr * dev_e)                               # :676
    norm_e = np.sqrt(np.where(sq > 0, sq, 0.0))
    rho_tr = 2 * (shear * norm_e)                             # :679
    p_tr = bulk * (iota.T @ e_tr)                             # :682
    denom_a = bulk * (eta ** 2)                               # :687-690
    denom_s = shear + denom_a
    crit1 = rho_tr / np.sqrt(2) + eta * p_tr - c
    crit2 = eta * p_tr - denom_a * rho_tr / (shear * np.sqrt(2)) - c
    ind_p = crit1 > 0                                         # :693-699
    ind_s = np.logical_and(crit1 > 0, crit2 <= 0)
    ind_a = np.logical_and(crit1 > 0, crit2 > 0)

    s = s_tr
    ds = 2 * dev3.reshape(-1, 1) * shear + vol3.reshape(-1, 1) * bulk          # :703
    lam = np.zeros((1, n_int))
    ep = np.zeros((4, n_int))
    if tsx_variant and not ind_p.any():                       # tsx :1103
        return {'s': s, 'ds': ds, 'ind_p': ind_p, 'lambda_final': lam, 'ep': ep,
                'n_smooth': 0, 'n_apex': 0}

    lam_s = crit1[ind_s] / denom_s[ind_s]                     # :710
    n_hat = dev_e[:, ind_s] / np.tile(norm_e[ind_s], (4, 1))  # :718-721
    m_hat = np.tile(np.sqrt(2) * shear[ind_s], (4, 1)) * n_hat + np.outer(iota, bulk[ind_s] * eta[ind_s])
    s[:, ind_s] = s[:, ind_s] - np.tile(lam_s, (4, 1)) * m_hat
    s[:, ind_a] = np.outer(iota, c[ind_a] / eta[ind_a])
    n_smooth, n_apex = int(ind_s.sum()), int(ind_a.sum())
    ident = np.outer(dev3.flatten(), np.ones(n_smooth))       # :724-728
    nn = np.tile(n_hat[0:3], (3, 1)) * np.repeat(n_hat[0:3], 3, axis=0)
    mm = np.tile(m_hat[0:3], (3, 1)) * np.repeat(m_hat[0:3], 3, axis=0)
    coef = 2 * np.sqrt(2) * (shear[ind_s] ** 2) * lam_s / rho_tr[ind_s]
    ds[:, ind_s] = ds[:, ind_s] - np.tile(coef, (9, 1)) * (ident - nn) - mm / np.tile(denom_s[ind_s], (9, 1))
    ds[:, ind_a] = np.zeros((9, n_apex))
    lam[0, ind_s] = lam_s                                     # :740
    lam[0, ind_a] = np.nan                                    # reference: whole array -> None (:743-746)
    if apply_plastic_strain:                                  # :750-755
        ep = ep_prev
        ep[:, ind_s] += np.outer(np.array([1, 1, 2, 1]), lam_s) * (n_hat / np.sqrt(2) + np.outer(iota, eta[ind_s] / 3))
        if n_apex > 0:
            ep[:, ind_a] = e4[:, ind_a] - np.outer(iota, c[ind_a] / (3 * bulk[ind_a] * eta[ind_a]))
    return {'s': s, 'ds': ds, 'ind_p': ind_p, 'lambda_final': lam, 'ep': ep,
            'n_smooth': n_smooth, 'n_apex': n_apex}


# ----------------------------------------------------------------------------
# Newton-loop glue (Plasticity2D_DP/pythonFEM.py:1043-1075, tsx-tunnel:1771-1792)
# ----------------------------------------------------------------------------
def strain(B, U):
    """E = reshape(B @ U(:), (3, n_int), 'F'); :1043."""
    return (B @ U.reshape((-1, 1), order='F')).reshape((3, -1), order='F')


def tangent_stiffness(K_elast, B, D_elast, weight, ds, i_d, j_d):
    """K_tangent = K_elast + B^T (D_p - D_elast) B; :1047-1050."""
    n_int = ds.shape[1]
    vd = np.tile(np.reshape(weight, (1, -1)), (9, 1)) * ds
    d_p = sp.csr_matrix((vd.ravel(), (i_d.ravel() - 1, j_d.ravel() - 1)), shape=(3 * n_int, 3 * n_int))
    return K_elast + B.T * (d_p - D_elast) * B


def internal_force(B, weight, s):
    """F = B^T vec_F(w * s[0:3]) as a flat (2 n_n,) vector; :1058."""
    n_int = s.shape[1]
    ws = np.tile(np.reshape(weight, (1, -1)), (3, 1)) * s[0:3, :]
    return np.asarray(B.T * np.reshape(ws, (3 * n_int, 1), order='F')).ravel()


def masked_dense_solve(K, rhs_flat, q_mask, transposed=False):
    """x[Q] = K[Q,Q]^-1 rhs[Q] (node-major interleaved free-DOF order); :1062-1066.

    ``transposed``: Plasticity2D_DP extracts the masked block row-major and reshapes it
    F-order (:1064-1065), i.e. it factorises K[Q,Q]^T; tsx-tunnel (:1779) undoes that."""
    qf = np.asarray(q_mask).flatten(order='F')
    kqq = K.tocsr()[qf][:, qf].toarray()
    if transposed:
        kqq = np.ascontiguousarray(kqq.T)
    x = np.zeros(qf.shape[0])
    x[qf] = np.linalg.solve(kqq, rhs_flat[qf])
    return x


def newton_criterion(K_elast, d_u, u_it, u_new):
    """sqrt(dU'K dU) / (sqrt(U_it'K U_it) + sqrt(U_new'K U_new)); :1072-1075 (flat F-order inputs)."""
    q1 = np.sqrt(d_u @ (K_elast @ d_u))
    q2 = np.sqrt(u_it @ (K_elast @ u_it))
    q3 = np.sqrt(u_new @ (K_elast @ u_new))
    return q1 / (q2 + q3)


def transform(q_int, elements, weight):
    """Integration-point -> nodal weighted average; Plasticity2D_DP/pythonFEM.py:760-816."""
    n_p, n_e = elements.shape
    w = np.reshape(weight, (1, -1))
    n_q = w.shape[1] // n_e
    node = np.repeat(np.asarray(elements).astype(np.int64), n_q, axis=1)
    num = np.zeros(int(node.max()) + 1)
    den = np.zeros_like(num)
    np.add.at(num, node.ravel(), np.tile(w * q_int, (n_p, 1)).ravel())
    np.add.at(den, node.ravel(), np.tile(w, (n_p, 1)).ravel())
    return num / den


# ----------------------------------------------------------------------------
# Canonical forms used by the parity tests
# ----------------------------------------------------------------------------
def canonical_csr(K):
    Kc = sp.csr_matrix(K).copy()
    Kc.sum_duplicates()
    Kc.sort_indices()
    return Kc


def structural_pattern(B, D):
    """Pattern of B^T D B with every stored entry (explicit zeros included) set to one."""
    b1, d1 = B.copy(), D.copy()
    b1.data[:] = 1.0
    d1.data[:] = 1.0
    s = (b1.T @ d1 @ b1).tocsr()
    s.sort_indices()
    return s.indptr.astype(np.int64), s.indices.astype(np.int32)


# ----------------------------------------------------------------------------
# Drivers restated around the hot path (used for Newton-trace parity)
# ----------------------------------------------------------------------------
def footing_constants():
    """Plasticity2D_DP/pythonFEM.py:910-933."""
    young, poisson, c0, phi = 1e7, 0.48, 450, np.pi / 9
    shear = young / (2 * (1 + poisson))
    bulk = young / (3 * (1 - 2 * poisson))
    eta = 3 * np.tan(phi) / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    c = 3 * c0 / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    return shear, bulk, eta, c, c0


def tsx_constants():
    """tsx-tunnel/pythonFEM.py:1663-1681 with scalar indexing (the original raises on NumPy 2)."""
    young, poisson = 60000, 0.2
    shear = young / (2 * (1 + poisson))
    bulk = young / (3 * (1 - 2 * poisson))
    cohesion, phi = 18.7, 49 * np.pi / 180
    eta = 3 * np.tan(phi) / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    c = 3 * cohesion / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    s0 = np.array([-45.0, -11.0, 0.0, -60.0]).reshape((-1, 1))
    tr = s0[0, 0] + s0[1, 0] + s0[3, 0]
    e0 = np.array([-poisson * tr + (1 + poisson) * s0[0, 0],
                   -poisson * tr + (1 + poisson) * s0[1, 0],
                   0,
                   -poisson * tr + (1 + poisson) * s0[3, 0]], dtype=float).reshape((-1, 1)) / young
    return shear, bulk, eta, c, s0, e0


def tsx_q_mask(coords):
    """tsx-tunnel/pythonFEM.py:1695-1699."""
    q = np.ones(coords.shape, dtype=bool)
    q[0, coords[0] < -49.99] = 0
    q[0, coords[0] > 49.99] = 0
    q[1, coords[1] < -49.99] = 0
    q[1, coords[1] > 49.99] = 0
    return q


def tsx_driver(coords, elem, et=ElementType.P1, max_steps=100, record=None):
    """Restatement of tsx-tunnel/pythonFEM.py:1661-1830 for P1 (no midpoints, :1690-1692 skipped).

    Returns dict(U, steps, trace) where trace holds per Newton iteration
    (zeta, it, n_plast, criterion)."""
    shear0, bulk0, eta0, c0_, s0, e0 = tsx_constants()
    xi, wf = quadrature_volume(et)
    _, d1, d2 = local_basis_volume(et, xi)
    n_n, n_e = coords.shape[1], elem.shape[1]
    n_int = n_e * wf.size
    shear, bulk = shear0 * np.ones(n_int), bulk0 * np.ones(n_int)
    eta, c = eta0 * np.ones(n_int), c0_ * np.ones(n_int)
    q = tsx_q_mask(coords)
    K, B, weight, i_d, j_d, D = elastic_stiffness(elem, coords, shear, bulk, d1, d2, wf)
    weight = weight.flatten(order='F')
    f0 = internal_force(B, weight, np.tile(s0, (1, n_int)))                       # :1737
    u_elast = masked_dense_solve(K, -f0, q).reshape((2, n_n), order='F')           # :1748
    d_zeta = 1 / 17
    d_zeta_min, d_zeta_old = d_zeta / 10, d_zeta
    zeta_old, zeta_max = 0, 1
    u_it = d_zeta * u_elast
    U = np.zeros((2, n_n))
    u_old = -u_it
    ep_old = np.zeros((4, n_int))
    trace, step, f = [], 0, None
    k_tangent = None
    while step < max_steps:
        zeta = zeta_old + d_zeta
        e0z = zeta * e0
        criterion = np.inf
        for it in range(25):
            E = strain(B, u_it)
            cp = constitutive_problem(E, ep_old, shear, bulk, eta, c, e0=e0z, tsx_variant=True)
            k_tangent = tangent_stiffness(K, B, D, weight, cp['ds'], i_d, j_d)
            f = internal_force(B, weight, cp['s'])
            du = masked_dense_solve(k_tangent, -f, q)
            ui = u_it.flatten(order='F')
            un = ui + du
            criterion = newton_criterion(K, du, ui, un)
            trace.append((zeta, it, int(cp['ind_p'].sum()), float(criterion)))
            if np.isnan(criterion):
                break
            u_it = un.reshape((2, n_n), order='F')
            if criterion < 1e-12:
                break
        if criterion < 1e-10:
            u_old, U = U, u_it
            cp = constitutive_problem(strain(B, U), ep_old, shear, bulk, eta, c, e0=e0z, tsx_variant=True)
            ep_old = cp['ep']                                  # stays zero: SURVEY B-5
            zeta_old, d_zeta_old = zeta, d_zeta
            step += 1
        else:
            d_zeta = d_zeta / 2
        u_it = d_zeta * (U - u_old) / d_zeta_old + U
        if zeta_old >= zeta_max or d_zeta < d_zeta_min:
            break
    return {'U': U, 'steps': step, 'trace': trace, 'K_elast': K, 'K_tangent': k_tangent, 'F': f, 'F0': f0,
            'Q': q, 'B': B, 'weight': weight}


def footing_driver(level=1, et=ElementType.P1, max_steps=1000, mesh=None):
    """Restatement of Plasticity2D_DP/pythonFEM.py:986-1131 (load stepping + semismooth Newton)."""
    shear0, bulk0, eta0, c_, c0 = footing_constants()
    mesh = mesh or footing_mesh(level, et)
    coords, elem, q = mesh['coordinates'], mesh['elements'], mesh['Q']
    q_nd = mesh['dirichlet_nodes'][1, :] > 0
    xi, wf = quadrature_volume(et)
    _, d1, d2 = local_basis_volume(et, xi)
    n_n, n_e = coords.shape[1], elem.shape[1]
    n_int = n_e * np.size(wf)
    shear, bulk = shear0 * np.ones(n_int), bulk0 * np.ones(n_int)
    eta, c = eta0 * np.ones(n_int), c_ * np.ones(n_int)
    K, B, weight, i_d, j_d, D = elastic_stiffness(elem, coords, shear, bulk, d1, d2, wf)
    d_zeta = 1 / 1000
    d_zeta_min, d_zeta_old = d_zeta / 1300, d_zeta
    zeta_old, zeta_max = 0, 1
    ud = -d_zeta * mesh['dirichlet_nodes']                                           # :997-1004
    f = -(K @ ud.flatten(order='F'))
    u_it = ud.copy()
    sol = masked_dense_solve(K, f, q, transposed=True)
    qf = q.flatten(order='F')
    ui = u_it.flatten(order='F')
    ui[qf] = sol[qf]
    u_it = ui.reshape((2, n_n), order='F')
    U = np.zeros((2, n_n))
    u_old = -u_it
    ep_old = np.zeros((4, n_int))
    pressure_old = 0
    trace, hist, step = [], [], 1
    while step <= max_steps:
        zeta = zeta_old + d_zeta
        criterion = np.inf
        for it in range(25):
            E = strain(B, u_it)
            cp = constitutive_problem(E, ep_old, shear, bulk, eta, c)
            k_tangent = tangent_stiffness(K, B, D, weight, cp['ds'], i_d, j_d)
            fint = internal_force(B, weight, cp['s'])
            du = masked_dense_solve(k_tangent, -fint, q, transposed=True)
            ui = u_it.flatten(order='F')
            un = ui + du
            criterion = newton_criterion(K, du, ui, un)
            trace.append((zeta, it, int(cp['ind_p'].sum()), float(criterion)))
            if np.isnan(criterion):
                break
            u_it = un.reshape((2, n_n), order='F')
            if criterion < 1e-12:
                break
        if criterion < 1e-10:
            u_old, U = U, u_it
            cp = constitutive_problem(strain(B, U), ep_old, shear, bulk, eta, c, apply_plastic_strain=True)
            ep_old = cp['ep']
            zeta_old, d_zeta_old = zeta, d_zeta
            step += 1
            pa = transform(cp['s'][1, :], elem, weight)                               # :1105-1112
            pressure = -np.mean(pa[q_nd]) / c0
            hist.append((zeta, pressure))
            if pressure - pressure_old < 0.1 and criterion < 1e-12:
                d_zeta *= 2
            pressure_old = pressure
        else:
            d_zeta /= 2
        u_it = d_zeta * (U - u_old) / d_zeta_old + U
        if zeta_old >= zeta_max or d_zeta < d_zeta_min:
            break
    return {'U': U, 'steps': step, 'trace': trace, 'hist': hist, 'Ep': ep_old}


# ---- load vectors of the linear-elastic demo ------------------------------------------------------
def quadrature_surface(et: ElementType):
    """(Xi_s (n_q_s,), WF_s (n_q_s,)); Elasticity2D/pythonFEM.py:112-132."""
    pt = 1 / np.sqrt(3)
    if et in (ElementType.P1, ElementType.Q1):
        return np.array([0]), np.array([2])
    return np.array([-pt, pt]), np.array([1, 1])


def local_basis_surface(et: ElementType, xi_s):
    """(HatP_s (n_p_s,n_q_s), DHatP1_s); Elasticity2D/pythonFEM.py:212-243 (the linear case returns a (2,1) derivative)."""
    xi = xi_s
    if et in (ElementType.P1, ElementType.Q1):
        return 0.5 * np.array([1 - xi, 1 + xi]), np.array([[-0.5], [0.5]])
    return (np.array([np.multiply(xi, (xi - 1) / 2), np.multiply(xi, (xi + 1) / 2), np.multiply(xi + 1, 1 - xi)]),
            np.array([xi - 0.5, xi + 0.5, -2 * xi]))


def _scatter_in_input_order(vals, nodes, n_n):
    """What csc_matrix((v, (0, j)), shape=(1, n_n)) does with duplicates: summed per column in input order."""
    out = np.zeros(n_n)
    for v, j in zip(vals, nodes.astype(np.int64)):
        out[j] = out[j] + v
    return out


def vector_volume(elements, coordinates, f_v_int, hatp, weight):
    """f_V (2, n_n) dense; Elasticity2D/pythonFEM.py:246-292.  Entry (a, g) contributes hatp[a, q] * (weight[g] * f[c, g])
    to node elements[a, e]; duplicates are summed in the flatten('F') order, i.e. ascending g, then a."""
    n_n = coordinates.shape[1]
    n_p, n_e = elements.shape
    hatphi = np.tile(hatp, (1, n_e))
    nodes = np.kron(elements, np.ones((1, hatp.shape[1]))).flatten(order='F')
    out = []
    for c in range(2):
        v = np.multiply(hatphi, np.ones((n_p, 1)) * np.multiply(weight, f_v_int[c,])).flatten(order='F')
        out.append(_scatter_in_input_order(v, nodes, n_n))
    return np.array(out)


def vector_traction(elements_s, coordinates, f_t_int, hatp_s, dhatp1_s, wf_s):
    """f_t (2, n_n) dense; Elasticity2D/pythonFEM.py:295-364.  The Jacobian uses the x coordinate only (:345) and the load
    is the LAST integration point's value for every point (f_t_int[c, -1], :356-357) - both kept."""
    n_n = coordinates.shape[1]
    n_p_s, n_e_s = elements_s.shape
    n_q_s = wf_s.shape[0]
    dhatphi1_s = np.tile(dhatp1_s, (1, n_e_s))
    hatphi_s = np.tile(hatp_s, (1, n_e_s))
    coords1 = np.reshape(coordinates[0, elements_s.flatten(order='F').astype(int)], (n_p_s, n_e_s), order='F')
    coord_int1 = np.kron(coords1, np.ones((1, n_q_s)))
    j11 = sum(np.multiply(coord_int1, dhatphi1_s))
    weight_s = np.multiply(abs(j11), np.tile(wf_s, (1, n_e_s)))
    nodes = np.kron(elements_s, np.ones((1, n_q_s))).flatten(order='F')
    out = []
    for c in range(2):
        v = np.multiply(hatphi_s, np.dot(np.ones((n_p_s, 1)), np.multiply(weight_s, f_t_int[c, -1]))).flatten(order='F')
        out.append(_scatter_in_input_order(v, nodes, n_n))
    return np.array(out)


def elasticity2d_solve(et: ElementType, elements, coordinates, neumann_nodes, dirichlet_nodes, q_mask, shear, bulk,
                       volume_force=(0.0, -1.0), traction_force=(0.0, 450.0)):
    """The linear-elastic demo after mesh generation (Elasticity2D/pythonFEM.py:1100-1171): K, f = f_t + f_V - K u_D with
    u_D = dirichlet_nodes / 2, dense solve on the free DOFs, stored energy 0.5 u'Ku - (f_t + f_V)'u.  ``elements`` 0-based."""
    xi, wf = quadrature_volume(et)
    hatp, d1, d2 = local_basis_volume(et, xi)
    xs, ws = quadrature_surface(et)
    hs, ds = local_basis_surface(et, xs)
    n_int = elements.shape[1] * np.size(wf)
    K, _, weight, _, _, _ = elastic_stiffness(elements, coordinates, shear * np.ones(n_int), bulk * np.ones(n_int), d1, d2, wf)
    n_int_s = neumann_nodes.shape[1] * len(ws)
    f_v = vector_volume(elements, coordinates, np.dot(np.array([volume_force]).T, np.ones((1, n_int))), hatp, weight)
    f_t = vector_traction(neumann_nodes, coordinates, np.dot(np.array([traction_force]).T, np.ones((1, n_int_s))), hs, ds, ws)
    load = (f_t + f_v).flatten(order='F')
    ud = (0.5 * dirichlet_nodes).flatten(order='F')
    f = load - K @ ud
    qf = np.asarray(q_mask, dtype=bool).flatten(order='F')
    u = ud.copy()
    u[qf] = np.linalg.solve(K.tocsr()[qf][:, qf].toarray(), f[qf])
    return {"u": u, "energy": 0.5 * u @ (K @ u) - load @ u, "f_V": f_v, "f_t": f_t, "K": K}
