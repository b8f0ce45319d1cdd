"""Generate tests/golden/*.npz by EXECUTING THE REAL REFERENCE (container only).

    python -m oracle.make_golden

Every array below is produced by the unmodified reference functions imported
from /root/reference (oracle.ref_loader) or read from its tsx-tunnel CSV
fixtures; nothing here comes from the oracle restatement or from the CUDA path.
The fixtures are small and committed so that the GPU box (which has no
/root/reference) can check the oracle and the kernels against the reference.
"""
import contextlib
import io
import logging
import os
import re
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def canon(m):
    c = sp.csr_matrix(m).copy()
    c.sum_duplicates()
    c.sort_indices()
    return {"indptr": c.indptr.astype(np.int64), "indices": c.indices.astype(np.int32), "data": c.data.copy()}


def pack(prefix, d):
    return {f"{prefix}_{k}": v for k, v in d.items()}


def structural(B, D):
    b1, d1 = B.copy(), D.copy()
    b1.data[:] = 1.0
    d1.data[:] = 1.0
    s = (b1.T @ d1 @ b1).tocsr()
    s.sort_indices()
    return {"indptr": s.indptr.astype(np.int64), "indices": s.indices.astype(np.int32)}


def footing_consts():
    young, poisson, c0, phi = 1e7, 0.48, 450, np.pi / 9
    return (young / (2 * (1 + poisson)), young / (3 * (1 - 2 * poisson)),
            3 * np.tan(phi) / np.sqrt(9 + 12 * (np.tan(phi)) ** 2), 3 * c0 / np.sqrt(9 + 12 * (np.tan(phi)) ** 2))


def main():
    os.makedirs(OUT, exist_ok=True)
    rp, re_, rt = ref_loader.load("plasticity"), ref_loader.load("elasticity"), ref_loader.load("tsx")
    tsx_dir = os.path.join(ref_loader.REF_ROOT, "tsx-tunnel")
    P1 = rp.LagrangeElementType.P1
    xi, wf = rp.get_quadrature_volume(P1)
    _, d1, d2 = rp.get_local_basis_volume(P1, xi)
    G0, K0, eta0, c0 = footing_consts()

    # -- 1. tsx-tunnel P1 assembly + the reference's CSV goldens --------------------------------
    coords = np.genfromtxt(os.path.join(tsx_dir, "coord.csv"), delimiter=",")
    elem = np.genfromtxt(os.path.join(tsx_dir, "elem.csv"), delimiter=",", dtype=int) - 1
    n_e = elem.shape[1]
    Gt, Kt = 60000 / (2 * 1.2), 60000 / (3 * (1 - 2 * 0.2))
    K, B, w, iD, jD, D = rt.get_elastic_stiffness_matrix(elem, coords, Gt * np.ones(n_e), Kt * np.ones(n_e), d1, d2, wf)
    np.savez_compressed(os.path.join(OUT, "assembly_tsx_p1.npz"), coordinates=coords, elements=elem,
                        shear=Gt, bulk=Kt, weight=w, B_data=canon(B)["data"], **pack("K", canon(K)),
                        **pack("S", structural(B, D)))
    kqq = np.genfromtxt(os.path.join(tsx_dir, "k_tangent_qq.csv"), delimiter=",")
    np.savez_compressed(os.path.join(OUT, "tsx_csv_golden.npz"), **pack("kqq", canon(sp.csr_matrix(kqq))),
                        kqq_shape=np.array(kqq.shape),
                        fq=np.genfromtxt(os.path.join(tsx_dir, "fq.csv"), delimiter=","),
                        f0q=np.genfromtxt(os.path.join(tsx_dir, "f0q.csv"), delimiter=","))

    # -- 2. footing meshes: L1 full K, nnz known answers for L2/L3 (log line :598) --------------
    nnz = {}
    for level in (1, 2, 3):
        mesh = rp.assemble_mesh(level, P1, 10)
        ne = mesh["elements"].shape[1]
        Kf, Bf, wfoot, _, _, Df = rp.get_elastic_stiffness_matrix(mesh["elements"], mesh["coordinates"],
                                                                 G0 * np.ones(ne), K0 * np.ones(ne), d1, d2, wf)
        nnz[level] = Kf.nnz
        if level == 1:
            np.savez_compressed(os.path.join(OUT, "assembly_footing_p1_l1.npz"), coordinates=mesh["coordinates"],
                                elements=mesh["elements"], Q=mesh["Q"], dirichlet_nodes=mesh["dirichlet_nodes"],
                                shear=G0, bulk=K0, weight=wfoot, **pack("K", canon(Kf)), **pack("S", structural(Bf, Df)))
            mesh1, K1, B1, w1, D1 = mesh, Kf, Bf, wfoot, Df
            iD1, jD1 = _, _
    np.savez_compressed(os.path.join(OUT, "nnz_known_answers.npz"), levels=np.array(list(nnz)), nnz=np.array(list(nnz.values())))

    # -- 3. config 1: Elasticity2D variant, 1-based float elements (comparison_assembly...py:74-80)
    from oracle import fem_oracle as fo  # mesh generator only (numbering checked against the reference in tests)
    m = fo.square_mesh_p1(16, 16, 1.0, 1.0)
    el1 = (m["elements"] + 1).astype(float)
    ne = el1.shape[1]
    Kc1, wc1 = re_.get_elastic_stiffness_matrix(el1.copy(), m["coordinates"], 0.5 * np.ones(ne), 1.5 * np.ones(ne), d1, d2, wf)
    np.savez_compressed(os.path.join(OUT, "assembly_elasticity2d_n16.npz"), coordinates=m["coordinates"],
                        elements_1based=el1, shear=0.5, bulk=1.5, weight=wc1, **pack("K", canon(Kc1)))

    # -- 4. return map (all three branches, both signatures) ------------------------------------
    rng = np.random.default_rng(0)
    n = 2048
    E = np.array([[-3e-4], [-3e-4], [0]]) + 2e-4 * rng.standard_normal((3, n))
    E[:, :128] *= 8
    Ep = 1e-5 * rng.standard_normal((4, n))
    G = G0 * (1 + 0.1 * rng.random(n))
    Kb = K0 * (1 + 0.1 * rng.random(n))
    eta, c = eta0 * np.ones(n), c0 * np.ones(n)
    out = {"E": E, "Ep": Ep, "shear": G, "bulk": Kb, "eta": eta, "c": c}
    for apply in (False, True):
        r = rp.construct_constitutive_problem(E.copy(), Ep.copy(), G, Kb, eta, c, apply)
        for k in ("s", "ds", "ind_p", "ep"):
            out[f"pl{int(apply)}_{k}"] = r[k]
    # tsx signature: e0 small enough that points still yield (tsx1) and large enough that none does (tsx0: the
    # early-out at tsx-tunnel/pythonFEM.py:1103 then returns ep = zeros even with apply_plastic_strain=True)
    for tag, scale in (("tsx1", 0.05), ("tsx0", 0.7)):
        e0 = scale * np.array([[-1.04e-3], [-3.6e-4], [0], [-1.34e-3]])
        r = rt.construct_constitutive_problem(E.copy(), e0, Ep.copy(), G, Kb, eta, c, True)
        out[f"{tag}_e0"] = e0
        for k in ("s", "ds", "ind_p", "ep"):
            out[f"{tag}_{k}"] = r[k]
    np.savez_compressed(os.path.join(OUT, "return_map.npz"), **out)

    # -- 5. Newton glue on footing L1 in a plastic state (Plasticity2D_DP:1043-1058) ------------
    ne = mesh1["elements"].shape[1]
    n_n = mesh1["coordinates"].shape[1]
    K1, B1, w1, iD1, jD1, D1 = rp.get_elastic_stiffness_matrix(mesh1["elements"], mesh1["coordinates"], G0 * np.ones(ne),
                                                               K0 * np.ones(ne), d1, d2, wf)
    U = 3e-3 * rng.standard_normal((2, n_n))
    Eg = (B1 @ U.reshape((-1, 1), order="F")).reshape((3, -1), order="F")
    cp = rp.construct_constitutive_problem(Eg.copy(), np.zeros((4, ne)), G0 * np.ones(ne), K0 * np.ones(ne),
                                           eta0 * np.ones(ne), c0 * np.ones(ne))
    vD = np.tile(w1, (9, 1)) * cp["ds"]
    D_p = sp.csr_matrix((rp.flatten_row(vD)[0], (rp.flatten_row(iD1)[0] - 1, rp.flatten_row(jD1)[0] - 1)), shape=(3 * ne, 3 * ne))
    Kt = K1 + B1.T * (D_p - D1) * B1
    F = B1.T * np.reshape(np.tile(w1, (3, 1)) * cp["s"][0:3, :], (3 * ne, 1), order="F")
    np.savez_compressed(os.path.join(OUT, "newton_glue_footing_l1.npz"), U=U, E=Eg, s=cp["s"], ds=cp["ds"], ind_p=cp["ind_p"],
                        F=np.asarray(F).ravel(), **pack("Kt", canon(Kt)))

    # -- 6. the reference driver's own Newton trace on footing P1 level 1 (from its INFO log) ---
    class Grab(logging.Handler):
        def __init__(self):
            super().__init__()
            self.msgs = []

        def emit(self, rec):
            self.msgs.append(rec.getMessage())

    h = Grab()
    logging.getLogger().addHandler(h)
    logging.getLogger().setLevel(logging.INFO)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        rp.elasticity_fem(P1, 1, False)
    logging.getLogger().setLevel(logging.ERROR)
    logging.getLogger().removeHandler(h)
    crit = np.array([float(mm.split(":")[1]) for mm in h.msgs if mm.startswith("stopping criterion")])
    cnt = np.array([[int(x) for x in re.findall(r"= (\d+)", mm)[:2]] for mm in h.msgs if mm.startswith("plastic integration")])
    zeta = np.array([float(mm.split("=")[1]) for mm in h.msgs if mm.startswith("load factor")])
    np.savez_compressed(os.path.join(OUT, "footing_l1_trace.npz"), criterion=crit, smooth_apex=cnt, load_factor=zeta)

    # -- 7. all four element types on the level-0 footing mesh (oracle coverage; P2/Q kernels later)
    allt = {}
    for name in ("P1", "P2", "Q1", "Q2"):
        et = rp.LagrangeElementType[name]
        mesh = rp.assemble_mesh(0, et, 10)
        x_, w_ = rp.get_quadrature_volume(et)
        _, a1, a2 = rp.get_local_basis_volume(et, x_)
        ni = mesh["elements"].shape[1] * np.size(w_)
        Ka, Ba, wa, _, _, Da = rp.get_elastic_stiffness_matrix(mesh["elements"], mesh["coordinates"], G0 * np.ones(ni),
                                                             K0 * np.ones(ni), a1, a2, w_)
        allt.update({f"{name}_coordinates": mesh["coordinates"], f"{name}_elements": mesh["elements"], f"{name}_weight": wa,
                     f"{name}_Q": mesh["Q"]})
        allt.update(pack(f"{name}_K", canon(Ka)))
        allt.update(pack(f"{name}_S", structural(Ba, Da)))
    np.savez_compressed(os.path.join(OUT, "assembly_all_types_l0.npz"), **allt)
    # -- 8. tsx-tunnel P2: mesh from the reference's own create_midpoints (:1508-1633), K and F0 = B^T (w sigma0) (:1737);
    #        F0[Q] is what the reference's f0q.csv holds (3594 free DOFs)
    d = rt.create_midpoints(rt.LagrangeElementType.P2, coords, elem)
    ce, ee = d["coord_ext"], d["elem_ext"].astype(np.int64)
    P2 = rt.LagrangeElementType.P2
    x2, w2 = rt.get_quadrature_volume(P2)
    _, b1, b2 = rt.get_local_basis_volume(P2, x2)
    ni = ee.shape[1] * np.size(w2)
    g_tsx, k_tsx = 60000 / (2 * 1.2), 60000 / (3 * (1 - 2 * 0.2))
    K2, B2, wt2, _, _, _ = rt.get_elastic_stiffness_matrix(ee, ce, g_tsx * np.ones(ni), k_tsx * np.ones(ni), b1, b2, w2)
    s0 = np.array([-45, -11, 0, -60]).reshape((-1, 1), order="F")
    F0 = B2.T @ np.reshape(np.tile(wt2.flatten(order="F"), (3, 1)) * s0[0:3, :], (3 * ni, 1), order="F")
    np.savez_compressed(os.path.join(OUT, "assembly_tsx_p2.npz"), coordinates=ce, elements=ee, shear=g_tsx, bulk=k_tsx, weight=wt2,
                        F0=np.asarray(F0).ravel(), **pack("K", canon(K2)))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
