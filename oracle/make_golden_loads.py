"""Golden vectors for the load vectors of the linear-elastic demo, produced by EXECUTING THE REAL REFERENCE
(Elasticity2D/pythonFEM.py:246-364 get_vector_volume / get_vector_traction, container only):

    python -m oracle.make_golden_loads

Follows the driver's own sequence (:1092-1139): assemble_mesh, get_elastic_stiffness_matrix (which turns mesh['elements']
0-based in place, :389), then the two load vectors with the driver's constant volume force (0,-1) and traction (0,450)
and, for a non-trivial case, with spatially varying loads.  P2 is absent: its Elasticity2D mesh generator raises (:698)."""
import contextlib
import io
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    re_ = ref_loader.load("elasticity")
    out = {}
    young, poisson = 206900, 0.29                                           # the driver's constants (:1072-1075)
    shear, bulk = young / (2 * (1 + poisson)), young / (3 * (1 - 2 * poisson))
    for name, et, level in (("P1", re_.LagrangeElementType.P1, 1), ("Q1", re_.LagrangeElementType.Q1, 1), ("Q2", re_.LagrangeElementType.Q2, 0)):
        mesh = re_.assemble_mesh(level, et, 10, 5)
        xi, wf = re_.get_quadrature_volume(et)
        xi_s, wf_s = re_.get_quadrature_surface(et)
        hatp, d1, d2 = re_.get_local_basis_volume(et, xi)
        hatp_s, d1_s = re_.get_local_basis_surface(et, xi_s)
        n_e = mesh["elements"].shape[1]
        n_int = n_e * wf.size
        K, weight = re_.get_elastic_stiffness_matrix(mesh["elements"], mesh["coordinates"], shear * np.ones(n_int), bulk * np.ones(n_int), d1, d2, wf)
        n_int_s = mesh["neumann_nodes"].shape[1] * len(wf_s)
        rng = np.random.default_rng(5)
        for tag, fv, ft in (("const", np.dot(np.array([[0, -1]]).T, np.ones((1, n_int))), np.dot(np.array([[0, 450]]).T, np.ones((1, n_int_s)))),
                            ("rand", rng.standard_normal((2, n_int)), rng.standard_normal((2, n_int_s)))):
            f_v = re_.get_vector_volume(mesh["elements"], mesh["coordinates"], fv, hatp, weight)
            f_t = re_.get_vector_traction(mesh["neumann_nodes"], mesh["coordinates"], ft, hatp_s, d1_s, wf_s)
            out[f"{name}_{tag}_fv_int"], out[f"{name}_{tag}_ft_int"] = fv, ft
            out[f"{name}_{tag}_f_V"], out[f"{name}_{tag}_f_t"] = f_v.toarray(), f_t.toarray()
        # the rest of the driver (:1141-1171): f = f_t + f_V - K ud, dense solve on the free DOFs, stored energy
        f_V = re_.get_vector_volume(mesh["elements"], mesh["coordinates"], np.dot(np.array([[0, -1]]).T, np.ones((1, n_int))), hatp, weight).reshape((-1, 1), order="F")
        f_t = re_.get_vector_traction(mesh["neumann_nodes"], mesh["coordinates"], np.dot(np.array([[0, 450]]).T, np.ones((1, n_int_s))), hatp_s, d1_s, wf_s).reshape((-1, 1), order="F")
        ud = 0.5 * mesh["dirichlet_nodes"]
        f = np.asarray(f_t + f_V - (K @ ud.reshape((-1, 1), order="F"))).reshape(-1)       # sparse + sparse - dense -> np.matrix, as in the driver
        qf = mesh["Q"].flatten(order="F").astype(bool)
        u = ud.flatten(order="F").copy()
        u[qf] = np.linalg.solve(K.tocsr()[qf][:, qf].toarray(), f[qf])
        load = np.asarray((f_t + f_V).todense()).reshape(-1)
        out[f"{name}_u"], out[f"{name}_energy"] = u, 0.5 * u @ (K @ u) - load @ u
        buf = io.StringIO()                                                 # the driver's own "Stored energy" print pins the glue above
        with contextlib.redirect_stdout(buf):
            re_.elasticity_fem(et, level, draw=False)
        printed = float(re.search(r"Stored energy: ([-0-9.e+]+)", buf.getvalue()).group(1))
        assert abs(printed - out[f"{name}_energy"]) <= 1e-12 * abs(printed), (printed, out[f"{name}_energy"])
        out[f"{name}_energy_printed"] = printed
        out[f"{name}_shear"], out[f"{name}_bulk"] = shear, bulk
        out[f"{name}_Q"], out[f"{name}_dirichlet_nodes"] = mesh["Q"].astype(bool), mesh["dirichlet_nodes"]
        out[f"{name}_elements"] = np.asarray(mesh["elements"]).astype(np.int64)
        out[f"{name}_coordinates"] = mesh["coordinates"]
        out[f"{name}_neumann_nodes"] = np.asarray(mesh["neumann_nodes"]).astype(np.int64)
        out[f"{name}_weight"] = np.asarray(weight)
        print(name, "energy", out[f"{name}_energy"], mesh["elements"].shape, mesh["neumann_nodes"].shape, np.abs(out[f"{name}_const_f_t"]).sum(), np.abs(out[f"{name}_const_f_V"]).sum())
    np.savez_compressed(os.path.join(OUT, "load_vectors.npz"), **out)


if __name__ == "__main__":
    main()
