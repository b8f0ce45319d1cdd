"""Newton-trace parity: the device-resident drivers against the reference driver's own log (footing L1)
and against the oracle's restated tsx driver.  The reference solves each Newton system with a dense LU
(Plasticity2D_DP/pythonFEM.py:1066); here it is Jacobi-PCG driven to rtol 1e-13."""
import numpy as np
import pytest

from oracle import fem_oracle as fo

pytestmark = pytest.mark.gpu


def test_footing_l1_trace_matches_reference_log(golden):
    from fem_elastoplasticity_b200 import newton
    g, m = golden("footing_l1_trace.npz"), golden("assembly_footing_p1_l1.npz")
    mesh = {k: m[k] for k in ("coordinates", "elements", "Q", "dirichlet_nodes")}
    out = newton.footing_driver(mesh)
    crit = np.array([t[3] for t in out["trace"]])
    ref = g["criterion"]
    assert crit.shape == ref.shape == (109,), "same number of Newton iterations as the reference driver"
    # PCG (rtol 1e-13) instead of a dense LU: criteria agree to 1e-5 relative down to the 1e-12 noise floor
    np.testing.assert_allclose(crit, ref, rtol=1e-5, atol=1e-12)
    assert out["steps"] - 1 == len(g["load_factor"]) == 16
    # plastic point counts: the reference logs smooth/apex per return-map call
    n_plast = [t[2] for t in out["trace"]]
    it = iter(g["smooth_apex"].sum(axis=1))
    assert all(any(c == r for r in it) for c in n_plast)
    # displacement parity against the oracle's dense-solve driver
    oref = fo.footing_driver(1)
    err = np.abs(out["U"] - oref["U"]).max() / np.abs(oref["U"]).max()
    print("footing L1 displacement error vs dense-solve oracle:", err)
    assert err < 1e-9
    np.testing.assert_allclose(out["Ep"], oref["Ep"], rtol=1e-6, atol=1e-9 * np.abs(oref["Ep"]).max())


def test_tsx_driver_matches_oracle(golden):
    from fem_elastoplasticity_b200 import newton
    m = golden("assembly_tsx_p1.npz")
    out = newton.tsx_driver(m["coordinates"], m["elements"])
    oref = fo.tsx_driver(m["coordinates"], m["elements"])
    assert out["steps"] == oref["steps"] == 17
    assert [t[1] for t in out["trace"]] == [t[1] for t in oref["trace"]]          # Newton iteration indices
    assert [t[2] for t in out["trace"]] == [t[2] for t in oref["trace"]]          # plastic point counts
    err = np.abs(out["U"] - oref["U"]).max() / np.abs(oref["U"]).max()
    print("tsx displacement error vs dense-solve oracle:", err)
    assert err < 1e-10
    assert np.array_equal(out["F0"], oref["F0"])                                   # B^T (w sigma0): bit-exact
    # golden fq.csv: the converged residual on the free DOFs is ~0
    g = golden("tsx_csv_golden.npz")
    ns = out["solver"]
    qf = out["Q"].flatten(order="F")
    res = ns.F.cpu().numpy()[qf]
    assert res.shape == g["fq"].shape
    assert np.abs(res).max() < 1e-8 * np.abs(out["F0"]).max()


def test_tangent_reference_mode_driver(golden):
    """Same driver with K_tangent evaluated in the reference's own order (bit-identical matrices)."""
    from fem_elastoplasticity_b200 import newton
    m = golden("assembly_tsx_p1.npz")
    a = newton.tsx_driver(m["coordinates"], m["elements"], tangent_mode="reference")
    b = newton.tsx_driver(m["coordinates"], m["elements"], tangent_mode="direct")
    assert a["steps"] == b["steps"] == 17
    assert np.abs(a["U"] - b["U"]).max() <= 1e-10 * np.abs(a["U"]).max()


def test_drivers_with_two_level_preconditioner(golden):
    """Whole load-stepping runs with the two-level PCG as the inner solver: same Newton traces and displacements."""
    from fem_elastoplasticity_b200 import newton
    m = golden("assembly_tsx_p1.npz")
    a = newton.tsx_driver(m["coordinates"], m["elements"], precond="twolevel", coarse_cells=8)
    oref = fo.tsx_driver(m["coordinates"], m["elements"])
    assert a["steps"] == oref["steps"] == 17
    assert [t[1:3] for t in a["trace"]] == [t[1:3] for t in oref["trace"]]
    assert np.abs(a["U"] - oref["U"]).max() <= 1e-10 * np.abs(oref["U"]).max()
    f = golden("assembly_footing_p1_l1.npz")
    mesh = {k: f[k] for k in ("coordinates", "elements", "Q", "dirichlet_nodes")}
    b = newton.footing_driver(mesh, precond="twolevel", coarse_cells=4)
    g = golden("footing_l1_trace.npz")
    crit = np.array([t[3] for t in b["trace"]])
    assert crit.shape == (109,)
    np.testing.assert_allclose(crit, g["criterion"], rtol=1e-5, atol=1e-12)
    j = newton.footing_driver(mesh)                                   # same run with point-Jacobi as the inner preconditioner
    tl_its, j_its = sum(t[4] for t in b["trace"]), sum(t[4] for t in j["trace"])
    print("inner iterations over the run: two-level", tl_its, "jacobi", j_its)
    assert tl_its < j_its
    assert np.abs(b["U"] - j["U"]).max() <= 1e-9 * np.abs(j["U"]).max()
