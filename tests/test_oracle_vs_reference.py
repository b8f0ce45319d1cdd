"""Oracle vs the LIVE reference imported from /root/reference (build container only)."""
import numpy as np
import pytest

from oracle import fem_oracle as fo, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")


def same(a, b):
    a, b = fo.canonical_csr(a), fo.canonical_csr(b)
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data, b.data))


@pytest.mark.parametrize("name,level", [("P1", 2), ("Q1", 1), ("P2", 1), ("Q2", 0)])
def test_assembly_bit_exact(name, level):
    rp = ref_loader.load("plasticity")
    et_r, et_o = rp.LagrangeElementType[name], fo.ElementType[name]
    mesh = rp.assemble_mesh(level, et_r, 10)
    xi, wf = rp.get_quadrature_volume(et_r)
    _, d1, d2 = rp.get_local_basis_volume(et_r, xi)
    xo, wo = fo.quadrature_volume(et_o)
    _, e1, e2 = fo.local_basis_volume(et_o, xo)
    assert np.array_equal(xi, xo) and np.array_equal(wf, wo) and np.array_equal(d1, e1) and np.array_equal(d2, e2)
    n_int = mesh["elements"].shape[1] * np.size(wf)
    rng = np.random.default_rng(1)
    G, K = 3.4e6 * (1 + rng.random(n_int)), 8.3e7 * (1 + rng.random(n_int))
    R = rp.get_elastic_stiffness_matrix(mesh["elements"], mesh["coordinates"], G, K, d1, d2, wf)
    O = fo.elastic_stiffness(mesh["elements"], mesh["coordinates"], G, K, e1, e2, wo)
    assert same(R[0], O[0]) and same(R[1], O[1]) and same(R[5], O[5])
    assert np.array_equal(R[2], O[2]) and np.array_equal(R[3], O[3]) and np.array_equal(R[4], O[4])
    if name in ("P1", "Q1"):
        m2 = fo.footing_mesh(level, et_o)
        assert all(np.array_equal(mesh[k], m2[k]) for k in ("coordinates", "elements", "dirichlet_nodes", "Q"))


def test_return_map_bit_exact_both_variants():
    rp, rt = ref_loader.load("plasticity"), ref_loader.load("tsx")
    G0, K0, eta0, c0, _ = fo.footing_constants()
    rng = np.random.default_rng(2)
    n = 20000
    E = np.array([[-3e-4], [-3e-4], [0]]) + 2e-4 * rng.standard_normal((3, n))
    E[:, :400] *= 6
    Ep = 1e-5 * rng.standard_normal((4, n))
    G, K = G0 * (1 + 0.1 * rng.random(n)), K0 * (1 + 0.1 * rng.random(n))
    eta, c = eta0 * np.ones(n), c0 * np.ones(n)
    for apply in (False, True):
        r = rp.construct_constitutive_problem(E.copy(), Ep.copy(), G, K, eta, c, apply)
        o = fo.constitutive_problem(E.copy(), Ep.copy(), G, K, eta, c, apply)
        assert all(np.array_equal(r[k], o[k]) for k in ("s", "ds", "ind_p", "ep"))
        assert r["lambda_final"] is None                      # SURVEY B-3
    e0 = 0.5 * fo.tsx_constants()[5]
    r = rt.construct_constitutive_problem(E.copy(), e0, Ep.copy(), G, K, eta, c, True)
    o = fo.constitutive_problem(E.copy(), Ep.copy(), G, K, eta, c, True, e0=e0, tsx_variant=True)
    assert all(np.array_equal(r[k], o[k]) for k in ("s", "ds", "ind_p", "ep"))


def test_material_constants():
    rp = ref_loader.load("plasticity")
    # the constants are literals inside elasticity_fem (:910-933); restated values must reproduce eta, c
    phi = np.pi / 9
    assert fo.footing_constants()[2] == 3 * np.tan(phi) / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    assert rp.LagrangeElementType.P1.value == fo.ElementType.P1.value == 1


def test_transform_matches():
    rp = ref_loader.load("plasticity")
    m = fo.footing_mesh(1, fo.ElementType.P1)
    rng = np.random.default_rng(3)
    n_e = m["elements"].shape[1]
    w, q = rng.random((1, n_e)) + 0.1, rng.standard_normal(n_e)
    ref = np.asarray(rp.transform(q, m["elements"], w)).ravel()
    np.testing.assert_allclose(fo.transform(q, m["elements"], w), ref, rtol=1e-13)
