"""Host logic of the geometric multigrid (no GPU): level layouts of the strip partition, Chebyshev coefficients, and the
NumPy statement of the V-cycle (tests/mg_reference.py, the checker of the GPU tests) as a CG preconditioner."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem_oracle as fo
from mg_reference import ReferenceMG, prolongation
from fem_elastoplasticity_b200.mg import check_schedule, chebyshev_coefficients, level_layouts, next_check


def strip_owned(ny_cells, world):
    loc = ny_cells // world
    return [(0 if r == 0 else r * loc + 1, (r + 1) * loc + 1) for r in range(world)]


@pytest.mark.parametrize("nx,ny,world", [(2828, 2828, 1), (2828, 2832, 8), (2828, 22624, 8), (100, 64, 2), (37, 40, 4), (10, 12, 3), (5, 3, 1)])
def test_level_layouts_cover_what_the_kernels_read(nx, ny, world):
    LX, NY = nx + 1, ny + 1
    owned = strip_owned(ny, world)
    lays = level_layouts(LX, NY, owned)
    assert lays[-1]["replicated"] and 2 * lays[-1]["nxn"] * lays[-1]["nrows_global"] <= max(2500, 8)
    f_own, f_N, f_nxn, f_rep = owned, NY, LX, world == 1
    f_local = [(max(0, lo - 1), min(NY, hi + 1)) for lo, hi in owned]            # level 0: owned rows + one ghost row per side
    for lay in lays:
        N, nxn = lay["nrows_global"], lay["nxn"]
        assert nxn == -(-(f_nxn - 1) // 2) + 1 and N == -(-(f_N - 1) // 2) + 1
        share = lay["owned_global"]
        if f_rep:                                          # below a replicated level every rank holds (and computes) all rows
            assert all(sh == (0, N) for sh in share)
        else:
            assert share[0][0] == 0 and share[-1][1] == N and all(share[i][1] == share[i + 1][0] for i in range(world - 1))
        for rk, (g0, nrows, olo, ohi, rlo, rhi) in enumerate(lay["ranks"]):
            assert 0 <= olo < ohi <= nrows and 0 <= rlo <= rhi <= nrows and g0 + nrows <= N
            if lay["replicated"]:
                assert (g0, nrows, olo, ohi) == (0, N, 0, N)
            else:
                assert (g0 + olo, g0 + ohi) == share[rk] and ohi - olo >= 3
            # restriction (gather): coarse row J of this rank's share reads fine rows 2J-1 .. 2J+1 - all local on the fine level
            flo, fhi = (0, f_N) if f_rep else f_local[rk]
            for J in {g0 + rlo, g0 + rhi - 1} if rhi > rlo else ():
                for fj in (2 * J - 1, 2 * J, 2 * J + 1):
                    if 0 <= fj < f_N:
                        assert flo <= fj < fhi, (lay["nxn"], rk, J, fj)
            # prolongation (gather): every owned fine row reads its parents floor(j/2), ceil(j/2) - all local on this level
            own_f = (0, f_N) if f_rep else f_own[rk]
            for j in {own_f[0], own_f[1] - 1}:
                for J in (j // 2, (j + 1) // 2):
                    assert g0 <= J < g0 + nrows, (rk, j, J)
        # the shares of the restricted rows tile the level
        if lay["first_replicated"] and world > 1:
            assert [(g0 + rlo, g0 + rhi) for g0, _, _, _, rlo, rhi in lay["ranks"]] == share
        f_rep = lay["replicated"]
        f_own = [(0, N)] * world if f_rep else share
        f_local = [(g0, g0 + nrows) for g0, nrows, *_ in lay["ranks"]]
        f_N, f_nxn = N, nxn


def test_chebyshev_coefficients_damp_the_target_interval():
    lmax, ratio = 1.9, 8.0
    for k in (1, 2, 3, 5):
        c1, c2 = chebyshev_coefficients(lmax, ratio, k)
        lam = np.linspace(lmax / ratio, lmax, 400)
        x, d = np.zeros_like(lam), np.zeros_like(lam)      # scalar model problem lam*x = 1, D = 1
        for j in range(k):
            d = c1[j] * d + c2[j] * (1.0 - lam * x)
            x = x + d
        bound = 1.0 / np.cosh(k * np.arccosh((ratio + 1) / (ratio - 1)))
        assert np.abs(1.0 - lam * x).max() <= bound * (1 + 1e-9)


def test_numpy_vcycle_is_an_h_independent_spd_preconditioner():
    et = fo.ElementType.P1
    xi, wf = fo.quadrature_volume(et)
    _, d1, d2 = fo.local_basis_volume(et, xi)
    its = {}
    for N in (24, 48):
        m = fo.square_mesh_p1(N, N, 10.0, 10.0)
        n_e = m["elements"].shape[1]
        G0, K0, eta0, c0, _ = fo.footing_constants()
        K = fo.elastic_stiffness(m["elements"], m["coordinates"], G0 * np.ones(n_e), K0 * np.ones(n_e), d1, d2, wf)[0].tocsr()
        q = m["Q"].flatten(order="F")
        levels = 2 if N == 24 else 3
        mg = ReferenceMG(K, q, N + 1, N + 1, levels, 3, 8.0)
        mg.set_bounds([1.0] * levels)                       # builds the block inverses; bounds from their spectra next
        lmax = [1.1 * np.abs(np.linalg.eigvals((Di @ A).toarray())).max() for A, Di in zip(mg.A[:-1], mg.dinv)]
        mg.set_bounds(lmax)
        rng = np.random.default_rng(0)
        u, v = rng.standard_normal(K.shape[0]) * q, rng.standard_normal(K.shape[0]) * q
        assert abs(u @ mg.vcycle(v) - v @ mg.vcycle(u)) <= 1e-10 * abs(u @ mg.vcycle(v)) and u @ mg.vcycle(u) > 0
        b = rng.standard_normal(K.shape[0]) * q
        x, r = np.zeros_like(b), b.copy()
        z = mg.vcycle(r)
        p, rz = z.copy(), r @ z
        for it in range(1, 200):
            Ap = q * (K @ p)
            al = rz / (p @ Ap)
            x, r = x + al * p, r - al * Ap
            if np.linalg.norm(r) <= 1e-10 * np.linalg.norm(b):
                break
            z = mg.vcycle(r)
            rz, rz_old = r @ z, rz
            p = z + (rz / rz_old) * p
        its[N] = it
    assert its[48] <= its[24] + 6 and max(its.values()) <= 50, its


def test_prolongation_reproduces_linear_fields():
    P, nxc, nyc = prolongation(9, 6)
    xc, yc = np.meshgrid(np.arange(nxc) * 2.0, np.arange(nyc) * 2.0)
    fc = np.stack([(1 + 2 * xc + 3 * yc).ravel(), (xc - yc).ravel()], axis=1).ravel()
    xf, yf = np.meshgrid(np.arange(9) * 1.0, np.arange(6) * 1.0)
    ff = np.stack([(1 + 2 * xf + 3 * yf).ravel(), (xf - yf).ravel()], axis=1).ravel()
    np.testing.assert_allclose(P @ fc, ff, atol=1e-12)


def test_convergence_check_schedule():
    """Checks start four iterations before the previous solve's count; the stopping rule (first check with rel <= rtol) is kept."""
    def run(n_conv, hint, check_every=2, maxit=500, start=0):
        it, checks, skip = start, [], check_schedule(hint, check_every)
        while it < maxit:
            it = next_check(it, maxit, check_every, skip)
            checks.append(it)
            if it >= n_conv:
                break
        return it, checks
    assert run(28, None) == (28, list(range(2, 29, 2)))
    assert run(28, 28) == (28, [24, 26, 28]) and run(27, 28) == (28, [24, 26, 28])
    assert run(12, 28) == (24, [24])                      # much faster than its predecessor: a few iterations too many
    assert run(40, 28)[1][:2] == [24, 26] and run(40, 28)[0] == 40
    assert run(3, 3) == (4, [2, 4]) and run(28, 28, start=2)[1] == [24, 26, 28]   # start=2: the two eager iterations before the graph
    assert run(30, 28, check_every=3) == (30, [24, 27, 30]) and run(5, 2, maxit=4) == (4, [2, 4])
    for hint in (None, 0, 1, 5, 28, 499):
        for ce in (1, 2, 3, 7):
            it, last = 0, -1
            skip = check_schedule(hint, ce)
            while it < 50:
                it = next_check(it, 50, ce, skip)
                assert it > last and it <= 50
                last = it
