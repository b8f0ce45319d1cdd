"""NumPy/SciPy statement of the geometric multigrid V-cycle of fem_elastoplasticity_b200/mg.py + csrc/mg.cu (test helper:
the reference itself has no iterative solver).  Same hierarchy (every second lattice point, ceil), same
bilinear transfers, Galerkin operators, unit diagonal on coarse DOFs without free support, Chebyshev smoothing on block Jacobi (2x2 node blocks) with
the coefficients of mg.chebyshev_coefficients, dense solve on the last level."""
import numpy as np
import scipy.sparse as sp


def interp_1d(n_f, n_c):
    rows, cols, vals = [], [], []
    for i in range(n_f):
        if i % 2 == 0:
            rows.append(i), cols.append(i // 2), vals.append(1.0)
        else:
            rows.extend([i, i]), cols.extend([(i - 1) // 2, (i + 1) // 2]), vals.extend([0.5, 0.5])
    return sp.csr_matrix((vals, (rows, cols)), shape=(n_f, n_c))


def prolongation(nxf, nyf):
    """(2 nxf nyf) x (2 nxc nyc) interpolation between node lattices (node = ix + iy*nx, DOF = 2 node + comp)."""
    nxc, nyc = -(-(nxf - 1) // 2) + 1, -(-(nyf - 1) // 2) + 1
    P = sp.kron(interp_1d(nyf, nyc), interp_1d(nxf, nxc), format="csr")
    return sp.kron(P, sp.identity(2), format="csr"), nxc, nyc


class ReferenceMG:
    def __init__(self, K_setup, q, nx, ny, n_levels, degree, ratio):
        """K_setup: scipy matrix on the lattice-ordered DOFs; q: free-DOF mask; n_levels structured levels."""
        self.q = q.astype(float)
        M = sp.diags(self.q)
        self.degree, self.ratio = degree, ratio
        cur = (M @ K_setup @ M).tocsr()
        self.A, self.P, self.dims = [cur], [], [(nx, ny)]
        for _ in range(n_levels):
            P, nx, ny = prolongation(nx, ny)
            cur = (P.T @ cur @ P).tocsr()                 # the Galerkin chain runs on the raw products ...
            d = cur.diagonal()
            Ac = (cur + sp.diags((d <= 1e-14 * d.max()).astype(float))).tocsr()   # ... a level's own dead DOFs get a unit diagonal
            self.P.append(P), self.A.append(Ac), self.dims.append((nx, ny))
        self.coarse = np.linalg.inv(self.A[-1].toarray())

    def set_fine(self, K):
        M = sp.diags(self.q)
        self.A[0] = (M @ K @ M).tocsr()

    def set_bounds(self, lmax):
        """Chebyshev coefficients from the given eigenvalue bounds of D^-1 A, D = the nodes' 2x2 diagonal blocks (block
        Jacobi); rows/columns of masked DOFs (level 0) are zero, dead coarse DOFs have a unit diagonal and no coupling."""
        from fem_elastoplasticity_b200.mg import chebyshev_coefficients
        self.coef = [chebyshev_coefficients(lm, self.ratio, self.degree) for lm in lmax]
        self.dinv = []
        for l, A in enumerate(self.A[:-1]):
            n = A.shape[0] // 2
            dg = A.diagonal()
            k00, k11 = dg[0::2].copy(), dg[1::2].copy()
            k01 = 0.5 * (A.diagonal(1)[0::2] + A.diagonal(-1)[0::2])
            f0 = (self.q[0::2] != 0) if l == 0 else np.ones(n, dtype=bool)
            f1 = (self.q[1::2] != 0) if l == 0 else np.ones(n, dtype=bool)
            both = f0 & f1
            det = np.where(both, k00 * k11 - k01 * k01, 1.0)
            i00 = np.where(both, k11 / det, np.where(f0 & (k00 != 0), 1.0 / np.where(k00 != 0, k00, 1.0), 0.0))
            i11 = np.where(both, k00 / det, np.where(f1 & (k11 != 0), 1.0 / np.where(k11 != 0, k11, 1.0), 0.0))
            i01 = np.where(both, -k01 / det, 0.0)
            ev, od = np.arange(0, 2 * n, 2), np.arange(1, 2 * n, 2)
            self.dinv.append(sp.csr_matrix((np.concatenate([i00, i01, i01, i11]), (np.concatenate([ev, ev, od, od]), np.concatenate([ev, od, ev, od]))),
                                           shape=A.shape))

    def smooth(self, l, b, x):
        A, dinv, (c1, c2) = self.A[l], self.dinv[l], self.coef[l]
        d = np.zeros_like(b)
        for k in range(self.degree):
            r = b if x is None else b - A @ x
            d = (c1[k] * d if k else 0.0) + c2[k] * (dinv @ r)
            x = d if x is None else x + d
        return x

    def vcycle(self, b, l=0):
        if l == len(self.A) - 1:
            return self.coarse @ b
        x = self.smooth(l, b, None)
        r = b - self.A[l] @ x
        if l == 0:
            r = r * self.q
        xc = self.vcycle(self.P[l].T @ r, l + 1)
        corr = self.P[l] @ xc
        x = x + (corr * self.q if l == 0 else corr)
        return self.smooth(l, b, x)
