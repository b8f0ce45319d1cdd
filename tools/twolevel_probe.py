"""CPU probe of the two-level preconditioner's ALGORITHM (NumPy/SciPy, test infrastructure: uses oracle/): does a coarse
grid with hx != hy / non-nested cells, or the partitioned Galerkin product (owned rows x free columns summed over strips),
change the iteration count?  Written to bisect the one 8-GPU weak-scaling run that did not converge.

    python tools/twolevel_probe.py --nx 40 --ny 320 --size-y 80 --grids 5x40 5x44 --world 1 4
"""
import argparse
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fem_oracle as fo  # noqa: E402


def strip_problem(nx, ny, sx, sy):
    m = fo.square_mesh_p1(nx, ny, sx, sy)
    coord, elem = m["coordinates"], m["elements"]
    xi, wf = fo.quadrature_volume(fo.ElementType.P1)
    _, d1, d2 = fo.local_basis_volume(fo.ElementType.P1, xi)
    n_int = elem.shape[1] * wf.size
    young, poisson = 1e7, 0.48
    G = np.full(n_int, young / (2 * (1 + poisson)))
    Kb = np.full(n_int, young / (3 * (1 - 2 * poisson)))
    K = fo.elastic_stiffness(elem, coord, G, Kb, d1, d2, wf)[0].tocsr()
    return coord, K, m["Q"].astype(bool).flatten(order="F")


def prolongation(coord, free, ncx, ncy):
    x0, y0 = coord[0].min(), coord[1].min()
    hx, hy = (coord[0].max() - x0) / ncx, (coord[1].max() - y0) / ncy
    fx, fy = (coord[0] - x0) / hx, (coord[1] - y0) / hy
    cx, cy = np.clip(fx.astype(int), 0, ncx - 1), np.clip(fy.astype(int), 0, ncy - 1)
    xi, et = np.clip(fx - cx, 0, 1), np.clip(fy - cy, 0, 1)
    w = np.stack([(1 - xi) * (1 - et), xi * (1 - et), xi * et, (1 - xi) * et])
    base = cx + cy * (ncx + 1)
    ids = np.stack([base, base + 1, base + 1 + (ncx + 1), base + (ncx + 1)])
    n_n = coord.shape[1]
    rows, cols, vals = [], [], []
    for c in range(2):
        for k in range(4):
            rows.append(2 * np.arange(n_n) + c)
            cols.append(2 * ids[k] + c)
            vals.append(w[k])
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * n_n, 2 * (ncx + 1) * (ncy + 1)))
    return sp.diags(free.astype(float)) @ P


def pcg(K, b, free, apply_m, rtol=1e-10, maxit=20000):
    x = np.zeros_like(b)
    r = np.where(free, b, 0.0)
    z = apply_m(r)
    p = z.copy()
    rz = r @ z
    b2 = r @ r
    for it in range(1, maxit + 1):
        q = np.where(free, K @ p, 0.0)
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if r @ r <= rtol * rtol * b2:
            return it
        z = apply_m(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return maxit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=40)
    ap.add_argument("--ny", type=int, default=320)
    ap.add_argument("--size-x", type=float, default=10.0)
    ap.add_argument("--size-y", type=float, default=80.0)
    ap.add_argument("--grids", nargs="+", default=["5x40", "5x44"])
    ap.add_argument("--world", type=int, nargs="+", default=[1, 4])
    a = ap.parse_args()
    coord, K, free = strip_problem(a.nx, a.ny, a.size_x, a.size_y)
    n = K.shape[0]
    b = np.random.default_rng(0).standard_normal(n)
    dinv = np.where(free, 1.0 / K.diagonal(), 0.0)
    print("jacobi:", pcg(K, b, free, lambda r: dinv * r))
    for gs in a.grids:
        ncx, ncy = map(int, gs.split("x"))
        P = prolongation(coord, free, ncx, ncy)
        for world in a.world:
            # partitioned product: rank w owns node rows (w*ny_loc, (w+1)*ny_loc] (rank 0 also row 0); rows of K restricted
            # to the owned nodes times ALL free columns, summed over the ranks
            ny_loc = a.ny // world
            row_of_node = np.arange(coord.shape[1]) // (a.nx + 1)
            Ac = np.zeros((P.shape[1], P.shape[1]))
            for w in range(world):
                own = (row_of_node > w * ny_loc) & (row_of_node <= (w + 1) * ny_loc)
                if w == 0:
                    own |= row_of_node == 0
                own2 = np.repeat(own, 2).astype(float)
                Ac += ((sp.diags(own2) @ P).T @ (K @ P)).toarray()
            Ac = 0.5 * (Ac + Ac.T)
            d = np.diag(Ac).copy()
            dead = d <= 1e-14 * d.max()
            Ac[dead, :] = 0
            Ac[:, dead] = 0
            Ac[dead, dead] = 1
            Aci = np.linalg.inv(Ac)
            Aci[dead, :] = 0
            Aci[:, dead] = 0
            its = pcg(K, b, free, lambda r: dinv * r + P @ (Aci @ (P.T @ r)))
            print(f"grid {gs} world {world}: n_c={P.shape[1]} dead={int(dead.sum())} cond(Ac)={np.linalg.cond(Ac):.2e} iterations={its}")


if __name__ == "__main__":
    main()
