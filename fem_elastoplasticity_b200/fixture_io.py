"""On-disk formats either side of the hot path: the tsx-tunnel mesh CSVs and the MATLAB-exported
golden arrays of the reference (tsx-tunnel/{coord,elem,k_tangent_qq,fq,f0q}.csv).

    coord.csv  (2, n_n)  node coordinates;   elem.csv  (n_p, n_e)  1-based node ids   (tsx-tunnel/pythonFEM.py:1687-1688)
    *_qq.csv   dense K[Q,Q] (free-DOF block, node-major interleaved DOF order)        (:1742-1746)
    fq.csv / f0q.csv   F[Q] as a single row                                           (:1747)
"""
import numpy as np


def read_mesh_csv(coord_path, elem_path):
    """-> (coordinates (2,n_n) float64, elements (n_p,n_e) int64 0-based), as the reference does at :1687-1688."""
    coords = np.genfromtxt(coord_path, delimiter=',')
    elem = np.genfromtxt(elem_path, delimiter=',', dtype=int) - 1
    return coords, elem


def tsx_dirichlet_mask(coords, limit=49.99):
    """Q (2,n_n) bool: u_x fixed where |x| > limit, u_y fixed where |y| > limit (:1695-1699)."""
    q = np.ones(coords.shape, dtype=bool)
    q[0, np.abs(coords[0]) > limit] = False
    q[1, np.abs(coords[1]) > limit] = False
    return q


def free_block(K, q_mask):
    """K[Q,Q] as a dense array in the golden layout (free DOFs in node-major interleaved order)."""
    qf = np.asarray(q_mask).flatten(order='F')
    return K.tocsr()[qf][:, qf].toarray()


def write_qq_csv(path, K, q_mask):
    """Export K[Q,Q] in the layout of k_tangent_qq.csv / the missing kelast_qq.csv."""
    np.savetxt(path, free_block(K, q_mask), delimiter=',', fmt='%.17g')


def write_fq_csv(path, F, q_mask):
    """Export F[Q] in the layout of fq.csv / f0q.csv (one row)."""
    qf = np.asarray(q_mask).flatten(order='F')
    np.savetxt(path, np.asarray(F).reshape(-1)[qf][None, :], delimiter=',', fmt='%.17g')


def read_qq_csv(path):
    return np.genfromtxt(path, delimiter=',')
