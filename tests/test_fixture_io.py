"""On-disk formats either side of the path, on the CPU: mesh CSV reader, Dirichlet mask of the tsx
driver, and the K[Q,Q] / F[Q] exporters in the layout of the reference's MATLAB-exported CSVs."""
import numpy as np
import scipy.sparse as sp

from fem_elastoplasticity_b200 import fixture_io as fio


def test_qq_block_layout_matches_reference_csv(golden, tmp_path):
    a, g = golden("assembly_tsx_p1.npz"), golden("tsx_csv_golden.npz")
    n = a["K_indptr"].size - 1
    K = sp.csr_matrix((a["K_data"], a["K_indices"], a["K_indptr"]), shape=(n, n))       # the reference's own K_elast
    q = fio.tsx_dirichlet_mask(a["coordinates"])
    assert q.shape == a["coordinates"].shape and int(q.sum()) == 908                     # 908 free DOFs
    kqq = fio.free_block(K, q)
    ref = sp.csr_matrix((g["kqq_data"], g["kqq_indices"], g["kqq_indptr"]), shape=tuple(g["kqq_shape"])).toarray()
    assert kqq.shape == ref.shape == (908, 908)
    assert np.array_equal(kqq != 0, ref != 0)                                            # same free-DOF order, same pattern
    assert np.abs(kqq - ref).max() <= 2e-4 * np.abs(ref).max()                           # the CSV holds ~5 digits
    fio.write_qq_csv(tmp_path / "kelast_qq.csv", K, q)                                   # the fixture the reference lacks
    assert np.array_equal(fio.read_qq_csv(tmp_path / "kelast_qq.csv"), kqq)              # %.17g round-trips doubles
    F = np.arange(n, dtype=float) * 0.1
    fio.write_fq_csv(tmp_path / "fq.csv", F, q)
    back = np.atleast_1d(np.genfromtxt(tmp_path / "fq.csv", delimiter=","))
    assert back.shape == g["fq"].shape and np.array_equal(back, F[q.flatten(order="F")])


def test_mesh_csv_reader_is_one_based(golden, tmp_path):
    a = golden("assembly_tsx_p1.npz")
    np.savetxt(tmp_path / "coord.csv", a["coordinates"], delimiter=",", fmt="%.17g")
    np.savetxt(tmp_path / "elem.csv", a["elements"] + 1, delimiter=",", fmt="%d")          # the files hold 1-based ids (:1688)
    coords, elem = fio.read_mesh_csv(tmp_path / "coord.csv", tmp_path / "elem.csv")
    assert np.array_equal(coords, a["coordinates"]) and np.array_equal(elem, a["elements"])
    assert elem.min() == 0
