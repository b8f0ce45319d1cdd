"""Golden vectors for the P1 -> P2 midpoint enrichment, produced by EXECUTING THE REAL REFERENCE
(tsx-tunnel/pythonFEM.py:1508-1626, create_midpoints_P2; container only):

    python -m oracle.make_golden_midpoints

Kept apart from oracle/make_golden.py so that the other fixtures are not rewritten.  Meshes: the reference's own tsx-tunnel
mesh (unstructured, 887 triangles) and the level-1 strip-footing mesh (800 triangles).  Under NumPy 2 the reference raises
on a boundary edge met as an element's FIRST edge V2-V3 (:1556 assigns a (3,1) array to a (3,) slice; the other two
branches :1588/:1620 assign (3,) arrays and work), which the tsx mesh never does; the footing mesh's triangles are therefore
rotated cyclically where needed so that their first edge is an interior one (same mesh, same orientation)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rotate_boundary_off_first_edge(elem):
    n = int(elem.max()) + 1
    a = np.concatenate([elem[1], elem[2], elem[0]])
    b = np.concatenate([elem[2], elem[0], elem[1]])
    key = np.minimum(a, b) * n + np.maximum(a, b)
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    boundary = (cnt[inv] == 1).reshape(3, -1)            # [edge slot][element]
    out = elem.copy()
    for e in np.nonzero(boundary[0])[0]:
        for _ in range(2):
            out[:, e] = out[[1, 2, 0], e]
            boundary[:, e] = boundary[[1, 2, 0], e]      # slot s of the rotated triangle was slot s+1
            if not boundary[0, e]:
                break
        assert not boundary[0, e]
    return out


def main():
    rt, rp = ref_loader.load("tsx"), ref_loader.load("plasticity")
    g = np.load(os.path.join(OUT, "assembly_tsx_p1.npz"))
    meshes = {"tsx": (g["coordinates"], g["elements"].astype(np.int64))}
    f = np.load(os.path.join(OUT, "assembly_footing_p1_l1.npz"))
    meshes["footing_l1"] = (f["coordinates"], rotate_boundary_off_first_edge(f["elements"].astype(np.int64)))
    out = {}
    for name, (coord, elem) in meshes.items():
        d = rt.create_midpoints_P2(coord, elem)
        out[f"{name}_coord"], out[f"{name}_elem"] = coord, elem
        for k, v in d.items():
            out[f"{name}_{k}"] = np.asarray(v)
        print(name, {k: np.asarray(v).shape for k, v in d.items()})
    np.savez_compressed(os.path.join(OUT, "midpoints_p2.npz"), **out)


if __name__ == "__main__":
    main()
