"""Synthetic uniform P1 meshes built directly in HBM (BASELINE.json configs 4/5).

Numbering is the reference's get_nodes_1 (Plasticity2D_DP/pythonFEM.py:73-122): node id = ix + iy*(n_x+1),
cell (ix, iy) -> triangles (V1,V2,V4), (V2,V3,V4), cell-major with ix fastest; footing boundary conditions
(:178-184).  Coordinates come from numpy.linspace on the host (n+1 values per axis) so they are bit-identical to the
reference generator's; only the O(n_n) tiling happens on the device."""
import numpy as np
import torch


def square_mesh_p1(n_x, n_y, size_x=10.0, size_y=10.0, device="cuda", iy0=0, n_y_global=None, size_y_global=None):
    """Rows iy0 .. iy0+n_y of cells of a (n_x x n_y_global) mesh.  Returns dict of CUDA tensors:
    elements (3, n_e) int32 (local node ids), coordinates (2, n_n) f64, Q (2, n_n) bool, dirichlet_nodes (2, n_n) f64."""
    n_y_global = n_y if n_y_global is None else n_y_global
    size_y_global = size_y if size_y_global is None else size_y_global
    dev = torch.device(device)
    cx = torch.as_tensor(np.linspace(0, size_x, n_x + 1)).to(dev)
    cy = torch.as_tensor(np.linspace(0, size_y_global, n_y_global + 1)[iy0:iy0 + n_y + 1].copy()).to(dev)
    coord = torch.stack([cx.repeat(n_y + 1), cy.repeat_interleave(n_x + 1)])
    ix = torch.arange(n_x, device=dev, dtype=torch.int32)
    iy = torch.arange(n_y, device=dev, dtype=torch.int32)
    v1 = (ix[None, :] + iy[:, None] * (n_x + 1)).reshape(-1)
    elem = torch.empty((3, 2 * n_x * n_y), dtype=torch.int32, device=dev)
    elem[0, 0::2], elem[1, 0::2], elem[2, 0::2] = v1, v1 + 1, v1 + (n_x + 1)
    elem[0, 1::2], elem[1, 1::2], elem[2, 1::2] = v1 + 1, v1 + (n_x + 2), v1 + (n_x + 1)
    top_left = (coord[1] == size_y_global) & (coord[0] <= 1.0001)
    q = coord > 0
    q[1, top_left] = False
    q[0, coord[0] == size_x] = False
    dirichlet = torch.zeros_like(coord)
    dirichlet[1, top_left] = 1.0
    return {"elements": elem, "coordinates": coord, "Q": q, "dirichlet_nodes": dirichlet, "n_x": n_x, "n_y": n_y}


def footing_materials(n_int, device="cuda"):
    """Constant material arrays of the strip-footing benchmark (Plasticity2D_DP/pythonFEM.py:910-933)."""
    young, poisson, c0, phi = 1e7, 0.48, 450, np.pi / 9
    shear = young / (2 * (1 + poisson))
    bulk = young / (3 * (1 - 2 * poisson))
    eta = 3 * np.tan(phi) / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    c = 3 * c0 / np.sqrt(9 + 12 * (np.tan(phi)) ** 2)
    full = lambda v: torch.full((n_int,), float(v), dtype=torch.float64, device=device)  # noqa: E731
    return full(shear), full(bulk), full(eta), full(c)


def synthetic_strain(n_int, device="cuda", seed=0, mean=(-3e-4, -3e-4, 0.0), std=2e-4):
    """E = mean + std*randn(3, n_int): ~2 % plastic / 0.6 % apex at the footing constants."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    e = torch.randn((3, n_int), generator=g, dtype=torch.float64, device=device) * std
    e += torch.tensor(mean, dtype=torch.float64, device=device)[:, None]
    return e
