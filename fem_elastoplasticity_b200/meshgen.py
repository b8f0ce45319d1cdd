"""Synthetic uniform P1 meshes built directly in HBM (BASELINE.json configs 4/5) and the P1 -> P2 midpoint enrichment.

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


def _s64(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def hash_uniform(counter):
    """Counter-based uniform numbers in (0, 1): splitmix64 of an int64 tensor of counters (wrap-around int64 arithmetic,
    logical shifts emulated by masking).  A value depends only on its counter, so ranks of a partitioned mesh that share
    an element or a node generate identical data for it, whatever their local numbering."""
    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)
    z = counter + _s64(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * _s64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    return (lsr(z, 11).to(torch.float64) + 0.5) * (1.0 / float(1 << 53))


def hash_normal(ids, stream, seed=0):
    """Standard normal numbers (Box-Muller on two hashed uniforms) indexed by global ids: value = f(seed, stream, id)."""
    base = (ids.to(torch.int64) * 16 + 2 * int(stream)) + int(seed) * 1000003 * 16
    u1, u2 = hash_uniform(base), hash_uniform(base + 1)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * np.pi * u2)


def synthetic_strain_global(n_int, first_id=0, device="cuda", seed=0, mean=(-3e-4, -3e-4, 0.0), std=2e-4):
    """Like ``synthetic_strain`` (same distribution: ~2 % plastic / 0.6 % apex at the footing constants) but a function of
    the GLOBAL integration-point id ``first_id + local index``: the ghost cell row a rank keeps for its interface nodes
    carries exactly the strains its owner has, so the partitioned matrix is the matrix of one global problem."""
    ids = torch.arange(first_id, first_id + n_int, dtype=torch.int64, device=device)
    return torch.stack([hash_normal(ids, k, seed) * std + mean[k] for k in range(3)])


def synthetic_nodal_global(n_n, first_node=0, device="cuda", seed=1, scale=1e-3):
    """DOF vector (2 n_n,) of hashed normals indexed by global node id (consistent on ghost node rows)."""
    ids = torch.arange(first_node, first_node + n_n, dtype=torch.int64, device=device)
    return (torch.stack([hash_normal(ids, 8 + k, seed) for k in range(2)], dim=1).reshape(-1) * scale).contiguous()


def create_midpoints_p2(coord, elem):
    """P1 -> P2 enrichment with the numbering of the reference's create_midpoints_P2 (tsx-tunnel/pythonFEM.py:1508-1626),
    without its O(n_e^2) ``np.where`` walk: every edge is keyed by its vertex pair, a midpoint is numbered by the FIRST
    (element, edge) visit of the reference's loop (edges in the order V2-V3, V3-V1, V1-V2), i.e. by the rank of the smallest
    occurrence index ``3*element + edge`` among the unique keys - one sort instead of a search per edge.  Runs on whatever
    device ``coord``/``elem`` live on (torch tensors; (2, n_n) float64, (3, n_e) integer, 0-based, counter-clockwise).

    Returns tensors named like the reference's dict: coord_mid (2, n_mid), surf (3, n_boundary_edges) = (second vertex,
    first vertex, midpoint) per boundary edge, coord_ext, elem_ext (6, n_e), elem_ed (3, n_e) midpoint index per element
    edge, edge_el (2, 2*n_e) the elements on either side of each midpoint (0 where the reference leaves its zeros).
    The reference finds the neighbour's edge slot from the position of one shared vertex, which presumes that the two
    triangles traverse the edge in opposite directions; a mesh where they do not is rejected here.

    CUDA tensors take the hand-written kernels (csrc/midpoints.cu: hash table of the edge keys + prefix sums, no sort); the
    torch formulation below is the host-side statement of the same rule (CPU tensors only)."""
    if coord.is_cuda:
        return _create_midpoints_p2_cuda(coord, elem)
    elem = elem.to(torch.int64)
    dev = elem.device
    n_e, n_n = elem.shape[1], coord.shape[1]
    a = torch.stack([elem[1], elem[2], elem[0]], dim=1).reshape(-1)      # occurrence o = 3*element + edge: from vertex
    b = torch.stack([elem[2], elem[0], elem[1]], dim=1).reshape(-1)      # ... to vertex
    key = torch.minimum(a, b) * n_n + torch.maximum(a, b)
    uniq, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    if int(cnt.max()) > 2:
        raise ValueError("an edge is shared by more than two triangles")
    occ = torch.arange(3 * n_e, device=dev)
    first = torch.full((uniq.numel(),), 3 * n_e, dtype=torch.int64, device=dev).scatter_reduce(0, inv, occ, reduce="amin")
    last = torch.full((uniq.numel(),), -1, dtype=torch.int64, device=dev).scatter_reduce(0, inv, occ, reduce="amax")
    shared = cnt == 2
    if not bool(((a[first] == b[last]) & (b[first] == a[last]))[shared].all()):
        raise ValueError("inconsistently oriented triangles: the reference's neighbour-slot rule is undefined")
    order = torch.argsort(first)                                         # unique edges in the order the reference meets them
    n_mid = order.numel()
    ind_of_edge = torch.empty_like(order)
    ind_of_edge[order] = torch.arange(n_mid, device=dev)
    ind = ind_of_edge[inv].reshape(n_e, 3).t()                           # (3, n_e): midpoint index of every element edge
    fa, fb = a[first[order]], b[first[order]]
    coord_mid = (coord[:, fa] + coord[:, fb]) / 2
    edge_el = torch.zeros((2, max(2 * n_e, n_mid)), dtype=torch.float64, device=dev)   # reference: (2, 2 n_e), enough for n_e >= boundary edges
    edge_el[0, :n_mid] = (first[order] // 3).to(torch.float64)
    edge_el[1, :n_mid] = torch.where(shared[order], last[order] // 3, torch.zeros_like(order)).to(torch.float64)
    bnd = ~shared[order]
    mids = torch.arange(n_mid, device=dev)[bnd] + n_n
    surf = torch.stack([fb[bnd], fa[bnd], mids]).to(torch.float64)
    return {"coord_mid": coord_mid, "surf": surf, "coord_ext": torch.cat([coord, coord_mid], dim=1),
            "elem_ext": torch.cat([elem, ind + n_n], dim=0), "elem_ed": ind.to(torch.float64), "edge_el": edge_el}


def _create_midpoints_p2_cuda(coord, elem):
    """create_midpoints_p2 through fem_midpoints_p2_count / _fill (include/fem_b200.h)."""
    import ctypes as C
    from ._lib import call
    from .plan import _ptr, _stream
    dev = coord.device
    n_n, n_e = coord.shape[1], elem.shape[1]
    coord = coord.to(torch.float64).contiguous()
    e32 = elem.to(device=dev, dtype=torch.int32).contiguous()
    handle, n_mid, n_bnd, status = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int()
    with torch.cuda.device(dev):
        call("fem_midpoints_p2_count", n_n, n_e, _ptr(e32), C.byref(handle), C.byref(n_mid), C.byref(n_bnd), C.byref(status), _stream())
        try:
            if status.value & 1:
                raise ValueError("an edge is shared by more than two triangles")
            if status.value & 2:
                raise ValueError("inconsistently oriented triangles: the reference's neighbour-slot rule is undefined")
            n_mid, n_bnd = n_mid.value, n_bnd.value
            coord_mid = torch.empty((2, n_mid), dtype=torch.float64, device=dev)
            ind = torch.empty((3, n_e), dtype=torch.int32, device=dev)
            edge_el32 = torch.empty((2, n_mid), dtype=torch.int32, device=dev)
            surf32 = torch.empty((3, n_bnd), dtype=torch.int32, device=dev)
            call("fem_midpoints_p2_fill", handle, _ptr(coord), _ptr(coord_mid), _ptr(ind), _ptr(edge_el32), _ptr(surf32), _stream())
        finally:
            call("fem_midpoints_p2_destroy", handle, _stream())
    ind = ind.to(torch.int64)
    edge_el = torch.zeros((2, max(2 * n_e, n_mid)), dtype=torch.float64, device=dev)
    edge_el[:, :n_mid] = edge_el32.to(torch.float64)
    return {"coord_mid": coord_mid, "surf": surf32.to(torch.float64), "coord_ext": torch.cat([coord, coord_mid], dim=1),
            "elem_ext": torch.cat([elem.to(device=dev, dtype=torch.int64), ind + n_n], dim=0), "elem_ed": ind.to(torch.float64),
            "edge_el": edge_el}
