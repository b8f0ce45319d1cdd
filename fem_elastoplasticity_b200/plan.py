"""Device-side objects of the hot path: FemPlan (mesh -> CSR pattern + geometry + kernels) and the
Drucker-Prager return map.  PyTorch is used only to own device buffers and streams; all arithmetic
is done by the CUDA kernels behind the C ABI (include/fem_b200.h)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import call


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


_UPLOAD_MIN_BYTES = 64 << 20     # pageable arrays of at least this size take the staged upload
_UPLOAD_CHUNK = 4 << 20          # doubles per staging chunk (32 MB)
_UPLOAD_THREADS = 6
_upload_state = {}


def _upload_pageable(a, device):
    """Large pageable NumPy array -> device tensor through two page-locked staging chunks: a few threads copy chunk k + 1 into
    its staging buffer (NumPy releases the GIL for large copies) while the DMA of chunk k runs.  A plain `.to(device)` of
    pageable memory goes through the driver's single-threaded bounce buffers at ~10 GB/s; this reaches the host's memcpy rate."""
    from concurrent.futures import ThreadPoolExecutor
    st = _upload_state.get(device)
    if st is None:
        st = _upload_state[device] = {
            "pool": ThreadPoolExecutor(_UPLOAD_THREADS),
            "stage": [torch.empty(_UPLOAD_CHUNK, dtype=torch.float64, pin_memory=True) for _ in range(2)],
            "ev": [torch.cuda.Event() for _ in range(2)], "used": [False, False]}
    flat = a.reshape(-1)
    n = flat.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=device)
    for k, start in enumerate(range(0, n, _UPLOAD_CHUNK)):
        j = k & 1
        m = min(_UPLOAD_CHUNK, n - start)
        if st["used"][j]:
            st["ev"][j].synchronize()                      # the DMA that last read this staging buffer is done
        dst = st["stage"][j].numpy()
        step = -(-m // _UPLOAD_THREADS)
        futs = [st["pool"].submit(np.copyto, dst[o:min(o + step, m)], flat[start + o:start + min(o + step, m)]) for o in range(0, m, step)]
        for f in futs:
            f.result()
        out[start:start + m].copy_(st["stage"][j][:m], non_blocking=True)
        st["ev"][j].record()
        st["used"][j] = True
    return out.reshape(a.shape)


def _dev_f64(x, device, shape=None):
    """Host array / tensor -> contiguous float64 CUDA tensor (no copy when already there)."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.float64)
    else:
        a = np.ascontiguousarray(x, dtype=np.float64)
        h = torch.as_tensor(a)
        t = None
        if a.nbytes >= _UPLOAD_MIN_BYTES and not h.is_pinned():
            try:
                t = _upload_pageable(a, torch.device(device))
            except RuntimeError:               # no page-locked staging memory on this host: ordinary upload
                t = None
        if t is None:
            t = h.to(device)
    t = t.contiguous()
    if shape is not None:
        t = t.reshape(shape)
    return t


class _DevView:
    """Zero-copy torch view of plan-owned device memory (kept alive by the plan)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("fem_elastoplasticity_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    _lib.load()


class FemPlan:
    """Mesh topology + reference-element tables -> structural CSR pattern, incidence lists, geometry.

    ``elements`` is (n_p, n_e) 0-based, ``coordinates`` (2, n_n), ``dhatp1/dhatp2`` (n_p, n_q), ``wf`` (1, n_q)
    exactly as passed to the reference's get_elastic_stiffness_matrix
    (Plasticity2D_DP/pythonFEM.py:491-497)."""

    def __init__(self, elements, coordinates, dhatp1, dhatp2, wf, device=None):
        require_cuda()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if isinstance(elements, torch.Tensor):
            el = elements.to(device=self.device, dtype=torch.int32).contiguous()
        else:
            el = torch.as_tensor(np.ascontiguousarray(np.asarray(elements).astype(np.int32))).to(self.device)
        co = _dev_f64(coordinates, self.device)
        if el.dim() != 2 or co.dim() != 2 or co.shape[0] != 2:
            raise ValueError("elements must be (n_p, n_e) and coordinates (2, n_n)")
        self.n_p, self.n_e = int(el.shape[0]), int(el.shape[1])
        self.n_n = int(co.shape[1])
        self.coord = co                                   # (2, n_n) device coordinates (two-level preconditioner)
        d1 = np.ascontiguousarray(np.asarray(dhatp1, dtype=np.float64).reshape(self.n_p, -1))
        d2 = np.ascontiguousarray(np.asarray(dhatp2, dtype=np.float64).reshape(self.n_p, -1))
        w = np.ascontiguousarray(np.asarray(wf, dtype=np.float64).ravel())
        self.n_q = int(w.size)
        if d1.shape != (self.n_p, self.n_q) or d2.shape != (self.n_p, self.n_q):
            raise ValueError("dhatp1/dhatp2 must be (n_p, n_q)")
        self._h = C.c_void_p()
        dp = C.POINTER(C.c_double)
        with torch.cuda.device(self.device):
            call("fem_plan_create", self.n_n, self.n_e, self.n_p, self.n_q, _ptr(el), _ptr(co), d1.ctypes.data_as(dp),
                 d2.ctypes.data_as(dp), w.ctypes.data_as(dp), _stream(), C.byref(self._h))
        sz = [C.c_int64() for _ in range(5)]
        md = C.c_int()
        call("fem_plan_sizes", self._h, *[C.byref(s) for s in sz], C.byref(md))
        self.n_n, self.n_e, self.n_int, self.n_dof, self.nnz = [int(s.value) for s in sz]
        self.max_degree = int(md.value)
        rp, ci, nn = C.c_void_p(), C.c_void_p(), C.c_int64()
        call("fem_plan_pattern", self._h, C.byref(rp), C.byref(ci), C.byref(nn))
        self.row_ptr = torch.as_tensor(_DevView(rp.value, (self.n_dof + 1,), "<i4"), device=self.device)
        self.col_idx = torch.as_tensor(_DevView(ci.value, (max(self.nnz, 1),), "<i4"), device=self.device)[: self.nnz]
        bp, bi, nb = C.c_void_p(), C.c_void_p(), C.c_int64()
        call("fem_plan_blocks", self._h, C.byref(bp), C.byref(bi), C.byref(nb))
        self.n_blocks = int(nb.value)
        self.nbr_ptr = torch.as_tensor(_DevView(bp.value, (self.n_n + 1,), "<i4"), device=self.device)
        self.nbr_idx = torch.as_tensor(_DevView(bi.value, (max(self.n_blocks, 1),), "<i4"), device=self.device)[: self.n_blocks]
        g1, g2, gw = C.c_void_p(), C.c_void_p(), C.c_void_p()
        call("fem_plan_geometry", self._h, C.byref(g1), C.byref(g2), C.byref(gw))
        self.dphi1 = torch.as_tensor(_DevView(g1.value, (self.n_p, self.n_int), "<f8"), device=self.device)
        self.dphi2 = torch.as_tensor(_DevView(g2.value, (self.n_p, self.n_int), "<f8"), device=self.device)
        self.weight = torch.as_tensor(_DevView(gw.value, (self.n_int,), "<f8"), device=self.device)
        self.bytes = int(_lib.load().fem_plan_bytes(self._h))

    def stage_info(self):
        """(stage_ok, TMA box width in elements) of the TMA-staged assembly kernel."""
        ok, cap = C.c_int(), C.c_int()
        call("fem_plan_stage_info", self._h, C.byref(ok), C.byref(cap))
        return int(ok.value), int(cap.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            for name in ("row_ptr", "col_idx", "nbr_ptr", "nbr_idx", "dphi1", "dphi2", "weight"):
                setattr(self, name, None)
            _lib.load().fem_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------
    def _f64(self, x, shape=None):
        return _dev_f64(x, self.device, shape)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def mask_u8(self, q_mask):
        """(2, n_n) boolean mask (reference layout) or flat (n_dof,) -> uint8 device vector in DOF order."""
        if isinstance(q_mask, torch.Tensor):
            m = q_mask.to(self.device)
            m = m.t().reshape(-1) if m.dim() == 2 else m.reshape(-1)
            return m.to(torch.uint8).contiguous()
        m = np.asarray(q_mask)
        m = m.flatten(order="F") if m.ndim == 2 else m.ravel()
        return torch.as_tensor(np.ascontiguousarray(m.astype(np.uint8))).to(self.device)

    # -- assembly ------------------------------------------------------------------------------
    def assemble_elastic(self, shear, bulk, out=None):
        sh, bu = self._f64(shear, (self.n_int,)), self._f64(bulk, (self.n_int,))
        out = self.empty(self.nnz) if out is None else out
        call("fem_assemble_elastic", self._h, _ptr(sh), _ptr(bu), _ptr(out), _stream())
        return out

    def elastic_dmat(self, shear, bulk):
        sh, bu = self._f64(shear, (self.n_int,)), self._f64(bulk, (self.n_int,))
        vd = self.empty(9, self.n_int)
        call("fem_elastic_dmat", self._h, _ptr(sh), _ptr(bu), _ptr(vd), _stream())
        return vd

    def assemble_tangent(self, ds, out=None):
        ds = self._f64(ds, (9, self.n_int))
        out = self.empty(self.nnz) if out is None else out
        call("fem_assemble_tangent", self._h, _ptr(ds), _ptr(out), _stream())
        return out

    def assemble_tangent_ref(self, ds, shear, bulk, k_elast_vals, out=None):
        ds = self._f64(ds, (9, self.n_int))
        sh, bu = self._f64(shear, (self.n_int,)), self._f64(bulk, (self.n_int,))
        out = self.empty(self.nnz) if out is None else out
        call("fem_assemble_tangent_ref", self._h, _ptr(ds), _ptr(sh), _ptr(bu), _ptr(k_elast_vals), _ptr(out), _stream())
        return out

    def assemble_tangent_force(self, ds, s, out_k=None, out_f=None):
        ds = self._f64(ds, (9, self.n_int))
        s = self._f64(s)
        out_k = self.empty(self.nnz) if out_k is None else out_k
        out_f = self.empty(self.n_dof) if out_f is None else out_f
        call("fem_assemble_tangent_force", self._h, _ptr(ds), _ptr(s), _ptr(out_k), _ptr(out_f), _stream())
        return out_k, out_f

    def strain(self, u, out=None):
        u = self._f64(u, (self.n_dof,))
        out = self.empty(3, self.n_int) if out is None else out
        call("fem_strain", self._h, _ptr(u), _ptr(out), _stream())
        return out

    def internal_force(self, s, out=None):
        s = self._f64(s)
        if s.dim() != 2 or s.shape[0] < 3 or s.shape[1] != self.n_int:
            raise ValueError("s must be (>=3, n_int)")
        out = self.empty(self.n_dof) if out is None else out
        call("fem_internal_force", self._h, _ptr(s), _ptr(out), _stream())
        return out

    # -- SpMV / PCG ----------------------------------------------------------------------------
    def spmv(self, k_vals, x, mask=None, out=None, dot=None):
        x = self._f64(x, (self.n_dof,))
        out = self.empty(self.n_dof) if out is None else out
        call("fem_spmv", self._h, _ptr(k_vals), _ptr(x), _ptr(out), _ptr(mask), _ptr(dot), _stream())
        return out

    def jacobi(self, k_vals, mask=None, out=None):
        out = self.empty(self.n_dof) if out is None else out
        call("fem_jacobi_setup", self._h, _ptr(k_vals), _ptr(mask), _ptr(out), _stream())
        return out

    def pcg(self, k_vals, rhs, mask=None, rtol=1e-10, maxit=100000, check_every=50, x0=None, work=None, raise_on_maxit=True, refine=0):
        """Jacobi-PCG on K[Q,Q]; returns (x, iterations, relative residual).
        ``refine`` > 0 adds that many steps of iterative refinement: the TRUE residual mask*(rhs - K x) is formed in FP64 and
        the correction equation is solved by the same PCG.  The recurrence residual of CG drifts from the true one near
        1e-13; refinement brings the solution to the cond(K)*eps accuracy of the reference's dense LU
        (Plasticity2D_DP/pythonFEM.py:1066), which the 1e-12 displacement parity needs."""
        rhs = self._f64(rhs, (self.n_dof,))
        x = torch.zeros(self.n_dof, dtype=torch.float64, device=self.device) if x0 is None else self._f64(x0, (self.n_dof,)).clone()
        work = self.empty(4 * self.n_dof) if work is None else work
        it, rel = C.c_int(), C.c_double()
        code = _lib.load().fem_pcg(self._h, _ptr(k_vals), _ptr(rhs), _ptr(mask), float(rtol), int(maxit), int(check_every),
                                   _ptr(x), _ptr(work), C.byref(it), C.byref(rel), _stream())
        if code != 0 and not (code == _lib.FEM_ERR_PCG_MAXIT and not raise_on_maxit):
            _lib.check(code)
        its, relres = int(it.value), float(rel.value)
        for _ in range(int(refine)):
            res = axpby(1.0, rhs, -1.0, self.spmv(k_vals, x, mask=mask))       # rows outside the mask are ignored by fem_pcg
            d = torch.zeros_like(x)
            code = _lib.load().fem_pcg(self._h, _ptr(k_vals), _ptr(res), _ptr(mask), min(1e-4, float(rtol) * 1e6), int(maxit), int(check_every),
                                       _ptr(d), _ptr(work), C.byref(it), C.byref(rel), _stream())
            if code != 0 and code != _lib.FEM_ERR_PCG_MAXIT:
                _lib.check(code)
            x += d
            its += int(it.value)
        return x, its, relres

    def energy_norms(self, k_vals, v0, v1, v2, work=None):
        """(v0'Kv0, v1'Kv1, v2'Kv2) as a device tensor of 3 doubles (Plasticity2D_DP/pythonFEM.py:1072-1074)."""
        work = self.empty(self.n_dof) if work is None else work
        out = self.empty(3)
        call("fem_energy_norms", self._h, _ptr(k_vals), _ptr(v0), _ptr(v1), _ptr(v2), _ptr(work), _ptr(out), _stream())
        return out

    def transform(self, q_int, out=None):
        q = self._f64(q_int, (self.n_int,))
        out = self.empty(self.n_n) if out is None else out
        call("fem_transform", self._h, _ptr(q), _ptr(out), _stream())
        return out

    # -- host views (façade / tests) -------------------------------------------------------------
    def pattern_host(self):
        return self.row_ptr.cpu().numpy().copy(), self.col_idx.cpu().numpy().copy()

    def to_scipy_csr(self, k_vals):
        import scipy.sparse as sp
        rp, ci = self.pattern_host()
        return sp.csr_matrix((k_vals.detach().cpu().numpy(), ci, rp), shape=(self.n_dof, self.n_dof))


def axpby(a, x, b, y, out=None):
    """out = a*x + b*y on the device (fem_vec_axpby)."""
    out = torch.empty_like(x) if out is None else out
    call("fem_vec_axpby", x.numel(), float(a), _ptr(x), float(b), _ptr(y), _ptr(out), _stream())
    return out


def dp_return_map(e, ep_prev, shear, bulk, eta, c, apply_plastic_strain=False, e0=None, want_lambda=False,
                  want_ep=True, want_counts=True, out=None, device=None):
    """Device Drucker-Prager return map.  Inputs are (3,n)/(4,n)/(n,) arrays or CUDA tensors; returns a dict of
    CUDA tensors: s (4,n), ds (9,n), ind_p (n,) uint8, ep (4,n), lambda (n,), counts (2,) int64 = (n_smooth, n_apex)."""
    require_cuda()
    dev = torch.device(device) if device is not None else (e.device if isinstance(e, torch.Tensor) and e.is_cuda
                                                          else torch.device(f"cuda:{torch.cuda.current_device()}"))
    E = _dev_f64(e, dev)
    n = int(E.shape[1])
    G, K = _dev_f64(shear, dev, (n,)), _dev_f64(bulk, dev, (n,))
    et, cc = _dev_f64(eta, dev, (n,)), _dev_f64(c, dev, (n,))
    ep_in = None if ep_prev is None else (ep_prev if (isinstance(ep_prev, torch.Tensor) and ep_prev.is_cuda and ep_prev.is_contiguous()
                                                      and ep_prev.dtype == torch.float64) else _dev_f64(ep_prev, dev, (4, n)))
    out = {} if out is None else out

    def buf(name, shape, dtype=torch.float64):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t

    S, DS, ind = buf("s", (4, n)), buf("ds", (9, n)), buf("ind_p", (n,), torch.uint8)
    lam = buf("lambda", (n,)) if want_lambda else None
    if apply_plastic_strain and ep_in is not None:
        ep_out = ep_in                                   # in-place, as the reference (ep = ep_prev)
        out["ep"] = ep_in
    else:
        ep_out = buf("ep", (4, n)) if want_ep else None
    counts = None
    if want_counts:
        counts = buf("counts", (2,), torch.int64)
        counts.zero_()
    e0p = None
    if e0 is not None:
        e0h = np.ascontiguousarray(np.asarray(e0, dtype=np.float64).ravel())
        if e0h.size != 4:
            raise ValueError("e0 must have 4 entries")
        e0p = e0h.ctypes.data_as(C.POINTER(C.c_double))
    with torch.cuda.device(dev):
        call("fem_dp_return_map", n, _ptr(E), e0p, _ptr(ep_in), _ptr(G), _ptr(K), _ptr(et), _ptr(cc), int(bool(apply_plastic_strain)),
             _ptr(S), _ptr(DS), _ptr(ind), _ptr(lam), _ptr(ep_out), _ptr(counts), _stream())
    return out
