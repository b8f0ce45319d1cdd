"""Two-level preconditioned CG on the device: M^-1 = D^-1 + P A_c^-1 P^T (kernels in csrc/twolevel.cu).

Point-Jacobi PCG needs O(N) iterations on the reference's near-incompressible footing problem (57 500 at 16M elements);
the coarse bilinear-grid correction makes the count depend on H/h only (~1 800 with a 64 x 64 coarse grid).  The coarse
operator A_c = P^T K P is assembled on the device, inverted once with a dense factorisation (torch.linalg, set-up only) and
applied every iteration by a hand-written dense GEMV.  Any SPD preconditioner yields the same solution as the reference's
dense LU (Plasticity2D_DP/pythonFEM.py:1062-1066), so parity is unaffected; a matrix that changes between Newton
iterations (K_tangent) can keep the coarse operator of K_elast."""
import ctypes as C

import torch

from ._lib import call
from .plan import _ptr, _stream


class TwoLevelPCG:
    def __init__(self, plan, mask, nc=64, part=None, max_coarse_dofs=12000, free_mask=None):
        """``mask``: the unknowns of this rank (free DOFs; free AND owned on a strip partition).  ``free_mask``: all free
        DOFs of the local vectors, ghost rows included (needed on a partition: the Galerkin product couples owned rows to
        ghost columns); ``nc``: coarse cells along the shorter side of the bounding box."""
        self.plan, self.mask, self.part = plan, mask, part
        self.free_mask = free_mask
        if part is not None and part.world > 1 and free_mask is None:
            raise ValueError("TwoLevelPCG on a partition needs free_mask (free DOFs including ghost rows)")
        dev = plan.device
        co = plan.coord
        lo = torch.stack([co[0].min(), co[1].min()])
        hi = torch.stack([co[0].max(), co[1].max()])
        if part is not None and part.world > 1:
            import torch.distributed as dist
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        (x0, y0), (x1, y1) = lo.tolist(), hi.tolist()
        lx, ly = max(x1 - x0, 1e-300), max(y1 - y0, 1e-300)
        ncx, ncy = (max(1, round(nc * lx / ly)), nc) if lx >= ly else (nc, max(1, round(nc * ly / lx)))
        while 2 * (ncx + 1) * (ncy + 1) > max_coarse_dofs:      # keep the dense inverse small (n_c^2 doubles)
            ncx, ncy = max(1, int(ncx * 0.9)), max(1, int(ncy * 0.9))
        self.grid = (float(x0), float(y0), float(lx / ncx), float(ly / ncy), int(ncx), int(ncy))
        self.ncd = 2 * (ncx + 1) * (ncy + 1)
        n = plan.n_dof
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        self.r, self.p, self.q, self.x, self.minv = (z(n) for _ in range(5))
        self.rc, self.zc, self.scal = z(self.ncd), z(self.ncd), z(8)
        self.Aci = None
        self.setup_seconds = None

    def _reduce(self, t):
        if self.part is not None:
            self.part.all_reduce(t)

    def setup(self, k_vals):
        """Coarse operator of ``k_vals`` (Galerkin), dense inverse; Jacobi part is refreshed in solve()."""
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        P = self.plan
        Ac = torch.empty((self.ncd, self.ncd), dtype=torch.float64, device=P.device)
        call("fem_coarse_galerkin", P._h, _ptr(k_vals), _ptr(self.mask), _ptr(self.free_mask), _ptr(P.coord), *self.grid, _ptr(Ac), _stream())
        self._reduce(Ac)
        Ac = 0.5 * (Ac + Ac.t())
        d = torch.diagonal(Ac)
        dead = d <= 1e-14 * d.max()                               # coarse nodes without free fine support
        Ac[dead, :] = 0.0
        Ac[:, dead] = 0.0
        Ac[dead, dead] = 1.0
        L = torch.linalg.cholesky(Ac)
        self.Aci = torch.cholesky_inverse(L).contiguous()
        self.Aci[dead, :] = 0.0
        self.Aci[:, dead] = 0.0
        del Ac, L
        torch.cuda.synchronize()
        self.setup_seconds = time.perf_counter() - t0
        return self

    def _coarse_solve(self, rz_slot):
        """z_c = A_c^-1 r_c (replicated on every rank) and the coarse part r_c'z_c of r'z (added once: by rank 0)."""
        self._reduce(self.rc)
        lead = self.part is None or self.part.rank == 0
        call("fem_dense_gemv", self.ncd, _ptr(self.Aci), _ptr(self.rc), _ptr(self.zc),
             _ptr(self.scal[rz_slot:rz_slot + 1]) if lead else None, _stream())

    def solve(self, k_vals, rhs, rtol=1e-10, maxit=100000, check_every=25, iters=None):
        """Returns (x, iterations, relative residual).  ``iters``: run exactly that many iterations (benchmarks)."""
        if self.Aci is None:
            self.setup(k_vals)
        P, part, g, s = self.plan, self.part, self.grid, self.scal
        n_n = P.n_n
        P.jacobi(k_vals, self.mask, out=self.minv)
        self.x.zero_()
        call("fem_tl_init", n_n, _ptr(rhs), None, _ptr(self.mask), _ptr(self.minv), _ptr(P.coord), *g, _ptr(self.r), _ptr(self.rc), _ptr(s),
             _stream())
        self._coarse_solve(0)                             # z_c, and r_c'z_c into s[0] (rank 0)
        self._reduce(s[0:5])                              # r'z = sum r'D^-1 r + r_c'z_c, r'r, |b|^2
        call("fem_tl_apply", n_n, 1, _ptr(self.r), _ptr(self.minv), _ptr(self.mask), _ptr(P.coord), *g, _ptr(self.zc), _ptr(self.p),
             _ptr(s), 0, 0, _stream())
        n_it = iters if iters is not None else maxit
        it, rel = 0, float("inf")
        while it < n_it:
            if part is not None:
                part.halo_exchange(self.p)
            call("fem_pcg_spmv_dot", P._h, _ptr(k_vals), _ptr(self.p), _ptr(self.q), _ptr(self.mask), _ptr(s), it, _stream())
            self._reduce(s[3:4])
            call("fem_tl_update_xr", n_n, _ptr(self.p), _ptr(self.q), _ptr(self.minv), _ptr(P.coord), *g, _ptr(self.x), _ptr(self.r),
                 _ptr(self.rc), _ptr(s), it, _stream())
            self._coarse_solve(0 if it & 1 else 2)        # + r_c'z_c into the r'z slot: no separate r'z pass
            self._reduce(s[1:3] if it % 2 == 0 else s[0:2])
            call("fem_tl_apply", n_n, 2, _ptr(self.r), _ptr(self.minv), _ptr(self.mask), _ptr(P.coord), *g, _ptr(self.zc), _ptr(self.p),
                 _ptr(s), 0, it, _stream())
            it += 1
            if (iters is None and it % check_every == 0) or it == n_it:
                h = s.cpu()
                if not torch.isfinite(h[1]):
                    raise ArithmeticError("two-level PCG breakdown: residual is not finite")
                rel = float((h[1] / h[4]).sqrt()) if h[4] > 0 else 0.0
                if iters is None and rel <= rtol:
                    break
        if part is not None:
            part.halo_exchange(self.x)
        return self.x, it, rel
