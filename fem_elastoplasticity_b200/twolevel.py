"""Two-level preconditioned CG on the device: M^-1 = D^-1 + P A_c^-1 P^T (kernels in csrc/twolevel.cu).

Point-Jacobi PCG needs O(N) iterations on the reference's near-incompressible footing problem (57 500 at 16M elements);
the coarse bilinear-grid correction makes the count depend on H/h only (~1 800 with a 64 x 64 coarse grid).  The coarse
operator A_c = P^T K P is assembled on the device, inverted once with a dense factorisation (torch.linalg, set-up only) and
applied every iteration by a hand-written dense GEMV.  Any SPD preconditioner yields the same solution as the reference's
dense LU (Plasticity2D_DP/pythonFEM.py:1062-1066), so parity is unaffected; a matrix that changes between Newton
iterations (K_tangent) can keep the coarse operator of K_elast."""
import torch

from ._lib import call
from .plan import _ptr, _stream


class CudaTLOps:
    """The kernels of the C ABI (include/fem_b200.h: fem_coarse_galerkin, fem_tl_*, fem_dense_gemv, fem_pcg_spmv_dot) on
    one plan.  ``grid`` = (x0, y0, hx, hy, ncx, ncy).  The CPU tests of the host logic inject a NumPy statement of the
    same steps (tests/test_distributed_cpu.py)."""

    def __init__(self, plan):
        self.plan = plan
        self.device, self.n_dof, self.n_n, self.coord = plan.device, plan.n_dof, plan.n_n, plan.coord

    def sync(self):
        torch.cuda.synchronize()

    def jacobi(self, k, mask, out):
        self.plan.jacobi(k, mask, out=out)

    def galerkin(self, k, row_mask, col_mask, grid, Ac):
        call("fem_coarse_galerkin", self.plan._h, _ptr(k), _ptr(row_mask), _ptr(col_mask), _ptr(self.coord), *grid, _ptr(Ac), _stream())

    def tl_init(self, rhs, mask, minv, grid, r, rc, scal):
        call("fem_tl_init", self.n_n, _ptr(rhs), None, _ptr(mask), _ptr(minv), _ptr(self.coord), *grid, _ptr(r), _ptr(rc), _ptr(scal), _stream())

    def gemv(self, n, A, x, y, dot):
        call("fem_dense_gemv", n, _ptr(A), _ptr(x), _ptr(y), _ptr(dot), _stream())

    def tl_apply(self, mode, r, minv, mask, grid, zc, p, scal, it):
        call("fem_tl_apply", self.n_n, mode, _ptr(r), _ptr(minv), _ptr(mask), _ptr(self.coord), *grid, _ptr(zc), _ptr(p), _ptr(scal), 0, it, _stream())

    def spmv_dot(self, k, p, q, mask, scal, it):
        call("fem_pcg_spmv_dot", self.plan._h, _ptr(k), _ptr(p), _ptr(q), _ptr(mask), _ptr(scal), it, _stream())

    def tl_update_xr(self, p, q, minv, grid, x, r, rc, scal, it):
        call("fem_tl_update_xr", self.n_n, _ptr(p), _ptr(q), _ptr(minv), _ptr(self.coord), *grid, _ptr(x), _ptr(r), _ptr(rc), _ptr(scal), it, _stream())


class TwoLevelPCG:
    def __init__(self, plan, mask, nc=64, part=None, max_coarse_dofs=12000, free_mask=None, ops=None):
        """``mask``: the unknowns of this rank (free DOFs; free AND owned on a strip partition).  ``free_mask``: all free
        DOFs of the local vectors, ghost rows included (needed on a partition: the Galerkin product couples owned rows to
        ghost columns); ``nc``: coarse cells along the shorter side of the bounding box."""
        self.ops = ops if ops is not None else CudaTLOps(plan)
        self.plan, self.mask, self.part = plan, mask, part
        self.free_mask = free_mask
        if part is not None and part.world > 1 and free_mask is None:
            raise ValueError("TwoLevelPCG on a partition needs free_mask (free DOFs including ghost rows)")
        dev = self.ops.device
        co = self.ops.coord
        lo = torch.stack([co[0].min(), co[1].min()])
        hi = torch.stack([co[0].max(), co[1].max()])
        if part is not None and part.world > 1:
            import torch.distributed as dist
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        (x0, y0), (x1, y1) = lo.tolist(), hi.tolist()
        lx, ly = max(x1 - x0, 1e-300), max(y1 - y0, 1e-300)
        ncx, ncy = (max(1, round(nc * lx / ly)), nc) if lx >= ly else (nc, max(1, round(nc * ly / lx)))
        asked = (ncx, ncy)
        while 2 * (ncx + 1) * (ncy + 1) > max_coarse_dofs:      # keep the dense inverse small (n_c^2 doubles)
            ncx, ncy = max(1, int(ncx * 0.9)), max(1, int(ncy * 0.9))
        self.grid_requested = asked
        if (ncx, ncy) != asked:                                  # not silently: the iteration count grows with H/h
            import warnings
            warnings.warn(f"TwoLevelPCG: coarse grid {asked[0]}x{asked[1]} shrunk to {ncx}x{ncy} to keep the dense coarse inverse at "
                          f"<= {max_coarse_dofs} DOFs (max_coarse_dofs); expect more CG iterations", RuntimeWarning, stacklevel=2)
        self.grid = (float(x0), float(y0), float(lx / ncx), float(ly / ncy), int(ncx), int(ncy))
        self.ncd = 2 * (ncx + 1) * (ncy + 1)
        n = self.ops.n_dof
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        self.r, self.p, self.q, self.x, self.minv = (z(n) for _ in range(5))
        self.rc, self.zc, self.scal = z(self.ncd), z(self.ncd), z(8)
        self.Aci = None
        self.setup_seconds = None
        self.inverse_residual = None

    def _reduce(self, t):
        if self.part is not None:
            self.part.all_reduce(t)

    def setup(self, k_vals):
        """Coarse operator of ``k_vals`` (Galerkin), dense inverse; Jacobi part is refreshed in solve()."""
        import time
        o = self.ops
        o.sync()
        t0 = time.perf_counter()
        Ac = torch.empty((self.ncd, self.ncd), dtype=torch.float64, device=o.device)
        o.galerkin(k_vals, self.mask, self.free_mask, self.grid, Ac)
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
        # self-check of the dense inverse on a few random vectors (set-up only): |A_c A_c^-1 v - v| / |v| on the live DOFs
        v = torch.randn((self.ncd, 4), dtype=torch.float64, device=o.device) * (~dead).to(torch.float64)[:, None]
        self.inverse_residual = float(((Ac @ (self.Aci @ v)) - v).norm() / v.norm())
        if not self.inverse_residual <= 1e-6:
            raise ArithmeticError(f"two-level PCG: the dense inverse of the coarse operator is inaccurate (residual {self.inverse_residual:.2e})")
        del Ac, L
        o.sync()
        self.setup_seconds = time.perf_counter() - t0
        return self

    def _coarse_solve(self, rz_slot):
        """z_c = A_c^-1 r_c (replicated on every rank) and the coarse part r_c'z_c of r'z (added once: by rank 0)."""
        self._reduce(self.rc)
        lead = self.part is None or self.part.rank == 0
        self.ops.gemv(self.ncd, self.Aci, self.rc, self.zc, self.scal[rz_slot:rz_slot + 1] if lead else None)

    def solve(self, k_vals, rhs, rtol=1e-10, maxit=100000, check_every=25, iters=None):
        """Returns (x, iterations, relative residual).  ``iters``: run exactly that many iterations (benchmarks)."""
        if self.Aci is None:
            self.setup(k_vals)
        o, part, g, s = self.ops, self.part, self.grid, self.scal
        o.jacobi(k_vals, self.mask, self.minv)
        self.x.zero_()
        o.tl_init(rhs, self.mask, self.minv, g, self.r, self.rc, s)
        self._coarse_solve(0)                             # z_c, and r_c'z_c into s[0] (rank 0)
        self._reduce(s[0:5])                              # r'z = sum r'D^-1 r + r_c'z_c, r'r, |b|^2
        o.tl_apply(1, self.r, self.minv, self.mask, g, self.zc, self.p, s, 0)
        n_it = iters if iters is not None else maxit
        it, rel = 0, float("inf")
        while it < n_it:
            if part is not None:
                part.halo_exchange(self.p)
            o.spmv_dot(k_vals, self.p, self.q, self.mask, s, it)
            self._reduce(s[3:4])
            o.tl_update_xr(self.p, self.q, self.minv, g, self.x, self.r, self.rc, s, it)
            self._coarse_solve(0 if it & 1 else 2)        # + r_c'z_c into the r'z slot: no separate r'z pass
            self._reduce(s[1:3] if it % 2 == 0 else s[0:2])
            o.tl_apply(2, self.r, self.minv, self.mask, g, self.zc, self.p, s, it)
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
        if iters is None and not rel <= rtol:
            from .distributed import PCGNotConverged
            raise PCGNotConverged("two-level PCG", it, rel, rtol)
        return self.x, it, rel
