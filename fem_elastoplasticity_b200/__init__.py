"""B200-native (sm_100a, FP64) hot path of the vectorised 2D elasto-plastic FEM solvers of
MartinBeseda/FEM-ElastoPlasticity: element stiffness evaluation, CSR assembly of the elastic and
consistent-tangent matrices, Drucker-Prager return mapping, CSR SpMV + PCG inside the Newton loop.

Host side stays Python and keeps the reference's pythonFEM.py function signatures
(``fem_elastoplasticity_b200.pythonFEM``); all arithmetic runs in hand-written CUDA kernels behind
a C ABI (``include/fem_b200.h``, loaded with ctypes).  There is no CPU fallback."""
from ._lib import FemError, NonFiniteJacobian, LIB_PATH  # noqa: F401

__all__ = ["FemPlan", "dp_return_map", "FemError", "NonFiniteJacobian", "pythonFEM", "newton", "meshgen", "distributed"]


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA (CPU-side tooling imports it)
    if name in ("FemPlan", "dp_return_map"):
        from . import plan
        return getattr(plan, name)
    if name in ("pythonFEM", "newton", "meshgen", "distributed"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
