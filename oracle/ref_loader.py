"""Import the real reference modules from /root/reference (TEST INFRASTRUCTURE ONLY).

Only usable in the build container: /root/reference does not exist on the GPU
box, so nothing marked ``gpu``, ``smoke()`` or ``bench.py`` may call this.
matplotlib is not installed and is imported at module top by the reference
(Plasticity2D_DP/pythonFEM.py:26-29), so empty stand-ins are seeded first
(SURVEY.md Appendix C)."""
import importlib.util
import logging
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("FEM_REFERENCE_ROOT", "/root/reference")
_CACHE = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Plasticity2D_DP", "pythonFEM.py"))


def load(which: str):
    """which in {'plasticity', 'elasticity', 'tsx'} -> the reference module object."""
    sub = {"plasticity": "Plasticity2D_DP", "elasticity": "Elasticity2D", "tsx": "tsx-tunnel"}[which]
    if which in _CACHE:
        return _CACHE[which]
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "matplotlib.collections"):
        sys.modules.setdefault(n, types.ModuleType(n))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location("_fem_ref_" + which, os.path.join(REF_ROOT, sub, "pythonFEM.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    logging.getLogger().setLevel(logging.ERROR)
    _CACHE[which] = mod
    return mod
