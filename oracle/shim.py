"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Import the UNMODIFIED reference (/root/reference, read-only, present only in
the build container — never on the GPU box) by pre-registering in sys.modules
  (i)  MagicMock stubs for matplotlib / mpl_toolkits / h5py (not installed), and
  (ii) the C restatements of the four f2py modules (oracle/kernels.py) under the
       module names the reference imports:
         pypde.bases.fortran.differentiate_cheby   (bases/chebyshev.py:11)
         pypde.bases.linalg.fortran.tdma           (bases/linalg/tdma.py:93)
         pypde.solver.linalg.fortran.fdma          (solver/plans.py:214,221,311)
         pypde.solver.linalg.fortran.twodma        (solver/plans.py:162,169)
The shipped .so files cannot be imported (NumPy-1.x f2py ABI, no libgfortran).

Used only by tests/golden/make_golden.py (fixture generation + validation of the
self-contained port in oracle/pypde_port.py).
"""
import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("PYPDE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pypde"))


def load_reference():
    """Returns (pypde, navier.rbc2d) modules of the unmodified reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "pypde" in sys.modules and getattr(sys.modules["pypde"], "__oracle_shim__", False):
        return sys.modules["pypde"], sys.modules["navier.rbc2d"]

    from . import kernels

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
                 "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "h5py"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)

    mods = kernels.as_modules()
    # the reference's `fortran` directories have no __init__.py: provide package
    # objects carrying the replacement modules as attributes
    for pkg, names in (
        ("pypde.bases.fortran", ("differentiate_cheby",)),
        ("pypde.bases.linalg.fortran", ("tdma",)),
        ("pypde.solver.linalg.fortran", ("fdma", "twodma")),
    ):
        p = types.ModuleType(pkg)
        p.__path__ = []
        for n in names:
            setattr(p, n, mods[n])
            sys.modules[pkg + "." + n] = mods[n]
        sys.modules[pkg] = p

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    pypde = importlib.import_module("pypde")
    pypde.__oracle_shim__ = True
    rbc2d = importlib.import_module("navier.rbc2d")
    return pypde, rbc2d
