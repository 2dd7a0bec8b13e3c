"""
ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the pypde_b200 product path).

ctypes front-end for oracle/fortran_kernels.c, presenting the exact call
signatures of the reference's four f2py modules so the same functions can be
(a) registered in sys.modules to run the UNMODIFIED reference Python
    (oracle/shim.py, only where /root/reference exists), and
(b) used by the self-contained NumPy port in oracle/pypde_port.py.

Reference call sites: pypde/bases/chebyshev.py:124,128 (diff_1d/diff_2d),
pypde/bases/linalg/tdma.py:96-98 (solve_tdma_*), pypde/solver/plans.py:164,171
(solve_twodma_*), :216,223 (solve_fdma_*), :313 (solve_fdma_type2).
"""
import ctypes
import os
import subprocess
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fortran_kernels.c")
_LIB = os.path.join(_HERE, "liboracle_kernels.so")

_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """gcc build of the C restatement. -ffp-contract=off: the f2py builds of the
    reference carry no FMA contraction, so neither may the oracle."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
               _SRC, "-o", _LIB, "-lm"]
        subprocess.check_call(cmd)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _estrides(a):
    return [s // a.itemsize for s in a.strides]


def _f64(a):
    a = np.asarray(a)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    return a


# --- differentiate_cheby -----------------------------------------------------
def diff_1d(c):
    c = _f64(c)
    n = c.shape[0]
    dc = np.empty(n)
    lib().orc_diff_2d(_p(c), ctypes.c_long(_estrides(c)[0]), ctypes.c_long(0), _p(dc), n, 1)
    return dc


def diff_2d(c):
    c = _f64(c)
    n, m = c.shape
    s0, s1 = _estrides(c)
    dc = np.empty((n, m))
    lib().orc_diff_2d(_p(c), ctypes.c_long(s0), ctypes.c_long(s1), _p(dc), n, m)
    return dc


# --- tdma --------------------------------------------------------------------
def solve_tdma_1d(a, b, c, d, k):
    d = _f64(d)
    n = d.shape[0]
    x = np.empty(n)
    a, b, c = _vec(a), _vec(b), _vec(c)
    lib().orc_solve_tdma_2d(_p(a), _p(b), _p(c), _p(d), ctypes.c_long(_estrides(d)[0]),
                            ctypes.c_long(0), int(k), _p(x), n, 1)
    return x


def solve_tdma_2d(a, b, c, d, k):
    d = _f64(d)
    n, m = d.shape
    s0, s1 = _estrides(d)
    x = np.empty((n, m))
    a, b, c = _vec(a), _vec(b), _vec(c)
    lib().orc_solve_tdma_2d(_p(a), _p(b), _p(c), _p(d), ctypes.c_long(s0), ctypes.c_long(s1),
                            int(k), _p(x), n, m)
    return x


# --- fdma (in place on x, like f2py intent(inout)) ------------------------------
def _check_inout(x):
    if not isinstance(x, np.ndarray) or x.dtype != np.float64:
        raise TypeError("intent(inout) argument must be a float64 ndarray")
    if not x.flags.writeable:
        raise ValueError("intent(inout) argument must be writeable")


def solve_fdma_1d(l, d, u1, u2, x):
    _check_inout(x)
    l, d, u1, u2 = _vec(l), _vec(d), _vec(u1), _vec(u2)
    lib().orc_solve_fdma_1d(_p(l), _p(d), _p(u1), _p(u2), _p(x),
                            ctypes.c_long(_estrides(x)[0]), x.shape[0])
    return x


def solve_fdma_2d(l, d, u1, u2, x, axis):
    _check_inout(x)
    n, m = x.shape
    s0, s1 = _estrides(x)
    l, d, u1, u2 = _vec(l), _vec(d), _vec(u1), _vec(u2)
    lib().orc_solve_fdma_2d(_p(l), _p(d), _p(u1), _p(u2), _p(x), ctypes.c_long(s0),
                            ctypes.c_long(s1), int(axis), n, m)
    return x


def solve_fdma_type2(A, C, lam, x, axis, singular):
    _check_inout(x)
    n, m = x.shape
    s0, s1 = _estrides(x)
    A = np.ascontiguousarray(A, dtype=np.float64)
    C = np.ascontiguousarray(C, dtype=np.float64)
    lam = _vec(lam)
    lib().orc_solve_fdma_type2(_p(A), _p(C), _p(lam), _p(x), ctypes.c_long(s0), ctypes.c_long(s1),
                               int(axis), int(bool(singular)), n, m)
    return x


# --- twodma ------------------------------------------------------------------
def solve_twodma_1d(d, u, x):
    _check_inout(x)
    d, u = _vec(d), _vec(u)
    lib().orc_solve_twodma_1d(_p(d), _p(u), _p(x), ctypes.c_long(_estrides(x)[0]), x.shape[0])
    return x


def solve_twodma_2d(d, u, x, axis):
    _check_inout(x)
    n, m = x.shape
    s0, s1 = _estrides(x)
    d, u = _vec(d), _vec(u)
    lib().orc_solve_twodma_2d(_p(d), _p(u), _p(x), ctypes.c_long(s0), ctypes.c_long(s1),
                              int(axis), n, m)
    return x


def as_modules():
    """Four module objects with the f2py module names the reference imports."""
    mods = {}
    for name, fns in (
        ("differentiate_cheby", (diff_1d, diff_2d)),
        ("tdma", (solve_tdma_1d, solve_tdma_2d)),
        ("fdma", (solve_fdma_1d, solve_fdma_2d, solve_fdma_type2)),
        ("twodma", (solve_twodma_1d, solve_twodma_2d)),
    ):
        m = types.ModuleType(name)
        for f in fns:
            setattr(m, f.__name__, f)
        mods[name] = m
    return mods
