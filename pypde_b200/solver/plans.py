"""
Solver plans with the reference's interface (pypde/solver/plans.py:54-341).

    PlanRHS(A, ndim, axis).solve(b)          b <- A b along axis
    PlanLHS(A, ndim, axis, method, ...)      b <- A^-1 b along axis ("fdma", "twodma",
                                             "poisson", "numpy") or A b ("multiply")

Setup (diagonal extraction, LU of the diagonals) runs on the host exactly like
the reference and is uploaded once; `solve` launches the sm_100a kernels on
device tensors.  NumPy in -> NumPy out.
"""
import numpy as np
import scipy.sparse as sp
import torch

from .. import _cabi as C
from .. import ops


class MetaPlan:
    def __init__(self, A, ndim, axis, **kwargs):
        self.A = A
        self.ndim = ndim
        self.axis = axis
        self.N = A.shape[1]
        self.flags = {"axis": axis}

    def solve(self, b):
        raise NotImplementedError

    def _check_b(self, b):
        assert isinstance(b, (np.ndarray, torch.Tensor))
        assert self.ndim == b.ndim, \
            "Dimensionality mismatch: ndim {:4d} | b.ndim {:4d}. Check ndim.".format(self.ndim, b.ndim)
        assert self.N == b.shape[self.axis], \
            "Shape mismatch. : N {:4d} | b.shape[axis] {:4d}. Check axis.".format(self.N, b.shape[self.axis])


def _diag(A, k):
    return np.asarray(A.diagonal(k) if sp.issparse(A) else np.diag(A, k), dtype=float).copy()


class PlanRHS(MetaPlan):
    """b <- A b.  Banded A (<= 8 diagonals) runs as a fused stencil, anything else as
    a dense fp64 tensor-core contraction."""

    def __init__(self, A, ndim, axis):
        MetaPlan.__init__(self, A=A, ndim=ndim, axis=axis)
        self.flags.update({"method": "multiply"})
        self.band = ops.Band.from_matrix(A)
        self.dense = None
        if self.band is None:
            self.dense = C.upload(A.toarray() if sp.issparse(A) else A)

    def solve(self, b):
        self._check_b(b)
        host = C.is_host(b)
        x = C.to_dev(b)
        if self.band is not None:
            y = ops.banded_mul(self.band, x, axis=self.axis)
        else:
            y = ops.dense_mul(self.dense, x, axis=self.axis)
        return C.give_back(y, host)


class _InPlacePlan(MetaPlan):
    """Shared driver of the in-place solves: NumPy input is solved on a device copy and
    written back into the caller's array (the reference mutates b)."""

    def solve(self, b):
        self._check_b(b)
        if self.ndim > 2:
            raise NotImplementedError("{} supports only ndim<3.".format(type(self).__name__))
        if C.is_host(b):
            x = C.to_dev(b).contiguous()
            self._solve_dev(x)
            b[...] = x.cpu().numpy()
            return b
        if b.dim() == 2 and b.stride(1) != 1:
            x = b.contiguous()
            self._solve_dev(x)
            b.copy_(x)
            return b
        self._solve_dev(b)
        return b


class Plan_twodma(_InPlacePlan):
    """A banded with diagonals at offsets 0, 2 (plans.py:137-181)."""

    def __init__(self, A, ndim, axis):
        MetaPlan.__init__(self, A=A, ndim=ndim, axis=axis)
        self.flags.update({"method": "twodma"})
        self.d, self.u = _diag(A, 0), _diag(A, 2)
        self._t = (C.upload(self.d), C.upload(self.u))

    def _solve_dev(self, x):
        ops.twodma_solve(self._t[0], self._t[1], x, axis=self.axis)


class Plan_fdma(_InPlacePlan):
    """A banded with diagonals at offsets -2, 0, 2, 4 (plans.py:184-251)."""

    def __init__(self, A, ndim, axis):
        MetaPlan.__init__(self, A=A, ndim=ndim, axis=axis)
        self.flags.update({"method": "fdma"})
        l, d, u1, u2 = _diag(A, -2), _diag(A, 0), _diag(A, 2), _diag(A, 4)
        self.FDMA_LU(l, d, u1, u2)
        self.l, self.d, self.u1, self.u2 = l, d, u1, u2
        self._t = tuple(C.upload(t if t.size else np.zeros(1)) for t in (l, d, u1, u2))

    @staticmethod
    def FDMA_LU(ld, d, u1, u2):
        """LU factorisation of the four diagonals, same recurrence order as the
        reference (plans.py:226-234 / fdma.f90:136-142)."""
        n = d.shape[0]
        for i in range(2, n):
            ld[i - 2] = ld[i - 2] / d[i - 2]
            d[i] = d[i] - ld[i - 2] * u1[i - 2]
            if i < n - 2:
                u1[i] = u1[i] - ld[i - 2] * u2[i - 2]

    def _solve_dev(self, x):
        ops.fdma_solve(*self._t, x, axis=self.axis)


class Plan_Poisson(_InPlacePlan):
    """(A + alpha_i C) x_i = b_i for every column i (plans.py:254-314): the core of the
    eigen-diagonalised 2-D Poisson solver.  The per-column LU is factored once on
    the device at construction."""

    def __init__(self, A, alpha, C, ndim, axis, singular=False):
        assert ndim == 2
        MetaPlan.__init__(self, A=A, ndim=ndim, axis=axis)
        self.flags.update({"method": "poisson"})
        self.alpha, self.C, self.singular = alpha, C, singular
        if axis != 0:
            raise NotImplementedError("Plan_Poisson: only axis=0 is on the time-step path")
        n = A.shape[0]
        Ad, Cd = np.zeros((4, n)), np.zeros((4, n))
        for k, off in enumerate((-2, 0, 2, 4)):
            for M, dst in ((A, Ad), (C, Cd)):
                dg = _diag(M, off)
                r0 = max(0, -off)
                dst[k, r0:r0 + dg.size] = dg
        self._Ad, self._Cd = Ad, Cd
        self._plan = ops.PoissonPlan(Ad, Cd, np.asarray(alpha, dtype=float), singular)

    def _check_b(self, b):
        MetaPlan._check_b(self, b)
        if np.asarray(self.alpha).size != b.shape[1]:
            raise ValueError("Size of eigenvalue {:3} array does not match to b {:3}!".format(
                np.asarray(self.alpha).size, b.shape[1]))

    def _solve_dev(self, x):
        self._plan.solve(x)


class Plan_numpy(MetaPlan):
    """General dense solve (plans.py:122-134); test helper, not on the time-step path."""

    def __init__(self, A, ndim, axis):
        MetaPlan.__init__(self, A=A, ndim=ndim, axis=axis)
        self.flags.update({"method": "numpy"})
        self._A = C.upload(A.toarray() if sp.issparse(A) else A)

    def solve(self, b):
        self._check_b(b)
        host = C.is_host(b)
        x = C.to_dev(b)
        if self.axis == 0:
            y = torch.linalg.solve(self._A, x)
        else:
            y = torch.linalg.solve(self._A, x.transpose(0, 1)).transpose(0, 1).contiguous()
        return C.give_back(y, host)


def PlanLHS(A, ndim, axis, method, **kwargs):
    """Plan factory (plans.py:77-119)."""
    all_method = {
        "numpy": Plan_numpy,
        "twodma": Plan_twodma,
        "fdma": Plan_fdma,
        "poisson": Plan_Poisson,
        "multiply": PlanRHS,
    }
    if method not in all_method:
        raise ValueError("Method name {:} not found in: {:}".format(method, all_method.keys()))
    return all_method[method](A=A, ndim=ndim, axis=axis, **kwargs)
