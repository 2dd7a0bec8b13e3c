"""
Tensor-level wrappers around the C ABI (one function per entry point of
include/pypde_b200.h).  All arguments are CUDA float64 tensors; outputs are
freshly allocated unless stated otherwise.  `axis` follows the reference: the
operator acts along that axis of a 1-D / 2-D array.
"""
import ctypes

import numpy as np
import torch

from . import _cabi as C


def _shape2(t):
    return (t.shape[0], 1) if t.dim() == 1 else tuple(t.shape)


def _axis_len(t, axis):
    return t.shape[axis]


def _out_like(x, axis, n_out):
    shp = list(x.shape)
    shp[axis] = n_out
    return torch.empty(shp, dtype=torch.float64, device=x.device)


def _batch(x2, axis):
    return x2.shape[1 - axis]


# ------------------------------------------------------------------ DCT-I plans
class DctPlan:
    """Owns a pde_dct_plan_t (constant tables only)."""
    _cache = {}

    def __init__(self, L, algo=0):
        self.L = int(L)
        h = ctypes.c_void_p()
        C.check(C.lib().pde_dct_plan_create(ctypes.byref(h), self.L, int(algo)))
        self.handle = h
        self.algo = int(C.lib().pde_dct_plan_algo(h))

    @classmethod
    def get(cls, L, algo=0):
        key = (int(L), int(algo), torch.cuda.current_device())
        if key not in cls._cache:
            cls._cache[key] = cls(L, algo)
        return cls._cache[key]

    def __del__(self):
        try:
            if self.handle:
                C.lib().pde_dct_plan_destroy(self.handle)
        except Exception:
            pass


RAW, FWD, BWD = 0, 1, 2


def dct1(plan, mode, x, axis=0, n_out=None):
    """y = DCT-I(x) along axis; x may be shorter than plan.L (implicit zero pad)."""
    n_in = x.shape[axis]
    n_out = plan.L if n_out is None else int(n_out)
    x2, ldx = C.mat2d(x)
    y = _out_like(x, axis, n_out)
    y2, ldy = C.mat2d(y)
    C.check(C.lib().pde_dct1(plan.handle, mode, C.p(x2), ldx, n_in, C.p(y2), ldy, n_out,
                             _batch(x2, axis), axis, C.stream()))
    return y


# ------------------------------------------------------------------ stencils
def to_cheb(s, v, axis=0, n_out=None):
    M = v.shape[axis]
    n_out = M + 2 if n_out is None else int(n_out)
    v2, ldv = C.mat2d(v)
    u = _out_like(v, axis, n_out)
    u2, ldu = C.mat2d(u)
    C.check(C.lib().pde_to_cheb(C.p(s), C.p(v2), ldv, M, C.p(u2), ldu, n_out, _batch(v2, axis), axis,
                                C.stream()))
    return u


def from_cheb(s, a, den, w, u, axis=0):
    M = s.numel()
    assert u.shape[axis] == M + 2
    u2, ldu = C.mat2d(u)
    v = _out_like(u, axis, M)
    v2, ldv = C.mat2d(v)
    C.check(C.lib().pde_from_cheb(C.p(s), C.p(a), C.p(den), C.p(w), C.p(u2), ldu, M, C.p(v2), ldv,
                                  _batch(u2, axis), axis, C.stream()))
    return v


def tdma2(a, den, w, d, axis=0):
    n = d.shape[axis]
    d2, ldd = C.mat2d(d)
    x = torch.empty(d.shape, dtype=torch.float64, device=d.device)
    x2, ldx = C.mat2d(x)
    C.check(C.lib().pde_tdma2_solve(C.p(a), C.p(den), C.p(w), C.p(d2), ldd, n, C.p(x2), ldx,
                                    _batch(d2, axis), axis, C.stream()))
    return x


def cheb_diff(c, order, axis=0, div=1.0):
    if order == 0:
        return c if div == 1.0 else c / div
    n = c.shape[axis]
    c2, ldc = C.mat2d(c)
    dc = torch.empty(c.shape, dtype=torch.float64, device=c.device)
    d2, ldd = C.mat2d(dc)
    C.check(C.lib().pde_cheb_diff(C.p(c2), ldc, C.p(d2), ldd, n, _batch(c2, axis), axis, int(order),
                                  float(div), C.stream()))
    return dc


# ------------------------------------------------------------------ banded product
class Band:
    """Device copy of a banded (n_out x n_in) matrix: diags[d, r] = A[r, r + offsets[d]]."""

    def __init__(self, offsets, diags, n_in):
        self.offsets = [int(o) for o in offsets]
        d = np.ascontiguousarray(diags, dtype=np.float64)
        self.ndiag, self.n_out = d.shape
        self.n_in = int(n_in)
        self.host = d                      # host copy (table layouts of the fused passes)
        self.diags = C.upload(d)
        self._off = (ctypes.c_int * self.ndiag)(*self.offsets)

    @staticmethod
    def from_matrix(A, max_diag=8):
        """Extract the non-zero diagonals of a dense / scipy.sparse matrix; None if not banded."""
        import scipy.sparse as sp
        A = sp.csr_matrix(A)
        n_out, n_in = A.shape
        coo = A.tocoo()
        nz = coo.data != 0
        offs = np.unique(coo.col[nz] - coo.row[nz])
        if offs.size == 0:
            offs = np.array([0])
        if offs.size > max_diag:
            return None
        diags = np.zeros((offs.size, n_out))
        for i, o in enumerate(offs):
            dg = A.diagonal(int(o))
            r0 = max(0, -int(o))
            diags[i, r0:r0 + dg.size] = dg
        return Band(offs, diags, n_in)


def banded_mul(band, x, axis=0, out=None, accumulate=False):
    assert x.shape[axis] == band.n_in, (x.shape, axis, band.n_in)
    x2, ldx = C.mat2d(x)
    if out is None:
        assert not accumulate
        out = _out_like(x, axis, band.n_out)
    y2, ldy = C.mat2d(out)
    C.check(C.lib().pde_banded_mul(C.p(band.diags), band._off, band.ndiag, C.p(x2), ldx, band.n_in,
                                   C.p(y2), ldy, band.n_out, _batch(x2, axis), axis, int(accumulate),
                                   C.stream()))
    return out


# ------------------------------------------------------------------ banded solves (in place)
def _inplace2d(x):
    if x.dim() == 1:
        if x.stride(0) != 1:
            raise ValueError("in-place solve needs a contiguous 1-D tensor")
        return x.unsqueeze(1), 1
    if x.stride(1) != 1:
        raise ValueError("in-place solve needs a row-major tensor (stride(1) == 1)")
    return x, x.stride(0) if x.shape[0] > 1 else x.shape[1]


def fdma_solve(l, d, u1, u2, x, axis=0):
    x2, ldx = _inplace2d(x)
    C.check(C.lib().pde_fdma_solve(C.p(l), C.p(d), C.p(u1), C.p(u2), C.p(x2), ldx, x.shape[axis],
                                   _batch(x2, axis), axis, C.stream()))
    return x


def twodma_solve(d, u, x, axis=0):
    x2, ldx = _inplace2d(x)
    C.check(C.lib().pde_twodma_solve(C.p(d), C.p(u), C.p(x2), ldx, x.shape[axis], _batch(x2, axis), axis,
                                     C.stream()))
    return x


class PoissonPlan:
    """Owns a pde_poisson_plan_t: per-column LU of (A + lam_i C), factored once on the device."""

    def __init__(self, Adiag, Cdiag, lam, singular):
        Adiag = np.ascontiguousarray(Adiag, dtype=np.float64)
        Cdiag = np.ascontiguousarray(Cdiag, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        assert Adiag.shape == Cdiag.shape and Adiag.shape[0] == 4
        self.n, self.m = Adiag.shape[1], lam.size
        C.device()
        h = ctypes.c_void_p()
        dp = ctypes.POINTER(ctypes.c_double)
        C.check(C.lib().pde_poisson_plan_create(ctypes.byref(h), Adiag.ctypes.data_as(dp),
                                                Cdiag.ctypes.data_as(dp), lam.ctypes.data_as(dp),
                                                self.n, self.m, int(bool(singular))))
        self.handle = h

    def solve(self, x):
        assert tuple(x.shape) == (self.n, self.m), (x.shape, self.n, self.m)
        x2, ldx = _inplace2d(x)
        C.check(C.lib().pde_poisson_solve(self.handle, C.p(x2), ldx, C.stream()))
        return x

    def __del__(self):
        try:
            if self.handle:
                C.lib().pde_poisson_plan_destroy(self.handle)
        except Exception:
            pass


# ------------------------------------------------------------------ dense contraction
def gemm(A, B, transB=False):
    """A (m x k) @ B, B = (k x n) or, with transB, (n x k)."""
    A2, lda = C.mat2d(A)
    B2, ldb = C.mat2d(B)
    m, k = A2.shape
    n = B2.shape[0] if transB else B2.shape[1]
    assert (B2.shape[1] if transB else B2.shape[0]) == k
    out = torch.empty((m, n), dtype=torch.float64, device=A.device)
    C.check(C.lib().pde_gemm_f64(int(transB), C.p(A2), lda, C.p(B2), ldb, C.p(out), n, m, n, k, C.stream()))
    return out


def dense_mul(Mat, x, axis=0):
    """PlanRHS with a dense matrix: axis 0 -> Mat @ x, axis 1 -> x @ Mat^T."""
    if x.dim() == 1:
        return gemm(Mat, x.unsqueeze(1)).squeeze(1)
    if axis == 0:
        return gemm(Mat, x)
    return gemm(x, Mat, transB=True)


def transpose(x):
    x2, ld = C.mat2d(x)
    n0, n1 = x2.shape
    out = torch.empty((n1, n0), dtype=torch.float64, device=x.device)
    C.check(C.lib().pde_transpose(C.p(x2), ld, C.p(out), n0, n0, n1, C.stream()))
    return out
