"""
Chebyshev function spaces with the reference's names and method signatures
(pypde/bases/chebyshev.py), running on the sm_100a kernels.

Every transform takes an extra keyword `axis` (default 0, the reference's only
mode); `SpectralSpace` uses it instead of swapaxes views, so axis-1 work runs in
row kernels on contiguous data.  NumPy in -> NumPy out, tensor in -> tensor out.
"""
from fractions import Fraction

import numpy as np
import scipy.sparse as sp
import torch

from .. import _cabi as C
from .. import ops
from .dmsuite import gauss_lobatto, pseudoinverse_sparse, pseudoinverse_spectral
from .memoize import memoized
from .spectralbase import MetaBase
from .utils import product

USE_PYFFTW = False  # kept for API compatibility; the transform is the CUDA DCT-I


def _io(fn):
    """NumPy in -> NumPy out; tensors stay on the device."""

    def wrapped(self, a, *args, **kwargs):
        host = C.is_host(a)
        out = fn(self, C.to_dev(a), *args, **kwargs)
        return C.give_back(out, host)

    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


class Chebyshev(MetaBase):
    """T_k on the Gauss-Lobatto grid x_j = -cos(pi j/(N-1)) (chebyshev.py:22-165)."""

    def __init__(self, N, dealias=None):
        MetaBase.__init__(self, N, gauss_lobatto(N - 1), dealias)
        self.id = "CH"
        self.family_id = "CH"

    @property
    def family(self):
        return self

    @property
    def plan(self):
        return ops.DctPlan.get(self.N)

    def get_basis(self, i=0, x=None):
        x = self.x if x is None else x
        return np.cos(i * np.arccos(x))

    def get_basis_derivative(self, i=0, k=0, x=None):
        from numpy.polynomial import chebyshev as npcheb
        x = np.atleast_1d(self.x if x is None else x)
        c = np.zeros(self.N)
        c[i] = 1
        b = npcheb.Chebyshev(c)
        return (b.deriv(k) if k > 0 else b)(x)

    # -- transforms (chebyshev.py:67-93) --------------------------------------------
    @_io
    def forward_fft(self, f, mass=True, axis=0, n_out=None):
        """Physical values -> Chebyshev coefficients: m_k (-1)^k DCT1(f)_k / (2(N-1))."""
        c = ops.dct1(self.plan, ops.FWD, f, axis=axis, n_out=n_out)
        if not mass:
            m = torch.full((c.shape[axis],), 2.0, dtype=torch.float64, device=c.device)
            m[0] = 1.0
            if c.shape[axis] == self.N:
                m[-1] = 1.0
            c = c / m.reshape([-1 if i == axis else 1 for i in range(c.dim())])
        return c

    @_io
    def backward_fft(self, c, axis=0, n_out=None):
        """Chebyshev coefficients (possibly fewer than N: zero padded) -> physical values."""
        return ops.dct1(self.plan, ops.BWD, c, axis=axis, n_out=n_out)

    @_io
    def dctn(self, f, axes=(0,), use_pyfftw=False):
        """Unnormalised DCT-I (scipy.fftpack.dctn(f, type=1, axes=axes))."""
        for ax in axes:
            f = ops.dct1(self.plan, ops.RAW, f, axis=ax)
        return f

    @_io
    def derivative(self, fhat, deriv, out_cheby=True, axis=0, div=1.0):
        """Chebyshev coefficients of the deriv-th derivative (chebyshev.py:117-131)."""
        return ops.cheb_diff(fhat, int(deriv), axis=axis, div=div)

    def derivative_physical(self, f, deriv, method="fft"):
        assert method in ["fft", "spectral"], "only the spectral path is provided"
        return self.backward_fft(self.derivative(self.forward_fft(f), deriv))

    def solve_mass(self, f):
        m = np.array([1.0, *[2.0] * (self.N - 2), 1.0])
        if isinstance(f, torch.Tensor):
            return product(C.upload(m), f)
        return product(m, f)

    # -- setup matrices (host) --------------------------------------------------------
    def B(self, deriv, discardrow=0):
        """Pseudo-inverse of D^deriv with the first rows dropped (chebyshev.py:155-161)."""
        if deriv > 2:
            raise ValueError("deriv>2 not supported")
        if deriv == 0:
            return np.eye(self.N)[discardrow:, :]
        return pseudoinverse_spectral(self.N, deriv)[discardrow:, :]

    def I(self, discardrow=0):
        return np.eye(self.N)[discardrow:, :]

    def B_sp(self, deriv, discardrow=0):
        if deriv == 0:
            return sp.identity(self.N, format="csr")[discardrow:, :]
        return pseudoinverse_sparse(self.N, deriv)[discardrow:, :]

    def I_sp(self, discardrow=0):
        return sp.identity(self.N, format="csr")[discardrow:, :]


class GalerkinChebyshev(MetaBase):
    """Composite bases phi_k = T_k + s_k T_{k+2} (chebyshev.py:168-363)."""

    def __init__(self, N, dealias=None):
        MetaBase.__init__(self, N, gauss_lobatto(N - 1), dealias)
        self._bc = None
        self.family_id = "CH"
        self.family = Chebyshev(self.N)
        self._dev = None

    def slice(self):
        return slice(0, self.N - 2)

    # stencil: sub-diagonal s_k of the N x M matrix S (diagonal is 1)
    def stencil_diag(self):
        raise NotImplementedError

    def _stencil(self):
        s = self.stencil_diag()
        S = np.zeros((self.N, self.M))
        for i in range(self.M):
            S[i, i], S[i + 2, i] = 1, s[i]
        return S

    def stencil(self, transpose=False):
        return self._stencil().T if transpose else self._stencil()

    def stencil_sparse(self):
        s = self.stencil_diag().copy()
        s[np.abs(s) < 1e-12] = 0          # tosparse(tol=1e-12), spectralbase.py:167-170
        S = sp.diags([np.ones(self.M), s], [0, -2], shape=(self.N, self.M), format="csc")
        S.eliminate_zeros()
        return S

    def get_basis(self, i=0, x=None):
        return self.get_basis_derivative(i=i, k=0, x=x)

    def get_basis_derivative(self, i=0, k=0, x=None):
        from numpy.polynomial import chebyshev as npcheb
        x = np.atleast_1d(self.x if x is None else x)
        if i >= self.M:
            raise ValueError("basis not known for i={:4d}".format(i))
        c = np.zeros(self.N)
        c[i], c[i + 2] = 1, self.stencil_diag()[i]
        b = npcheb.Chebyshev(c)
        return (b.deriv(k) if k > 0 else b)(x)

    # -- host tables for (S^T S) v = S^T u  (chebyshev.py:305-345, tdma.f90:82-89) ---
    @memoized
    def _init_stencil_inv(self):
        """Diagonals -2, 0, 2 of S^T S.  The main diagonal 1 + s_k^2 is rounded ONCE
        (what the reference's BLAS `S.T @ S` yields on FMA hardware)."""
        s = self.stencil_diag().copy()
        s[np.abs(s) < 1e-12] = 0
        d = np.array([float(Fraction(1) + Fraction(v) * Fraction(v)) for v in s])
        return s[: self.M - 2].copy(), d, s[: self.M - 2].copy()

    def _tables_host(self):
        """(s, a, den, w): stencil, sub-diagonal of S^T S, Thomas denominators and ratios (tdma.f90:82-89)."""
        if getattr(self, "_host", None) is None:
            l2, d, u2 = self._init_stencil_inv()
            n = self.M
            w = np.zeros(max(n - 2, 1))
            den = np.zeros(n)
            for i in range(n):
                den[i] = d[i] if i < 2 else d[i] - l2[i - 2] * w[i - 2]
                if i < n - 2:
                    w[i] = u2[i] / den[i]
            s = self.stencil_diag().copy()
            s[np.abs(s) < 1e-12] = 0
            self._host = (s, l2, den, w)
        return self._host

    def _tables(self):
        if self._dev is None:
            self._dev = tuple(C.upload(t) for t in self._tables_host())
        return self._dev

    # -- transforms --------------------------------------------------------------------
    @_io
    def to_chebyshev(self, vhat, axis=0, n_out=None):
        """u = S v; fewer than M input rows / more than N output rows are zero padding."""
        assert vhat.shape[axis] <= self.M, "{} {}".format(vhat.shape[axis], self.M)
        s = self._tables()[0]
        return ops.to_cheb(s, vhat, axis=axis, n_out=self.N if n_out is None else n_out)

    @_io
    def from_chebyshev(self, uhat, axis=0):
        assert uhat.shape[axis] == self.N, "Shape mismatch ({:3}) ({:3})".format(uhat.shape[axis], self.N)
        s, a, den, w = self._tables()
        return ops.from_cheb(s, a, den, w, uhat, axis=axis)

    @_io
    def forward_fft(self, f, bc=None, axis=0):
        if bc is not None:
            f = f - C.to_dev(self.eval_inhomogeneous(bc, axis=axis))
        c = ops.dct1(self.family.plan, ops.FWD, f, axis=axis)
        s, a, den, w = self._tables()
        return ops.from_cheb(s, a, den, w, c, axis=axis)

    @_io
    def backward_fft(self, c, bc=None, axis=0):
        s = self._tables()[0]
        u = ops.to_cheb(s, c, axis=axis, n_out=self.N)
        if bc is not None:
            u = u + C.to_dev(self.bc.to_chebyshev(bc, axis=axis))
        return ops.dct1(self.family.plan, ops.BWD, u, axis=axis)

    def eval_inhomogeneous(self, bchat, axis=0):
        return self.bc.backward_fft(bchat, axis=axis)

    @_io
    def derivative(self, vhat, deriv, out_cheby=True, axis=0, div=1.0):
        s = self._tables()[0]
        u = ops.to_cheb(s, vhat, axis=axis, n_out=self.N)
        du = ops.cheb_diff(u, int(deriv), axis=axis, div=div)
        if out_cheby:
            return du
        _, a, den, w = self._tables()
        return ops.from_cheb(s, a, den, w, du, axis=axis)


class ChebDirichlet(GalerkinChebyshev):
    """phi_k = T_k - T_{k+2} (chebyshev.py:366-408)."""

    def __init__(self, N, dealias=None):
        GalerkinChebyshev.__init__(self, N, dealias)
        self.id = "CD"
        self.bc = DirichletC(N)

    @memoized
    def stencil_diag(self):
        return -np.ones(self.M)


class ChebNeumann(GalerkinChebyshev):
    """phi_k = T_k - (k/(k+2))^2 T_{k+2} (chebyshev.py:411-458)."""

    def __init__(self, N, dealias=None):
        GalerkinChebyshev.__init__(self, N, dealias)
        self.id = "CN"
        self.bc = NeumannC(N)

    @memoized
    def stencil_diag(self):
        return np.array([-((i / (i + 2)) ** 2) for i in range(self.M)])


class _BoundaryBasis(GalerkinChebyshev):
    """Two-function spaces carrying inhomogeneous boundary values (chebyshev.py:533-597).
    They appear only in FieldBC (setup time), so the 2-column products are plain
    tensor expressions and the 2 x 2 solve runs on the host like the reference's."""

    coeff = None  # 2 x 2: rows T_0, T_1; columns phi_0, phi_1
    is_bc = True

    def slice(self):
        return slice(0, 2)

    def _stencil(self):
        S = np.zeros((self.N, self.M))
        S[:2, :2] = self.coeff
        return S

    def stencil_sparse(self):
        return sp.csc_matrix(self._stencil())

    @_io
    def to_chebyshev(self, vhat, axis=0, n_out=None):
        assert vhat.shape[axis] == 2
        v = vhat.movedim(axis, 0)
        n = self.N if n_out is None else n_out
        u = torch.zeros((n,) + tuple(v.shape[1:]), dtype=torch.float64, device=v.device)
        S = self.coeff
        u[0] = S[0][0] * v[0] + S[0][1] * v[1]
        u[1] = S[1][0] * v[0] + S[1][1] * v[1]
        return u.movedim(0, axis).contiguous()

    @_io
    def from_chebyshev(self, uhat, axis=0):
        u = uhat.movedim(axis, 0)[:2].cpu().numpy()
        v = np.linalg.solve(np.array(self.coeff, dtype=float), u)
        return C.upload(v).movedim(0, axis).contiguous()

    @_io
    def forward_fft(self, f, bc=None, axis=0):
        c = ops.dct1(self.family.plan, ops.FWD, f, axis=axis)
        return self.from_chebyshev(c, axis=axis)

    @_io
    def backward_fft(self, c, bc=None, axis=0):
        return ops.dct1(self.family.plan, ops.BWD, self.to_chebyshev(c, axis=axis), axis=axis)

    @_io
    def derivative(self, vhat, deriv, out_cheby=True, axis=0, div=1.0):
        du = ops.cheb_diff(self.to_chebyshev(vhat, axis=axis), int(deriv), axis=axis, div=div)
        return du if out_cheby else self.from_chebyshev(du, axis=axis)


class DirichletC(_BoundaryBasis):
    """phi_0 = T_0/2 - T_1/2, phi_1 = T_0/2 + T_1/2"""
    coeff = ((0.5, 0.5), (-0.5, 0.5))

    def __init__(self, N, dealias=None):
        GalerkinChebyshev.__init__(self, N, dealias)
        self.id = "DC"


class NeumannC(_BoundaryBasis):
    """phi_0 = T_0/2 - T_1/8, phi_1 = T_0/2 + T_1/8"""
    coeff = ((0.5, 0.5), (-1 / 8, 1 / 8))

    def __init__(self, N, dealias=None):
        GalerkinChebyshev.__init__(self, N, dealias)
        self.id = "NC"
