"""
ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the pypde_b200 product path).

Self-contained CPU restatement (NumPy / SciPy + oracle/fortran_kernels.c) of the
reference's Chebyshev spectral-Galerkin hot path, so that the checker exists on
the GPU box where /root/reference does not.  It calls the same third-party
arithmetic the reference calls (scipy.fftpack.dctn, scipy.sparse CSR/CSC
mat-mul, numpy.linalg inv/eig/solve) in the same order, and is pinned
bit-for-bit against the unmodified reference run under oracle/shim.py by
tests/golden/make_golden.py (fixtures in tests/golden/*.npz; the reference has
no golden vectors of its own).

Everything here acts on host float64 NumPy arrays.  `axis` arguments follow the
reference: every 1-D operator acts along axis 0, axis 1 is reached through
swapaxes views (pypde/bases/spectralspace.py:33-60).
"""
import numpy as np
import scipy.sparse as sp
from scipy.fftpack import dctn

from . import kernels as K


# -----------------------------------------------------------------------------
# bases
# -----------------------------------------------------------------------------
def gauss_lobatto(N):
    """pypde/bases/dmsuite.py:209-214 with n = N-1 (ascending -1..1)."""
    n = N - 1
    k = np.linspace(n, 0, n + 1)
    return np.sin(np.pi * (n - 2 * k) / (2 * n))


def _threshold_sparse(A, fmt):
    """pypde/bases/utils.py:5-22 (tosparse with tol=1e-12)."""
    A = np.array(A, dtype=float)
    A[np.abs(A) < 1e-12] = 0
    return sp.csc_matrix(A) if fmt == "csc" else sp.csr_matrix(A)


def pinv_d2(N):
    """Pseudo-inverse of the spectral D2 matrix, pypde/bases/dmsuite.py:358-368,
    passed through pinv.py:86-94 (tosparse(...).toarray() with no thresholding)."""
    d0 = np.zeros(N)
    d0[2:-2] = np.array([-1 / (2 * (i ** 2 - 1)) for i in range(2, N - 2)])
    d1 = np.zeros(N - 2)
    d1[2:-2] = np.array([1 / (4 * i * (i + 1)) for i in range(2, N - 4)])
    d2 = np.zeros(N - 2)
    d2[:] = np.array([1 / (4 * i * (i - 1)) for i in range(2, N)])
    d2[0] *= 2
    return sp.diags([d2, d0, d1], [-2, 0, 2]).toarray()


class Basis:
    """One 1-D function space.  kind: "CH" (pypde/bases/chebyshev.py:22-165),
    "CD"/"CN" (:168-363, :366-458), boundary bases "DC"/"NC" (:533-597)."""

    def __init__(self, N, kind, dealias=None):
        self.N, self.kind = int(N), kind
        self.x = gauss_lobatto(self.N)
        self.sign = np.array([(-1) ** k for k in np.arange(self.N)])
        if kind == "CH":
            self.M = self.N
            self.S = None
        else:
            self.M = 2 if kind in ("DC", "NC") else self.N - 2
            S = np.zeros((self.N, self.M))
            if kind == "CD":        # chebyshev.py:382-392
                for i in range(self.M):
                    S[i, i], S[i + 2, i] = 1, -1
            elif kind == "CN":      # chebyshev.py:426-436
                for i in range(self.M):
                    S[i, i], S[i + 2, i] = 1, -((i / (i + 2)) ** 2)
            elif kind == "DC":      # chebyshev.py:552-556
                S[0, 0], S[1, 0] = 0.5, -0.5
                S[0, 1], S[1, 1] = 0.5, 0.5
            elif kind == "NC":      # chebyshev.py:583-587
                S[0, 0], S[1, 0] = 0.5, -1 / 8
                S[0, 1], S[1, 1] = 0.5, 1 / 8
            else:
                raise ValueError(kind)
            self.S = S
            # spectralbase.py:210-232: S, ST memoised dense, *_sp thresholded CSC.
            # (tosparse zeroes |a|<1e-12 IN PLACE on the memoised dense S as well.)
            self.S_sp = _threshold_sparse(S, "csc")
            S[np.abs(S) < 1e-12] = 0
            self.ST_sp = _threshold_sparse(S.T.copy(), "csc")
            if kind in ("CD", "CN"):
                A = S.T @ S             # chebyshev.py:339-345
                self.l2, self.d, self.u2 = (np.diag(A, i) for i in (-2, 0, 2))
        self.bc = None
        if kind == "CD":
            self.bc = Basis(N, "DC")
        elif kind == "CN":
            self.bc = Basis(N, "NC")
        self.dealias = None
        if dealias is not None:    # spectralbase.py:77-78,94-96
            self.dealias = Basis(int(self.N * dealias), kind)

    @property
    def family(self):
        if self.kind == "CH":
            return self
        if not hasattr(self, "_family"):
            self._family = Basis(self.N, "CH")
        return self._family

    # -- pure Chebyshev transforms, chebyshev.py:67-93,143-149 -------------------
    def _cheb_forward(self, f):
        N = self.N
        c = 0.5 * dctn(f, type=1, axes=(0,)) / (N - 1)
        c = self.sign.reshape((N,) + (1,) * (c.ndim - 1)) * c
        m_inv = sp.diags([1.0, *[2.0] * (N - 2), 1.0], 0)
        return m_inv @ c

    def _cheb_backward(self, c):
        N = self.N
        f = self.sign.reshape((N,) + (1,) * (c.ndim - 1)) * c
        f[[0, -1]] = np.array([2, 2]).reshape((2,) + (1,) * (c.ndim - 1)) * f[[0, -1]]
        return 0.5 * dctn(f, type=1, axes=(0,))

    # -- stencil maps, chebyshev.py:287-337,564-566,595-597 ----------------------
    def to_cheb(self, v):
        if self.kind == "CH":
            return v
        assert v.shape[0] == self.M
        return self.S_sp @ v

    def from_cheb(self, u):
        if self.kind == "CH":
            return u
        if self.kind in ("DC", "NC"):
            return np.linalg.solve(self.S[:2, :2], u[:2])
        assert u.shape[0] == self.N
        rhs = self.ST_sp @ u
        if rhs.ndim == 1:
            return K.solve_tdma_1d(self.l2, self.d, self.u2, rhs, 2)
        return K.solve_tdma_2d(self.l2, self.d, self.u2, rhs, 2)

    # -- full 1-D transforms along axis 0, chebyshev.py:241-285 ------------------
    def forward(self, f):
        if self.kind == "CH":
            return self._cheb_forward(f)
        return self.from_cheb(self.family._cheb_forward(f))

    def backward(self, c):
        if self.kind == "CH":
            return self._cheb_backward(c)
        return self.family._cheb_backward(self.to_cheb(c))

    # -- derivative to Chebyshev coefficients, chebyshev.py:117-131,347-363 ------
    def deriv(self, vhat, order):
        u = self.to_cheb(vhat)
        for _ in range(order):
            u = K.diff_1d(u) if u.ndim == 1 else K.diff_2d(u)
        return u

    # -- setup matrices, chebyshev.py:155-165 -----------------------------------
    def B2(self):
        return pinv_d2(self.N)[2:, :]

    def I2(self):
        return np.eye(self.N)[2:, :]


def _ax(fn, a, axis):
    """spectralspace.py:33-60: axis 1 handled through swapaxes views."""
    if axis == 0:
        return fn(a)
    return np.swapaxes(fn(np.swapaxes(a, axis, 0)), axis, 0)


def _pad0(a, n, axis):
    """bases/utils.py:89-110"""
    p = n - a.shape[axis]
    if p <= 0:
        return a
    w = [(0, 0)] * a.ndim
    w[axis] = (0, p)
    return np.pad(a, pad_width=w, mode="constant", constant_values=0)


def _cut0(a, n, axis):
    """bases/utils.py:113-115"""
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(0, n)
    return a[tuple(sl)]


class Space:
    """pypde/field.py:9-56 + spectralspace.py:4-60 without the stored fields:
    a tuple of bases with the 2-D forward / backward loops."""

    def __init__(self, bases, dealiased=False, size_undealiased=None):
        self.xs = list(bases)
        self.ndim = len(self.xs)
        self.dealiased = dealiased
        self.size_undealiased = size_undealiased
        self.shape_physical = tuple(b.N for b in self.xs)
        self.shape_spectral = tuple(b.M for b in self.xs)
        self.dealias = None
        if not dealiased and all(b.dealias is not None for b in self.xs):
            # field.py:271-276
            self.dealias = Space([b.dealias for b in self.xs], True, [b.M for b in self.xs])

    def forward(self, v):
        vhat = v
        for axis in range(self.ndim):
            vhat = _ax(self.xs[axis].forward, vhat, axis)
            if self.dealiased:
                vhat = _cut0(vhat, self.size_undealiased[axis], axis)
        return vhat

    def backward(self, vhat):
        v = vhat
        for axis in range(self.ndim):
            if self.dealiased:
                v = _pad0(v, self.xs[axis].M, axis)
            v = _ax(self.xs[axis].backward, v, axis)
        return v

    def derivative(self, vhat, order, axis):
        return _ax(lambda a: self.xs[axis].deriv(a, order), vhat, axis)

    def grad(self, vhat, deriv, scale=None):
        """field_operations.py:8-56 (values only)."""
        d = vhat
        for axis in range(self.ndim):
            d = self.derivative(d, deriv[axis], axis)
            if scale is not None:
                d = d / scale[axis] ** deriv[axis]
        return d

    def to_cheb(self, vhat):
        """field_operations.py:71-80"""
        for axis in range(self.ndim):
            if self.xs[axis].kind != "CH":
                vhat = _ax(self.xs[axis].to_cheb, vhat, axis)
        return vhat

    def from_cheb(self, uhat):
        """field_operations.py:59-68"""
        for axis in range(self.ndim):
            if self.xs[axis].kind != "CH":
                uhat = _ax(self.xs[axis].from_cheb, uhat, axis)
        return uhat

    @property
    def x(self):
        return self.xs[0].x

    @property
    def y(self):
        return self.xs[1].x

    @staticmethod
    def _cellwidth(x):
        """field.py:68-80"""
        xm = np.zeros(x.size + 1)
        xm[0], xm[-1] = x[0], x[-1]
        xm[1:-1] = (x[1:] + x[:-1]) / 2.0
        return np.diff(xm)

    @property
    def dx(self):
        return self._cellwidth(self.x)

    @property
    def dy(self):
        return self._cellwidth(self.y)


def bc_space(bases, axis):
    """spectralspace.py:69-99: BC basis along `axis`, pure Chebyshev elsewhere."""
    xs = [b.bc if i == axis else b.family for i, b in enumerate(bases)]
    return Space(xs)


def field_bc(bases, axis, bc):
    """field.py:403-414 (FieldBC.add_bc).  Returns (space, v, vhat)."""
    s = bc_space(bases, axis)
    v = _ax(s.xs[axis].backward, bc, axis)
    return s, v, s.forward(v)


# -----------------------------------------------------------------------------
# solver plans
# -----------------------------------------------------------------------------
def _csr(A):
    """solver/plans.py:68 -> matrix.py:68-71 -> solver/utils.py:5-10 (always CSR)."""
    return sp.csr_matrix(A)


def _dot(Acsr, b, axis):
    """solver/matrix.py:48-53"""
    if axis == 0:
        return Acsr @ b
    return np.swapaxes(Acsr @ np.swapaxes(b, axis, 0), axis, 0)


def fdma_lu(A):
    """solver/plans.py:193-199,226-234"""
    l = np.diag(A, -2).copy()
    d = np.diag(A, 0).copy()
    u1 = np.diag(A, 2).copy()
    u2 = np.diag(A, 4).copy()
    n = d.shape[0]
    for i in range(2, n):
        l[i - 2] = l[i - 2] / d[i - 2]
        d[i] = d[i] - l[i - 2] * u1[i - 2]
        if i < n - 2:
            u1[i] = u1[i] - l[i - 2] * u2[i - 2]
    return l, d, u1, u2


class HelmholtzADI:
    """templates/hholtz.py:42-93"""

    def __init__(self, bases, lam, scale=(1, 1)):
        self.rhs, self.old, self.lu = [], [], []
        for axis, b in enumerate(bases):
            S = b.S_sp
            B = b.family.B2()
            I = b.family.I2()
            A = B @ S - lam * (1.0 / scale[axis] ** 2.0) * I @ S
            self.rhs.append(_csr(B))
            self.old.append(_csr(B @ S))
            self.lu.append(fdma_lu(A))

    def solve_rhs(self, b):
        for axis in (0, 1):
            b = _dot(self.rhs[axis], b, axis)
        return b

    def solve_old(self, b):
        for axis in (0, 1):
            b = _dot(self.old[axis], b, axis)
        return b

    def solve_lhs(self, b):
        for axis in (0, 1):
            K.solve_fdma_2d(*self.lu[axis], b, axis)
        return b


class Helmholtz1D:
    """templates/hholtz.py:4-39"""

    def __init__(self, b, lam):
        S, B, I = b.S_sp, b.family.B2(), b.family.I2()
        A = B @ S - lam * I @ S
        self.B, self.BS, self.lu = _csr(B), _csr(B @ S), fdma_lu(A)

    def solve(self, rhs, uold):
        r = self.B @ rhs
        r += self.BS @ uold
        return K.solve_fdma_1d(*self.lu, r)


class PoissonEig:
    """templates/poisson.py:43-110 (2-D, eigendecomposition along axis 1)."""

    def __init__(self, bases, singular=False, scale=(1, 1)):
        bx, by = bases
        Sx, Bx, Ix = bx.S_sp, bx.family.B2(), bx.family.I2()
        self.Ax = Ix @ Sx * (1.0 / scale[0] ** 2.0)
        self.Cx = Bx @ Sx
        Sy, By, Iy = by.S_sp, by.family.B2(), by.family.I2()
        Ay = Iy @ Sy * (1.0 / scale[1] ** 2.0)
        Cy = By @ Sy
        CyI = np.linalg.inv(Cy)
        w, Q = np.linalg.eig(CyI @ Ay)      # solver/utils.py:13-29
        order = np.argsort(w)[::-1]
        w, Q = w[order], Q[:, order]
        Qi = np.linalg.inv(Q)
        if singular:
            w[0] += 1e-20
        self.wy, self.Qy, self.Hy = w, Q, Qi @ CyI @ By
        self.Bx_csr, self.Hy_csr, self.Qy_csr = _csr(Bx), _csr(self.Hy), _csr(self.Qy)

    def solve_rhs(self, b):
        return _dot(self.Hy_csr, _dot(self.Bx_csr, b, 0), 1)

    def solve_lhs(self, b):
        # poisson.py:106 hard-codes singular=True in the LHS plan
        K.solve_fdma_type2(self.Ax, self.Cx, self.wy, b, 0, True)
        return _dot(self.Qy_csr, b, 1)


class Poisson1D:
    """templates/poisson.py:4-40"""

    def __init__(self, b, singular=False):
        S, B, I = b.S_sp, b.family.B2(), b.family.I2()
        A = I @ S
        if singular:
            assert A[0, 0] == 0
            A[0, 0] += 1e-20
        self.B = _csr(B)
        self.d, self.u = np.diag(A, 0), np.diag(A, 2)

    def solve(self, rhs):
        return K.solve_twodma_1d(self.d, self.u, self.B @ rhs)


# -----------------------------------------------------------------------------
# Rayleigh-Benard stepper (navier/rbc2d.py + navier/rbc2d_base.py)
# -----------------------------------------------------------------------------
class Diffusion2D:
    """The reference's 2-D diffusion example with an inhomogeneous Dirichlet wall
    (diffusion/diff_2d-bc.py:9-105): du/dt = kappa lap(u), bases (CD, CN), u(x=-1, y) = cos(pi y),
    theta-scheme with the ADI Helmholtz template (templates/hholtz.py:42-93)."""

    def __init__(self, shape=(20, 20), bases=("CD", "CN"), kappa=1.0, dt=0.2, beta=0.5):
        self.shape, self.kappa, self.dt, self.beta = tuple(shape), kappa, dt, beta
        self.space = Space([Basis(shape[0], bases[0]), Basis(shape[1], bases[1])])
        self.vhat = np.zeros(self.space.shape_spectral)
        self.time = 0.0
        # diff_2d-bc.py:79-85
        bc = np.zeros((2, shape[1]))
        bc[0, :] = np.cos(np.pi * self.space.y)
        self.sbc, self.bc_v, self.bc_vhat = field_bc(self.space.xs, 0, bc)
        # diff_2d-bc.py:74-77
        self.solver = HelmholtzADI(self.space.xs, lam=dt * kappa * beta)
        # diff_2d-bc.py:87-92
        self.fhat = dt * kappa * self.sbc.grad(self.bc_vhat, (0, 2))

    def update(self):
        """diff_2d-bc.py:94-105"""
        c = self.dt * self.kappa * (1.0 - self.beta)
        rhs = self.fhat.copy()
        rhs += c * self.space.grad(self.vhat, (0, 2))
        rhs += c * self.space.grad(self.vhat, (2, 0))
        rhs = self.solver.solve_rhs(rhs)
        rhs += self.solver.solve_old(self.vhat)
        self.vhat = self.solver.solve_lhs(rhs)
        self.time += self.dt

    def total(self):
        """Physical field including the lifting (diff_2d-bc.py:134-135)."""
        return self.space.backward(self.vhat) + self.bc_v


def transfer_function(TL, TM, TR, x, k=0.01):
    """navier/rbc2d.py:437-446"""
    arr = np.zeros(x.shape)
    L = x[-1] - x[0]
    for i in range(x.size):
        xs = x[i] * 2.0 / L
        if xs < 0:
            arr[i] = -k * xs / (k + xs + 1) * (TL - TM) + TM
        else:
            arr[i] = k * xs / (k - xs + 1) * (TR - TM) + TM
    return arr


class RBC2D:
    """State + IMEX stage loop of navier/rbc2d.py:28-434 on NumPy arrays."""

    def __init__(self, case="rbc", shape=(50, 50), ra=5e3, pr=1.0, dt=0.2, tsave=0.1,
                 dealias=True, integrator="eu", beta=1.0, aspect=1.0):
        if case not in ("rbc", "linear", "zero"):
            raise ValueError("Specified case is not available")
        self.case, self.shape, self.ra, self.pr, self.dt = case, tuple(shape), ra, pr, dt
        self.dealias, self.integrator, self.beta, self.aspect = dealias, integrator, beta, aspect
        self.tsave, self.time = tsave, 0.0
        # rbc2d_base.py:8-13,68-81 (normalize=True)
        self.nu = np.sqrt(pr / (ra / 1.0 ** 3.0))
        self.kappa = np.sqrt(1 / pr / (ra / 1.0 ** 3.0))
        self.scale = (aspect * 0.5, 0.5)
        N0, N1 = self.shape
        self.deriv = Space([Basis(N0, "CH", 3 / 2), Basis(N1, "CH", 3 / 2)])
        self.x = self.deriv.x * self.scale[0]
        self.y = self.deriv.y * self.scale[1]
        self.xx, self.yy = np.meshgrid(self.x, self.y, indexing="ij")
        side = "CN" if case == "rbc" else "CD"
        self.sT = Space([Basis(N0, side, 3 / 2), Basis(N1, "CD", 3 / 2)])
        self.sU = Space([Basis(N0, "CD", 3 / 2), Basis(N1, "CD", 3 / 2)])
        self.sV = Space([Basis(N0, "CD", 3 / 2), Basis(N1, "CD", 3 / 2)])
        self.sP = Space([Basis(N0, "CN"), Basis(N1, "CN")])
        self.That_ = np.zeros(self.sT.shape_spectral)   # T.vhat
        self.Uhat = np.zeros(self.sU.shape_spectral)
        self.Vhat = np.zeros(self.sV.shape_spectral)
        self.Phat = np.zeros(self.sP.shape_spectral)
        self.pres = np.zeros(self.shape)
        self.setup_solver()
        self.set_fieldbc()

    # rbc2d_base.py:86-114 / rbc2d.py:180-211
    def setup_solver(self):
        if self.integrator == "rk3":
            self.nstage = 3
            self.a = np.array([8.0 / 15.0, 2.0 / 15.0, 1.0 / 3.0])
            self.b = np.array([8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0])
            self.c = np.array([0, -17.0 / 60.0, -5.0 / 12.0])
        else:
            self.nstage = 1
            self.a, self.b, self.c = np.array([1.0]), np.array([1.0]), np.array([0])
        self.solver_U, self.solver_V, self.solver_T = [], [], []
        for rk in range(self.nstage):
            lam_nu = self.dt * self.a[rk] * self.beta * self.nu
            lam_ka = self.dt * self.a[rk] * self.beta * self.kappa
            self.solver_U.append(HelmholtzADI(self.sU.xs, lam_nu, self.scale))
            self.solver_V.append(HelmholtzADI(self.sV.xs, lam_nu, self.scale))
            self.solver_T.append(HelmholtzADI(self.sT.xs, lam_ka, self.scale))
        self.solver_P = PoissonEig(self.sP.xs, singular=True, scale=self.scale)

    # rbc2d.py:135-178
    def set_fieldbc(self):
        N0, N1 = self.shape
        if self.case == "zero":
            bc = np.zeros((2, N1))
            bc[0, :] = transfer_function(0.5, 0, -0.5, self.y, k=0.02)
            bc[1, :] = bc[0, :]
            axis = 0
        else:
            bc = np.zeros((N0, 2))
            bc[:, 0], bc[:, 1] = 0.5, -0.5
            axis = 1
        self.sTbc, self.Tbc_v, self.Tbc_vhat = field_bc(self.sT.xs, axis, bc)
        self.dTbcdz2 = self.sTbc.grad(self.Tbc_vhat, (0, 2), self.scale)
        vhat = self.sTbc.grad(self.Tbc_vhat, (0, 1), self.scale)
        self.dTbcdz1 = (self.deriv.dealias if self.dealias else self.deriv).backward(vhat)
        self.Tbc_cheby = self.sTbc.to_cheb(self.Tbc_vhat)

    # rbc2d.py:122-133
    def set_temperature(self, amplitude=0.5, m=1):
        v = amplitude * np.sin(m * np.pi * self.xx) * np.cos(np.pi * self.yy)
        self.That_ = self.sT.forward(v)

    def set_velocity(self, amplitude=0.5, m=1, n=1):
        x = (self.x - self.x[0]) / (self.x[-1] - self.x[0])
        y = (self.y - self.y[0]) / (self.y[-1] - self.y[0])
        xx, yy = np.meshgrid(x, y, indexing="ij")
        self.Uhat = self.sU.forward(-amplitude * np.sin(m * np.pi * xx) * np.cos(n * np.pi * yy))
        self.Vhat = self.sV.forward(amplitude * np.cos(m * np.pi * xx) * np.sin(n * np.pi * yy))

    # field_operations.py:83-169 via rbc2d.py:236-250
    def conv_term(self, space, vhat, ux, uz, add_bc=None):
        dsp = self.deriv.dealias if self.dealias else self.deriv
        conv = dsp.backward(space.grad(vhat, (1, 0), self.scale)) * ux
        conv += dsp.backward(space.grad(vhat, (0, 1), self.scale)) * uz
        if add_bc is not None:
            conv += add_bc
        return dsp.forward(conv)

    def _explicit_diffusion(self, space, vhat, coef, stage):
        """rbc2d.py:268-283 (only when beta != 1)"""
        f = self.dt * self.a[stage] * (1 - self.beta) * coef
        return f * space.grad(vhat, (2, 0), self.scale), f * space.grad(vhat, (0, 2), self.scale)

    def _helmholtz(self, solver, rhs, vhat):
        rhs = solver.solve_rhs(rhs)
        rhs += solver.solve_old(vhat)
        return solver.solve_lhs(rhs)

    def divergence(self):
        """rbc2d.py:225-234"""
        return self.sU.grad(self.Uhat, (1, 0), self.scale) + self.sV.grad(self.Vhat, (0, 1), self.scale)

    def update(self):
        """rbc2d.py:396-434 with update_U/V/P/pres/velocity/T (:252-394, :213-223) inlined."""
        dt, a, b, c = self.dt, self.a, self.b, self.c
        chspace = self.deriv  # pres lives in CH x CH
        ux_old = uz_old = 0
        for rk in range(self.nstage):
            That = self.sT.to_cheb(self.That_)
            That += self.Tbc_cheby
            if self.dealias:
                ux = self.sU.dealias.backward(self.Uhat)
                uz = self.sV.dealias.backward(self.Vhat)
            else:
                ux = self.sU.backward(self.Uhat)
                uz = self.sV.backward(self.Vhat)

            # update_U
            dpdx = chspace.grad(self.pres, (1, 0), self.scale)
            rhs = -dt * a[rk] * dpdx
            rhs -= dt * b[rk] * self.conv_term(self.sU, self.Uhat, ux, uz)
            if c[rk] != 0:
                rhs -= dt * c[rk] * self.conv_term(self.sU, self.Uhat, ux_old, uz_old)
            if self.beta != 1.0:
                e0, e1 = self._explicit_diffusion(self.sU, self.Uhat, self.nu, rk)
                rhs += e0
                rhs += e1
            self.Uhat[:] = self._helmholtz(self.solver_U[rk], rhs, self.Uhat)

            # update_V
            dpdz = chspace.grad(self.pres, (0, 1), self.scale)
            rhs = -dt * a[rk] * dpdz
            rhs -= dt * b[rk] * self.conv_term(self.sV, self.Vhat, ux, uz)
            if c[rk] != 0:
                rhs -= dt * c[rk] * self.conv_term(self.sV, self.Vhat, ux_old, uz_old)
            rhs += dt * a[rk] * That
            if self.beta != 1.0:
                e0, e1 = self._explicit_diffusion(self.sV, self.Vhat, self.nu, rk)
                rhs += e0
                rhs += e1
            self.Vhat[:] = self._helmholtz(self.solver_V[rk], rhs, self.Vhat)

            div = self.divergence()

            # update_P
            r = self.solver_P.solve_rhs(div)
            self.Phat[:] = self.solver_P.solve_lhs(r)
            self.Phat[0, 0] = 0

            # update_pres
            self.pres -= 1.0 * self.nu * div * self.beta
            self.pres += 1.0 / (dt * a[rk]) * self.sP.to_cheb(self.Phat)

            # update_velocity
            dpdx = self.sP.grad(self.Phat, (1, 0), self.scale)
            dpdz = self.sP.grad(self.Phat, (0, 1), self.scale)
            self.Uhat -= self.sU.from_cheb(dpdx * 1.0)
            self.Vhat -= self.sV.from_cheb(dpdz * 1.0)

            # update_T
            rhs = -dt * b[rk] * self.conv_term(self.sT, self.That_, ux, uz, add_bc=uz * self.dTbcdz1)
            if c[rk] != 0:
                rhs -= dt * c[rk] * self.conv_term(self.sT, self.That_, ux_old, uz_old,
                                                   add_bc=uz_old * self.dTbcdz1)
            rhs += dt * a[rk] * self.kappa * self.dTbcdz2
            if self.beta != 1.0:
                e0, e1 = self._explicit_diffusion(self.sT, self.That_, self.kappa, rk)
                rhs += e0
                rhs += e1
            self.That_[:] = self._helmholtz(self.solver_T[rk], rhs, self.That_)

            ux_old, uz_old = ux, uz
        self.time += dt

    def iterate(self, nsteps):
        for _ in range(nsteps):
            self.update()

    # rbc2d_base.py:344-386
    def eval_Nu(self):
        Lz = self.y[-1] - self.y[0]
        field = self.deriv
        T = self.sT.backward(self.That_).copy()
        T += self.Tbc_v.copy()
        That = field.forward(T)
        scale = Lz / 2.0
        dThat = field.derivative(That, 1, axis=1) / scale
        dT = field.backward(dThat)
        dTavg = np.sum(dT * field.dx[:, None], axis=0) / np.sum(field.dx)
        Nu = (-dTavg[0] * Lz + -dTavg[-1] * Lz) / 2.0
        V = self.sV.backward(self.Vhat).copy()
        Nuvol = (T * V / self.kappa - dT) * Lz
        favgx = np.sum(Nuvol * field.dx[:, None], axis=0) / np.sum(field.dx)
        Nuvol = np.sum(favgx * field.dy) / np.sum(field.dy)
        return Nu, Nuvol

    def state(self):
        return {"T": self.That_.copy(), "U": self.Uhat.copy(), "V": self.Vhat.copy(),
                "P": self.Phat.copy(), "pres": self.pres.copy()}


class RBC2DAdjoint:
    """navier/rbc2d_adj.py:15-356 on NumPy arrays (TEST INFRASTRUCTURE, like everything in oracle/):
    adjoint-descent iteration towards steady states.  Every stage (i) advances the forward model one step to
    get the residual (state_new - state) / dt (:189-197), (ii) smooths it with three non-singular Poisson solves
    into the adjoint fields (:199-205), (iii) updates U, V explicitly with the adjoint convective terms
    (:207-262), projects them divergence free (:289-303, :341-349) and (iv) updates T (:264-287)."""

    def __init__(self, **cfg):
        self.NS = RBC2D(**cfg)
        ns = self.NS
        self.dt, self.scale, self.dealias = ns.dt, ns.scale, ns.dealias
        self.nu, self.kappa = ns.nu, ns.kappa
        self.sT, self.sU, self.sV, self.sP, self.deriv = ns.sT, ns.sU, ns.sV, ns.sP, ns.deriv
        self.That_ = np.zeros(self.sT.shape_spectral)
        self.Uhat = np.zeros(self.sU.shape_spectral)
        self.Vhat = np.zeros(self.sV.shape_spectral)
        self.Phat = np.zeros(self.sP.shape_spectral)
        self.pres = np.zeros(ns.shape)
        self.TA = np.zeros(self.sT.shape_spectral)
        self.UA = np.zeros(self.sU.shape_spectral)
        self.VA = np.zeros(self.sV.shape_spectral)
        # rbc2d_adj.py:133-150
        self.a, self.b, self.c, self.nstage = ns.a, ns.b, ns.c, ns.nstage
        self.solver_P = ns.solver_P
        self.nabla_U = PoissonEig(self.sU.xs, singular=False, scale=self.scale)
        self.nabla_V = PoissonEig(self.sV.xs, singular=False, scale=self.scale)
        self.nabla_T = PoissonEig(self.sT.xs, singular=False, scale=self.scale)
        # rbc2d_adj.py:112-117: the lifting temperature in physical space
        dsp = self.deriv.dealias if self.dealias else self.deriv
        self.temp_bc = dsp.backward(ns.sTbc.to_cheb(ns.Tbc_vhat))
        self.time = 0.0

    def _phys(self, space, vhat):
        return (space.dealias if self.dealias else space).backward(vhat)

    def _conv(self, space, vhat, u, deriv):
        """field_operations.py:83-129: u * d(field)/dx_i in physical space"""
        dsp = self.deriv.dealias if self.dealias else self.deriv
        return dsp.backward(space.grad(vhat, deriv, self.scale)) * u

    def _conv_adj(self, deriv, ux, uz, temp):
        """rbc2d_adj.py:162-187"""
        conv = self._conv(self.sU, self.UA, ux, deriv)
        conv += self._conv(self.sV, self.VA, uz, deriv)
        conv += self._conv(self.sT, self.TA, temp, deriv)
        conv += self._conv(self.sT, self.TA, self.temp_bc, deriv)
        return (self.deriv.dealias if self.dealias else self.deriv).forward(conv)

    def _residual(self):
        """rbc2d_adj.py:189-205"""
        ns = self.NS
        ns.Uhat[:], ns.Vhat[:], ns.That_[:] = self.Uhat, self.Vhat, self.That_
        ns.update()
        ns.Uhat[:] = (ns.Uhat - self.Uhat) / self.dt
        ns.Vhat[:] = (ns.Vhat - self.Vhat) / self.dt
        ns.That_[:] = (ns.That_ - self.That_) / self.dt
        for plan, space, res, out, coef in ((self.nabla_U, self.sU, ns.Uhat, self.UA, self.nu),
                                            (self.nabla_V, self.sV, ns.Vhat, self.VA, self.nu),
                                            (self.nabla_T, self.sT, ns.That_, self.TA, self.kappa)):
            out[:] = plan.solve_lhs(plan.solve_rhs(space.to_cheb(res))) / coef

    def update(self):
        """rbc2d_adj.py:321-356"""
        ns, a, b, c = self.NS, self.a, self.b, self.c
        ux_old = uz_old = temp_old = 0
        for rk in range(self.nstage):
            ux, uz, temp = self._phys(self.sU, self.Uhat), self._phys(self.sV, self.Vhat), self._phys(self.sT, self.That_)
            self._residual()
            for deriv, space, adj, res, state in (((1, 0), self.sU, self.UA, ns.Uhat, self.Uhat),
                                                  ((0, 1), self.sV, self.VA, ns.Vhat, self.Vhat)):
                rhs = np.zeros(ns.shape)
                rhs -= a[rk] * self.deriv.grad(self.pres, deriv, self.scale)
                rhs += b[rk] * ns.conv_term(space, adj, ux, uz)
                rhs += b[rk] * self._conv_adj(deriv, ux, uz, temp)
                if c[rk] != 0:
                    rhs += c[rk] * ns.conv_term(space, adj, ux_old, uz_old)
                    rhs += c[rk] * self._conv_adj(deriv, ux_old, uz_old, temp_old)
                rhs += a[rk] * space.to_cheb(res)
                state += self.dt * space.from_cheb(rhs)
            # pressure projection (rbc2d_adj.py:289-303 with the forward model's divergence, rbc2d.py:225-234)
            div = self.sU.grad(self.Uhat, (1, 0), self.scale) + self.sV.grad(self.Vhat, (0, 1), self.scale)
            self.Phat[:] = self.solver_P.solve_lhs(self.solver_P.solve_rhs(div))
            self.Phat[0, 0] = 0
            self.pres += self.sP.to_cheb(self.Phat) / (self.dt * a[rk])
            dpdx = self.sP.grad(self.Phat, (1, 0), self.scale)
            dpdz = self.sP.grad(self.Phat, (0, 1), self.scale)
            self.Uhat -= self.sU.from_cheb(dpdx * 1.0)
            self.Vhat -= self.sV.from_cheb(dpdz * 1.0)
            # temperature (rbc2d_adj.py:264-287)
            rhs = np.zeros(ns.shape)
            rhs += b[rk] * ns.conv_term(self.sT, self.TA, ux, uz)
            if c[rk] != 0:
                rhs += c[rk] * ns.conv_term(self.sT, self.TA, ux_old, uz_old)
            rhs += a[rk] * self.sT.to_cheb(ns.That_)
            rhs += a[rk] * self.sV.to_cheb(self.VA)
            self.That_ += self.dt * self.sT.from_cheb(rhs)
            ux_old, uz_old, temp_old = ux, uz, temp
        self.time += self.dt

    def state(self):
        return {"T": self.That_.copy(), "U": self.Uhat.copy(), "V": self.Vhat.copy(), "P": self.Phat.copy(),
                "pres": self.pres.copy(), "TA": self.TA.copy(), "UA": self.UA.copy(), "VA": self.VA.copy()}
