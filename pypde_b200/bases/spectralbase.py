"""
`Base` factory and `MetaBase` (API of pypde/bases/spectralbase.py:12-247).

A basis object is host metadata (N, grid, stencil diagonals as NumPy) plus the
device-side constant tables its kernels need; transforms themselves run in the
sm_100a kernels behind pypde_b200._cabi.  Dense N x M stencil matrices
(`S`, `ST`) are only materialised on request (API compatibility / tests); the
hot path works from the diagonals.
"""
import numpy as np
import scipy.sparse as sp

from .memoize import memoized

SPARSE = True

# Size policy of the dealiased twin space (MetaBase.create_dealiased_base):
#   "reference": int(N * dealias) points, exactly like the reference (spectralbase.py:94-96)
#   "fft":       the next L >= int(N * dealias) with L - 1 = 2^a 3^b 5^c even, so the
#                dealiased DCT-I runs as a shared-memory FFT.  With the 3/2-rule the
#                truncated coefficients of a product are alias-free for ANY grid of at
#                least 3N/2 points, so both policies give the same time step up to
#                rounding (tests/test_oracle_cpu.py::test_dealias_is_alias_free).
DEALIAS_POLICY = "reference"


def fft_friendly_size(n):
    """Smallest L >= n such that L - 1 is even with prime factors 2, 3, 5 only."""
    L = max(int(n), 3)
    while True:
        P = L - 1
        if P % 2 == 0:
            m = P
            for f in (2, 3, 5):
                while m % f == 0:
                    m //= f
            if m == 1:
                return L
        L += 1


class dealias_policy:
    """Context manager: `with dealias_policy("fft"): Field([Base(N, "CD", dealias=3/2), ...])`."""

    def __init__(self, policy):
        assert policy in ("reference", "fft")
        self.policy = policy

    def __enter__(self):
        global DEALIAS_POLICY
        self.saved, DEALIAS_POLICY = DEALIAS_POLICY, self.policy

    def __exit__(self, *exc):
        global DEALIAS_POLICY
        DEALIAS_POLICY = self.saved


def Base(N, key, *args, **kwargs):
    """Initialise a basis from its key ("CH", "CD", "CN", "DC", "NC" or the class names)."""
    return _bases_from_key(key)(N, *args, **kwargs)


def _bases_from_key(key):
    from .chebyshev import Chebyshev, ChebDirichlet, ChebNeumann, DirichletC, NeumannC

    table = {
        "CH": Chebyshev, "Chebyshev": Chebyshev,
        "CD": ChebDirichlet, "ChebDirichlet": ChebDirichlet,
        "CN": ChebNeumann, "ChebNeumann": ChebNeumann,
        "DC": DirichletC, "DirichletC": DirichletC,
        "NC": NeumannC, "NeumannC": NeumannC,
    }
    if key in table:
        return table[key]
    if key in ("FO", "Fourier", "CDN", "ChebDirichletNeumann"):
        raise ValueError("Key {:} is outside the Chebyshev time-step path of pypde_b200.".format(key))
    raise ValueError("Key {:} not available.".format(key))


class MetaBase:
    """Common part of all function spaces: sizes, grid, optional dealiased twin."""

    def __init__(self, N, x, dealias=None):
        self._N = int(N)
        self._x = x
        self.name = self.__class__.__name__
        self.id = None
        if dealias is not None:
            self.create_dealiased_base(self.N * dealias)

    @property
    def x(self):
        return self._x

    @property
    def N(self):
        """Number of grid points in physical space"""
        return self._N

    @property
    def M(self):
        """Number of coefficients without BC"""
        return len(range(*self.slice().indices(self.N)))

    def slice(self):
        return slice(0, self.N)

    def create_dealiased_base(self, size):
        """Twin space of int(size) points for dealiased transforms (spectralbase.py:94-96)."""
        size = int(size)
        if DEALIAS_POLICY == "fft":
            size = fft_friendly_size(size)
        self.dealias = self.__class__(size, dealias=None)

    # -- stencil matrices on request ------------------------------------------------
    @property
    @memoized
    def S(self):
        if hasattr(self, "stencil"):
            return self.stencil()
        return np.eye(self.N)

    @property
    @memoized
    def ST(self):
        if hasattr(self, "stencil"):
            return self.stencil(transpose=True)
        return np.eye(self.N)

    @property
    @memoized
    def S_sp(self):
        if hasattr(self, "stencil_sparse"):
            return self.stencil_sparse()
        return sp.identity(self.N, format="csc")

    @property
    @memoized
    def ST_sp(self):
        return sp.csc_matrix(self.S_sp.T)
