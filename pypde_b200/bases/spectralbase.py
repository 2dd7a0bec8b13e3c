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
        self.dealias = self.__class__(int(size), dealias=None)

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
