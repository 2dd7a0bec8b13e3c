"""Multi-dimensional spaces (API of pypde/bases/spectralspace.py:4-99).  Per-axis
dispatch passes `axis` to the kernels instead of building swapaxes views."""
import numpy as np

from .spectralbase import MetaBase


class SpectralSpace:
    def __init__(self, bases):
        if isinstance(bases, MetaBase):
            bases = [bases]
        assert all(isinstance(i, MetaBase) for i in bases)
        self._set_bases(bases)
        self.ndim = len(self.shape_physical)
        self.shape = self.shape_physical

    def _set_bases(self, bases):
        self.xs = list(bases)
        self.shape_physical = tuple(b.N for b in self.xs)
        self.shape_spectral = tuple(b.M for b in self.xs)

    def forward_fft(self, v, axis):
        assert isinstance(axis, int)
        return self.xs[axis].forward_fft(v, axis=axis)

    def backward_fft(self, vhat, axis):
        assert isinstance(axis, int)
        return self.xs[axis].backward_fft(vhat, axis=axis)

    def derivative(self, vhat, deriv, axis, out_cheby=True, div=1.0):
        return self.xs[axis].derivative(vhat, deriv, out_cheby, axis=axis, div=div)


class SpectralSpaceBC(SpectralSpace):
    """Boundary-condition basis along `axis`, pure Chebyshev along the others
    (spectralspace.py:69-99)."""

    def __init__(self, bases, axis):
        SpectralSpace.__init__(self, bases)
        self.axis = axis
        xs = []
        for i, b in enumerate(self.xs):
            if i == axis:
                xs.append(b.bc if getattr(b, "bc", None) is not None else b)
            else:
                xs.append(b.family if hasattr(b, "family") else b)
        SpectralSpace.__init__(self, xs)
