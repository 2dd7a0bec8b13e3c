"""
Spectral field operations with the reference's signatures
(pypde/field_operations.py:8-220) on device tensors.
"""
import numpy as np
import torch

from . import _cabi as C
from .bases.spectralbase import Base
from .field import Field, FieldBC


def grad(field, deriv, return_field=False, scale=None):
    """Chebyshev coefficients of the mixed derivative d^deriv[0]_x d^deriv[1]_y of `field`
    (field_operations.py:8-56).  The 1/scale**deriv factor is applied inside the
    derivative kernel (same rounding: one division of the finished coefficient)."""
    assert isinstance(field, (Field, FieldBC))
    if isinstance(deriv, int):
        deriv = (deriv,)
    assert field.ndim == len(deriv)
    dvhat = field.vhat
    for axis in range(field.ndim):
        div = 1.0
        if scale is not None:
            assert len(scale) == field.ndim
            div = scale[axis] ** deriv[axis]
        dvhat = field.derivative(dvhat, deriv[axis], axis=axis, div=div)
    if return_field:
        xs = [b.family if hasattr(b, "family") else b for b in field.xs]
        field_deriv = Field(xs)
        field_deriv.vhat = dvhat
        return field_deriv
    return dvhat


def cheby_to_galerkin(uhat, galerkin_field):
    for axis in range(uhat.ndim):
        if hasattr(galerkin_field.xs[axis], "from_chebyshev"):
            uhat = galerkin_field.xs[axis].from_chebyshev(uhat, axis=axis)
    return uhat


def galerkin_to_cheby(vhat, galerkin_field):
    for axis in range(vhat.ndim):
        if hasattr(galerkin_field.xs[axis], "to_chebyshev"):
            vhat = galerkin_field.xs[axis].to_chebyshev(vhat, axis=axis)
    return vhat


def _default_deriv_field(v_field, dealias):
    d = 3 / 2 if dealias else None
    return Field([Base(v_field.shape[0], "CH", dealias=d), Base(v_field.shape[1], "CH", dealias=d)])


def conv_term(v_field, u, deriv, deriv_field=None, dealias=False, scale=None):
    """u * d(v)/dx_i in (dealiased) physical space (field_operations.py:83-129)."""
    assert isinstance(v_field, Field), "v_field must be instance Field"
    if deriv_field is None:
        deriv_field = _default_deriv_field(v_field, dealias)
    vhat = grad(v_field, deriv, return_field=False, scale=scale)
    space = deriv_field.dealias if dealias else deriv_field
    return space.backward(vhat) * C.to_dev(u)


def convective_term(v_field, ux, uz, deriv_field=None, add_bc=None, dealias=False, scale=None):
    """ux dv/dx + uz dv/dz (+ add_bc), transformed to Chebyshev coefficients
    (field_operations.py:132-169)."""
    if deriv_field is None:
        deriv_field = _default_deriv_field(v_field, dealias)
    conv = conv_term(v_field, ux, (1, 0), deriv_field, dealias, scale=scale)
    conv += conv_term(v_field, uz, (0, 1), deriv_field, dealias, scale=scale)
    if add_bc is not None:
        conv += C.to_dev(add_bc)
    space = deriv_field.dealias if dealias else deriv_field
    return space.forward(conv)


def _w(a, like):
    return C.to_dev(a) if isinstance(like, torch.Tensor) else a


def avg_x(f, dx):
    dx = _w(dx, f)
    return (f * dx[:, None]).sum(0) / dx.sum()


def avg_vol(f, dx, dy):
    dx, dy = _w(dx, f), _w(dy, f)
    favgx = (f * dx[:, None]).sum(0) / dx.sum()
    return (favgx * dy).sum() / dy.sum()


def interpolate(Field_old, Field_new, spectral=True):
    """Pad / truncate one field into another of different resolution
    (field_operations.py:184-220)."""
    F_old = Field_old.vhat if spectral else Field_old.v
    F_new = Field_new.vhat if spectral else Field_new.v
    if F_old.ndim != F_new.ndim:
        raise ValueError("Field must be of same dimension!")
    sl = tuple(slice(0, min(i, j)) for i, j in zip(F_old.shape, F_new.shape))
    F_new.zero_()
    F_new[sl] = F_old[sl]
    if spectral:
        Field_new.backward()
    else:
        Field_new.forward()
