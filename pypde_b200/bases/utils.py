"""Array helpers with the reference's names (pypde/bases/utils.py), accepting
NumPy arrays or CUDA tensors."""
import numpy as np
import torch
from scipy.sparse import csc_matrix, csr_matrix


def tosparse(A, tol=1e-12, format="csc"):
    """Zero |a| < tol (in place, like the reference) and return a sparse matrix."""
    A[np.abs(A) < tol] = 0
    if format in "csc":
        return csc_matrix(A)
    if format in "csr":
        return csr_matrix(A)


def _bcast(a, b):
    assert a.shape[0] == b.shape[0], "First dimension is different"
    assert b.ndim >= a.ndim, "a has more dimensions than b"
    return a.reshape(a.shape + (1,) * (b.ndim - a.ndim))


def product(a, b):
    """a (1-D) times b along b's first dimension."""
    return _bcast(a, b) * b


def add(a, b):
    """a (1-D) plus b along b's first dimension."""
    return _bcast(a, b) + b


def extract_diag(M, k=(-2, 0, 2)):
    return tuple([np.diag(M, i) for i in k])


def zero_pad(array, target_length, axis=0):
    pad = target_length - array.shape[axis]
    if pad <= 0:
        return array
    if isinstance(array, torch.Tensor):
        shp = list(array.shape)
        shp[axis] = target_length
        out = torch.zeros(shp, dtype=array.dtype, device=array.device)
        out.narrow(axis, 0, array.shape[axis]).copy_(array)
        return out
    npad = [(0, 0)] * array.ndim
    npad[axis] = (0, pad)
    return np.pad(array, pad_width=npad, mode="constant", constant_values=0)


def zero_unpad(array, target_length, axis=0):
    if isinstance(array, torch.Tensor):
        return array.narrow(axis, 0, min(target_length, array.shape[axis]))
    slc = [slice(None)] * array.ndim
    slc[axis] = slice(0, target_length)
    return array[tuple(slc)]
