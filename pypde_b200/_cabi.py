"""
ctypes binding of the C ABI in include/pypde_b200.h (libpypde_b200.so, built
in-tree by pypde_b200/build.py for sm_100a).

PyTorch is used for device memory and streams only: every wrapper takes CUDA
float64 tensors, passes `data_ptr()` + sizes + the current stream through the
C ABI and returns tensors.  There is NO CPU fallback: a missing library or a
non-CUDA tensor is an error.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYPDE_B200_LIB") or os.path.join(_HERE, "_lib", "libpypde_b200.so")

_c_dp = ctypes.c_void_p
_c_long = ctypes.c_long
_c_int = ctypes.c_int

# name -> (restype, argtypes); mirrors include/pypde_b200.h line by line
SIGNATURES = {
    "pde_last_error": (ctypes.c_char_p, []),
    "pde_version": (_c_int, []),
    "pde_device_info": (_c_int, [ctypes.POINTER(_c_int)] * 3),
    "pde_launch_count": (_c_long, []),
    "pde_launch_count_reset": (None, []),
    "pde_dct_plan_create": (_c_int, [ctypes.POINTER(ctypes.c_void_p), _c_int, _c_int]),
    "pde_dct_plan_destroy": (_c_int, [ctypes.c_void_p]),
    "pde_dct_plan_algo": (_c_int, [ctypes.c_void_p]),
    "pde_dct1": (_c_int, [ctypes.c_void_p, _c_int, _c_dp, _c_long, _c_int, _c_dp, _c_long, _c_int, _c_int,
                          _c_int, ctypes.c_void_p]),
    "pde_to_cheb": (_c_int, [_c_dp, _c_dp, _c_long, _c_int, _c_dp, _c_long, _c_int, _c_int, _c_int,
                             ctypes.c_void_p]),
    "pde_from_cheb": (_c_int, [_c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_long, _c_int, _c_dp, _c_long, _c_int,
                               _c_int, ctypes.c_void_p]),
    "pde_tdma2_solve": (_c_int, [_c_dp, _c_dp, _c_dp, _c_dp, _c_long, _c_int, _c_dp, _c_long, _c_int, _c_int,
                                 ctypes.c_void_p]),
    "pde_cheb_diff": (_c_int, [_c_dp, _c_long, _c_dp, _c_long, _c_int, _c_int, _c_int, _c_int,
                               ctypes.c_double, ctypes.c_void_p]),
    "pde_banded_mul": (_c_int, [_c_dp, ctypes.POINTER(_c_int), _c_int, _c_dp, _c_long, _c_int, _c_dp, _c_long,
                                _c_int, _c_int, _c_int, _c_int, ctypes.c_void_p]),
    "pde_fdma_solve": (_c_int, [_c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_long, _c_int, _c_int, _c_int,
                                ctypes.c_void_p]),
    "pde_twodma_solve": (_c_int, [_c_dp, _c_dp, _c_dp, _c_long, _c_int, _c_int, _c_int, ctypes.c_void_p]),
    "pde_poisson_plan_create": (_c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                         _c_int, _c_int, _c_int]),
    "pde_poisson_plan_destroy": (_c_int, [ctypes.c_void_p]),
    "pde_poisson_solve": (_c_int, [ctypes.c_void_p, _c_dp, _c_long, ctypes.c_void_p]),
    "pde_poisson_plan_export": (_c_int, [ctypes.c_void_p, _c_int, ctypes.c_void_p]),
    "pde_pass_run": (_c_int, [_c_int, _c_int, _c_int, _c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pde_pass_width": (_c_int, [_c_int]),
    "pde_ipc_alloc": (_c_int, [ctypes.POINTER(ctypes.c_void_p), _c_long, ctypes.c_void_p]),
    "pde_ipc_open": (_c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "pde_ipc_close": (_c_int, [ctypes.c_void_p]),
    "pde_ipc_free": (_c_int, [ctypes.c_void_p]),
    "pde_peer_barrier": (_c_int, [ctypes.c_void_p, ctypes.c_void_p, _c_int, _c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pde_gemm_f64": (_c_int, [_c_int, _c_dp, _c_long, _c_dp, _c_long, _c_dp, _c_long, _c_int, _c_int, _c_int,
                              ctypes.c_void_p]),
    "pde_transpose": (_c_int, [_c_dp, _c_long, _c_dp, _c_long, _c_int, _c_int, ctypes.c_void_p]),
}

class SweepJob(ctypes.Structure):
    """pde_sweep_job (include/pypde_b200.h)"""
    _fields_ = [("inp", ctypes.c_void_p * 5), ("ldin", ctypes.c_long * 5), ("out", ctypes.c_void_p),
                ("ldout", ctypes.c_long), ("tab", ctypes.c_void_p * 6), ("itab", ctypes.c_void_p),
                ("nseq", ctypes.c_int), ("flag", ctypes.c_int), ("sc", ctypes.c_double)]


class StencilJob(ctypes.Structure):
    """pde_stencil_job"""
    _fields_ = [("s", ctypes.c_void_p), ("v", ctypes.c_void_p), ("ldv", ctypes.c_long), ("M", ctypes.c_int),
                ("u", ctypes.c_void_p), ("ldu", ctypes.c_long), ("n_out", ctypes.c_int), ("batch", ctypes.c_int)]


class BandJob(ctypes.Structure):
    """pde_band_job"""
    _fields_ = [("diags", ctypes.c_void_p), ("ndiag", ctypes.c_int), ("off", ctypes.c_int * 8),
                ("x", ctypes.c_void_p), ("ldx", ctypes.c_long), ("n_in", ctypes.c_int), ("y", ctypes.c_void_p),
                ("ldy", ctypes.c_long), ("n_out", ctypes.c_int), ("batch", ctypes.c_int),
                ("accumulate", ctypes.c_int)]


class LincombJob(ctypes.Structure):
    """pde_lincomb_job"""
    _fields_ = [("nterm", ctypes.c_int), ("x", ctypes.c_void_p * 4), ("ldx", ctypes.c_long * 4),
                ("coef", ctypes.c_double * 4), ("y", ctypes.c_void_p), ("ldy", ctypes.c_long),
                ("n0", ctypes.c_int), ("n1", ctypes.c_int)]


SIGNATURES.update({
    "pde_sweep": (_c_int, [_c_int, _c_int, _c_int, _c_int, ctypes.POINTER(SweepJob), ctypes.c_void_p]),
    "pde_to_cheb_multi": (_c_int, [_c_int, _c_int, ctypes.POINTER(StencilJob), ctypes.c_void_p]),
    "pde_banded_multi": (_c_int, [_c_int, _c_int, ctypes.POINTER(BandJob), ctypes.c_void_p]),
    "pde_lincomb_multi": (_c_int, [_c_int, ctypes.POINTER(LincombJob), ctypes.c_void_p]),
    "pde_slab_repack": (_c_int, [_c_int, _c_dp, _c_dp, _c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_int),
                                 ctypes.c_void_p]),
    "pde_conv_products": (_c_int, [_c_long, ctypes.c_double, ctypes.c_double] + [_c_dp] * 11 + [ctypes.c_void_p]),
    "pde_conv_products_members": (_c_int, [_c_long, _c_int, _c_long, ctypes.c_double, ctypes.c_double] + [_c_dp] * 11
                                  + [ctypes.c_void_p]),
    "pde_dct1_batched": (_c_int, [ctypes.c_void_p, _c_int, _c_int, ctypes.c_void_p, _c_long, _c_int, ctypes.c_void_p,
                                  _c_long, _c_int, _c_int, _c_int, _c_int, ctypes.c_void_p]),
    "pde_gemm_f64_batched": (_c_int, [_c_int, _c_dp, ctypes.c_void_p, _c_long, _c_dp, ctypes.c_void_p, _c_long,
                                      ctypes.c_void_p, _c_long, _c_int, _c_int, _c_int, _c_int, _c_int, ctypes.c_void_p]),
    "pde_dct1_multi": (_c_int, [ctypes.c_void_p, _c_int, _c_int, ctypes.POINTER(ctypes.c_void_p), _c_long, _c_int,
                                ctypes.POINTER(ctypes.c_void_p), _c_long, _c_int, _c_int, _c_int, ctypes.c_void_p]),
})

_lib = None


class PdeError(RuntimeError):
    pass


def lib():
    """Load libpypde_b200.so (fails loudly when the extension was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pypde_b200: CUDA extension %s is missing - run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `python -m pypde_b200.build`). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PdeError("pypde_b200 C ABI error %d: %s" % (rc, lib().pde_last_error().decode()))


def device():
    if not torch.cuda.is_available():
        raise PdeError("pypde_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_dev(a):
    """numpy / tensor -> contiguous float64 CUDA tensor (copy only when needed)."""
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != torch.float64 or not t.is_cuda:
            t = t.to(device=device(), dtype=torch.float64)
        return t
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=device())


def upload(a):
    """Small host table -> device tensor (always a fresh contiguous copy)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device())


def is_host(a):
    return not isinstance(a, torch.Tensor)


def give_back(t, like_host):
    """Return numpy when the caller handed in numpy (drop-in behaviour), else the tensor."""
    return t.cpu().numpy() if like_host else t


def mat2d(t):
    """View a 1-D/2-D tensor as a row-major 2-D matrix: returns (tensor2d, ld)."""
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.dim() != 2:
        raise NotImplementedError("pypde_b200 operators support ndim < 3")
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    elif t.shape[1] == 1 and t.stride(0) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        t = t.contiguous()
        ld = t.shape[1]
    return t, ld


def p(t):
    return ctypes.c_void_p(t.data_ptr())


def launch_count():
    return int(lib().pde_launch_count())


def launch_count_reset():
    lib().pde_launch_count_reset()
