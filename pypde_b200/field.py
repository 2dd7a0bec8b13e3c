"""
Field / FieldBC / MultiField with the reference's interface (pypde/field.py:9-465).

`v` (physical) and `vhat` (spectral) are CUDA float64 tensors that stay resident
on the device; assigning a NumPy array uploads it.  Transforms loop over the
axes like the reference (field.py:14-56) but zero padding / truncation for the
3/2-rule is folded into the kernels' n_in / n_out arguments instead of np.pad
copies.
"""
import numpy as np
import torch

from . import _cabi as C
from .bases.spectralbase import MetaBase
from .bases.spectralspace import SpectralSpace, SpectralSpaceBC
from .bases.utils import zero_unpad


class _DeviceArray:
    """Descriptor: attribute is always a CUDA float64 tensor (NumPy is uploaded on assignment)."""

    def __init__(self, name):
        self.name = "_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return getattr(obj, self.name)

    def __set__(self, obj, value):
        setattr(obj, self.name, C.to_dev(value))


class FieldBase:
    v = _DeviceArray("v")
    vhat = _DeviceArray("vhat")

    def forward(self, v=None, undealias_after=None):
        """Full forward transform (field.py:14-34)."""
        host = v is not None and C.is_host(v)
        vhat = self.v if v is None else C.to_dev(v)
        if undealias_after is None:
            undealias_after = self.dealiased_space
        for axis in range(self.ndim):
            vhat = self.forward_fft(vhat, axis=axis)
            if undealias_after:
                vhat = zero_unpad(vhat, self.size_undealiased[axis], axis=axis)
        if undealias_after and not vhat.is_contiguous():
            vhat = vhat.contiguous()
        if v is None:
            self.vhat = vhat
        else:
            return C.give_back(vhat, host)

    def backward(self, vhat=None, dealias_before=None):
        """Full backward transform (field.py:36-56).  Coefficient arrays shorter than the
        space (dealiasing: field.py:48-51) are zero padded inside the kernels."""
        host = vhat is not None and C.is_host(vhat)
        v = self.vhat if vhat is None else C.to_dev(vhat)
        for axis in range(self.ndim):
            v = self.backward_fft(v, axis=axis)
        if vhat is None:
            self.v = v
        else:
            return C.give_back(v, host)

    @property
    def x(self):
        return self.xs[0].x

    @property
    def y(self):
        if self.ndim < 2:
            raise ValueError("Dimension y not defined for ndim<2.")
        return self.xs[1].x

    @staticmethod
    def _cellwidth(x):
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

    # -- checkpoint I/O: same keys as the reference's HDF5 layout (field.py:83-173), stored
    #    as .npz because h5py is not part of this image ---------------------------------
    def write(self, filename="file_0.h5", dict=None, leading_str="flow", add_time=True, grp_name=""):
        filename = self._filename(filename, leading_str, add_time)
        grp = grp_name + "/" if grp_name and grp_name[-1] != "/" else grp_name
        data = _npz_load(filename)
        data[grp + "v"] = self.v.cpu().numpy()
        data[grp + "vhat"] = self.vhat.cpu().numpy()
        data["time"] = np.asarray(self.t)
        data["x"] = self.x
        if self.ndim > 1:
            data["y"] = self.y
        if dict is not None:
            for key in dict:
                data[key] = np.asarray(dict[key])
        print("Write {:s} ...".format(filename))
        np.savez(_npz_name(filename), **data)

    def read(self, filename="file_0.h5", dict=None, leading_str="flow", add_time=True, grp_name=""):
        filename = self._filename(filename, leading_str, add_time)
        grp = grp_name + "/" if grp_name and grp_name[-1] != "/" else grp_name
        print("Read {:s} ...".format(filename))
        data = _npz_load(filename)
        if not data:
            raise FileNotFoundError(
                "%s: no checkpoint found (this build stores the reference's HDF5 layout as %s because h5py is "
                "not part of the image; HDF5 files written by pypde cannot be read here)" % (filename, _npz_name(filename)))
        for key in (grp + "v", grp + "vhat", "time"):
            if key not in data:
                raise KeyError("%s: dataset %r missing from checkpoint" % (_npz_name(filename), key))
        # in place, like the reference (`self.vhat[:] = ...`, field.py:150-152): a checkpoint of another
        # resolution is an error, and tensors bound into launch lists / CUDA graphs stay valid
        for name, arr in (("v", data[grp + "v"]), ("vhat", data[grp + "vhat"])):
            cur = getattr(self, name)
            if tuple(arr.shape) != tuple(cur.shape):
                raise ValueError("%s: checkpoint %s has shape %s, field expects %s"
                                 % (_npz_name(filename), grp + name, tuple(arr.shape), tuple(cur.shape)))
            cur.copy_(C.to_dev(arr))
        self.t = float(data["time"])
        if dict is not None:
            for key in dict:
                dict[key] = data[key][()] if key in data else 0.0

    def _filename(self, filename, leading_str, add_time):
        if filename is None:
            filename = leading_str
            if add_time:
                filename = filename + "_{:07.2f}".format(self.t)
            filename = filename + ".h5"
        return filename


def _npz_name(filename):
    return filename if filename.endswith(".npz") else filename + ".npz"


def _npz_load(filename):
    import os
    name = _npz_name(filename)
    if not os.path.exists(name):
        return {}
    with np.load(name) as f:
        return {k: f[k] for k in f.files}


class Field(SpectralSpace, FieldBase):
    """Field variable in physical (`v`) and spectral (`vhat`) space (field.py:180-356)."""

    def __init__(self, bases):
        if isinstance(bases, MetaBase):
            bases = [bases]
        SpectralSpace.__init__(self, bases)
        self.bases = bases
        dev = C.device()
        self.v = torch.zeros(self.shape_physical, dtype=torch.float64, device=dev)
        self.vhat = torch.zeros(self.shape_spectral, dtype=torch.float64, device=dev)
        self.field_bc = None
        self.t = 0
        self.V = []
        self.T = []
        self.dealiased_space = False
        self.create_dealiased_field(bases)

    def create_dealiased_field(self, bases):
        if all(hasattr(i, "dealias") for i in bases):
            self.dealias = Field([i.dealias for i in bases])
            self.dealias.size_undealiased = [self.xs[i].M for i in range(self.ndim)]
            self.dealias.dealiased_space = True

    def add_field_bc(self, field_bc):
        assert isinstance(field_bc, FieldBC)
        self.field_bc = field_bc

    def make_homogeneous(self):
        import warnings
        if self.field_bc is None:
            warnings.warn("No inhomogeneous field found. Call add_field_bc first!")
        else:
            assert self.v.shape == self.inhomogeneous.shape, "Shape mismatch in make_homogeneous"
        return self.v - self.inhomogeneous

    @property
    def total(self):
        return self.homogeneous + self.inhomogeneous

    @property
    def homogeneous(self):
        return self.v

    @property
    def inhomogeneous(self):
        if self.field_bc is not None:
            return self.field_bc.v
        return 0

    def save(self, transform=True):
        """Append a host copy of the physical field (field.py:325-329)."""
        if transform:
            self.backward()
        self.V.append(self.v.cpu().numpy())
        self.T.append(self.t)

    def dstack(self):
        self.VS = np.rollaxis(np.dstack(self.V).squeeze(), -1)
        self.TS = np.rollaxis(np.dstack(self.T).squeeze(), -1)


class FieldBC(SpectralSpaceBC, FieldBase):
    """Inhomogeneous lifting field from boundary values (field.py:359-414)."""

    def __init__(self, bases, axis):
        SpectralSpaceBC.__init__(self, bases, axis)
        dev = C.device()
        self.v = torch.zeros(self.shape_physical, dtype=torch.float64, device=dev)
        self.vhat = torch.zeros(self.shape_spectral, dtype=torch.float64, device=dev)
        self.dealiased_space = False
        self.t = 0

    def add_bc(self, bc):
        """bc: boundary coefficients along self.axis, physical values along the other axes."""
        expected_shape = list(self.shape_physical)
        expected_shape[self.axis] = self.shape_spectral[self.axis]
        assert tuple(bc.shape) == tuple(expected_shape)
        self.v = self.backward_fft(C.to_dev(bc), axis=self.axis)
        self.forward()


class MultiField:
    """Collection of fields with collective save / time update / I/O (field.py:417-465)."""

    def __init__(self, fields, names):
        self.fields, self.names = [], []
        for f, n in zip(fields, names):
            if not isinstance(f, Field):
                raise ValueError("Must be of type Field.")
            self.fields.append(f)
            self.names.append(n)

    def save(self):
        for f in self.fields:
            f.save()

    def update_time(self, dt):
        for f in self.fields:
            f.t += dt

    def read(self, filename=None, leading_str="", add_time=True, dict={}):
        for f, n in zip(self.fields, self.names):
            f.read(filename=filename, leading_str=leading_str, add_time=add_time, dict=dict, grp_name=n)

    def write(self, filename=None, leading_str="", add_time=True, dict={}):
        for f, n in zip(self.fields, self.names):
            f.backward()
            f.write(filename=filename, leading_str=leading_str, add_time=add_time, dict=dict, grp_name=n)

    def interpolate(self, old_fields, spectral=True):
        from .field_operations import interpolate
        for f, f_old in zip(self.fields, old_fields.fields):
            interpolate(f_old, f, spectral)
