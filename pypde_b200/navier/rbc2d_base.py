"""
Configuration, non-dimensional parameters, time-step coefficients and Nusselt
diagnostics of the Rayleigh-Benard solver (navier/rbc2d_base.py:8-257, :344-386).
The steady-state add-on (:260-341) is provided as a host-side SciPy Newton-Krylov driver around the device
update(); the stability add-on (:389-452, dense eigenproblems of pypde/stability) is outside the time-step path.
"""
import numpy as np
import torch

from .. import _cabi as C
from ..bases.spectralbase import Base
from ..field import Field
from ..field_operations import avg_vol, avg_x, grad


def nu(Ra, Pr, L):
    return np.sqrt(Pr / (Ra / L ** 3.0))


def kappa(Ra, Pr, L):
    return np.sqrt(1 / Pr / (Ra / L ** 3.0))


def Ra(nu, kappa, L):
    return 1 / (nu * kappa) * L ** 3.0


def Pr(nu, kappa):
    return nu / kappa


class NavierStokesBase:
    def __init__(self, **kwargs):
        if "Ra" in kwargs or "Pr" in kwargs:
            raise ValueError("Use small ra/pr!")
        self.CONFIG = {
            "shape": (50, 50),
            "ra": 5e3,
            "pr": 1.0,
            "dt": 0.2,
            "ndim": 2,
            "tsave": 0.1,
            "dealias": True,
            "integrator": "eu",
            "beta": 1.0,
            "aspect": 1.0,
        }
        self.CONFIG.update(**kwargs)
        self.__dict__.update(**self.CONFIG)
        self.normalize = True
        self.set_nu_kappa()
        # space for derivatives / nonlinear products (Chebyshev x Chebyshev, 3/2 twin)
        self.deriv_field = Field([
            Base(self.shape[0], "CH", dealias=3 / 2),
            Base(self.shape[1], "CH", dealias=3 / 2),
        ])
        self.x = self.deriv_field.x * self.scale[0]
        self.y = self.deriv_field.y * self.scale[1]
        self.xx, self.yy = np.meshgrid(self.x, self.y, indexing="ij")

    def set_nu_kappa(self, normalize=None):
        if normalize is None:
            normalize = self.normalize
        if normalize:
            self.nu = nu(self.ra, self.pr, L=1.0)
            self.kappa = kappa(self.ra, self.pr, L=1.0)
            self.scale = (self.aspect * 0.5, 0.5)
        else:
            self.nu = nu(self.ra, self.pr, L=2.0)
            self.kappa = kappa(self.ra, self.pr, L=2.0)
            self.scale = (self.aspect * 1.0, 1.0)

    def grad(self, field, deriv, return_field=False):
        return grad(field, deriv=deriv, return_field=return_field, scale=self.scale)

    def set_timestep_coefficients_rk3(self):
        """(1 - a_k L) phi_k = phi_k + b_k N_k + c_k N_{k-1}, diffusion implicit."""
        self.nstage = 3
        self.a = np.array([8.0 / 15.0, 2.0 / 15.0, 1.0 / 3.0])
        self.b = np.array([8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0])
        self.c = np.array([0, -17.0 / 60.0, -5.0 / 12.0])

    def set_timestep_coefficients_euler(self):
        self.nstage = 1
        self.a = np.array([1.0])
        self.b = np.array([1.0])
        self.c = np.array([0])

    def io_config(self):
        print("----------------------------")
        print("Input Parameter:")
        for k, v in self.CONFIG.items():
            print(k, ":", v)
        print("----------------------------")

    def callback(self):
        self.eval_Nu()
        print("|div| = {:4.2e}".format(float(torch.linalg.norm(self.divergence_velocity(self.U, self.V)))))

    def eval_Nu(self):
        Lz = self.y[-1] - self.y[0]
        Nuz = eval_Nu(self.T, self.deriv_field, Tbc=self.Tbc, Lz=Lz)
        Nuv = eval_Nuvol(self.T, self.V, self.kappa, self.deriv_field, Tbc=self.Tbc, Lz=Lz)
        return Nuz, Nuv

    def interpolate(self, NS_old, spectral=True):
        self.field.interpolate(NS_old.field)

    def write(self, filename=None, leading_str="", add_time=True):
        dict = {"nu": self.nu, "kappa": self.kappa,
                "ra": Ra(self.nu, self.kappa, L=self.y[-1] - self.y[0]), "pr": Pr(self.nu, self.kappa)}
        self.field.write(filename=filename, leading_str=leading_str, add_time=add_time, dict=dict)

    def read(self, filename=None, leading_str="", add_time=True):
        dict = {"ra": self.ra, "pr": self.pr}
        self.field.read(filename=filename, leading_str=leading_str, add_time=add_time, dict=dict)
        self.time = self.field.fields[0].t
        self.set_nu_kappa()
        self.CONFIG.update(dict)
        self.__dict__.update(**self.CONFIG)
        self.setup_solver()

    def write_from_Ra(self, folder=""):
        if folder and folder[-1] != "/":
            folder = folder + "/"
        self.write(filename=folder + self.fname_from_Ra(self.ra))

    def read_from_Ra(self, folder=""):
        if folder and folder[-1] != "/":
            folder = folder + "/"
        self.read(filename=folder + self.fname_from_Ra(self.ra))

    @staticmethod
    def fname_from_Ra(Ra):
        return "Flow_Ra{:3.3e}.h5".format(Ra)

    def save(self):
        self.field.save()



class NavierStokesSteadyState:
    """Steady states by Newton-Krylov (rbc2d_base.py:260-341): SciPy's `optimize.root(method="krylov")` runs on
    the host and drives the device time step.  The unknown is the flat vector [T, U, V] of Galerkin
    coefficients; one residual evaluation uploads it into the fields, advances them (one step, or `dt` time
    units) on the GPU and returns (new - old) / dt."""

    def solve_steady_state(self, X0=None, dt=None, maxiter=300, disp=True, tol=1e-8, jac_options=None):
        from scipy import optimize

        if jac_options is None:
            jac_options = {"inner_maxiter": 30}
        if disp:
            print("\nSolve steady state ...\n")
        options = {"maxiter": maxiter, "disp": disp, "fatol": tol, "jac_options": jac_options}
        if X0 is None:
            X0 = self.vectorify()
        return optimize.root(self.steady_fun, X0, args=(self, dt), method="krylov", options=options)

    def _state_fields(self):
        return (self.T, self.U, self.V)

    def flatten(self):
        return tuple(f.vhat.detach().cpu().numpy().ravel().copy() for f in self._state_fields())

    def vectorify(self):
        return np.concatenate(self.flatten())

    def get_masks(self):
        masks, pos = [], 0
        for f in self._state_fields():
            n = f.vhat.numel()
            masks.append(slice(pos, pos + n))
            pos += n
        return tuple(masks)

    def reshape(self, X):
        return tuple(np.array(X[m]).reshape(tuple(f.vhat.shape)) for m, f in zip(self.get_masks(), self._state_fields()))

    def steady_fun(self, X, NS, dt):
        """X: flat [T, U, V] -> residual (NS(X) - X) / dt, flat, on the host."""
        for f, part in zip(NS._state_fields(), NS.reshape(X)):
            f.vhat[:] = torch.as_tensor(part, dtype=f.vhat.dtype, device=f.vhat.device)
        if dt is None:
            dt = NS.dt
            NS.update()
        else:
            NS.reset_time()
            NS.iterate(dt, callback=False)
            NS.reset_time()
        return (NS.vectorify() - X) / dt


def _plate_gradient(T, field, Lz, Tbc):
    """dT/dz in physical space through the derivative space (rbc2d_base.py:348-356)."""
    T.backward()
    Tv = T.v.clone()
    if Tbc is not None:
        Tv += Tbc.v
    That = field.forward(Tv)
    dThat = field.derivative(That, 1, axis=1) / (Lz / 2.0)
    return Tv, field.backward(dThat)


def eval_Nu(T, field, Lz=1.0, Tbc=None):
    """Heat flux at the plates."""
    _, dT = _plate_gradient(T, field, Lz, Tbc)
    dTavg = avg_x(dT, field.dx).cpu().numpy()
    Nu_bot = -dTavg[0] * Lz
    Nu_top = -dTavg[-1] * Lz
    print("Nubot: {:10.6e}".format(Nu_bot))
    print("Nutop: {:10.6e}".format(Nu_top))
    return (Nu_bot + Nu_top) / 2.0


def eval_Nuvol(T, V, kappa, field, Lz=1.0, Tbc=None):
    """Heat flux through the box (volume average)."""
    V.backward()
    Tv, dT = _plate_gradient(T, field, Lz, Tbc)
    Nuvol = (Tv * V.v / kappa - dT) * Lz
    Nuvol = float(avg_vol(Nuvol, field.dx, field.dy))
    print("Nuvol: {:10.6e}".format(Nuvol))
    return Nuvol
