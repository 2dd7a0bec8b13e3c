"""
2-D diffusion with an inhomogeneous Dirichlet wall on the device path — the counterpart of the
reference's example script diffusion/diff_2d-bc.py:9-105 (BASELINE.json configs[2]).

    du/dt = kappa lap(u) on [-1,1]^2,  u(-1, y) = cos(pi y),  u(+1, y) = 0,  du/dy(x, +-1) = 0

Bases (CD, CN); theta-scheme (beta = implicit weight); the implicit part is the ADI Helmholtz
template (banded 4-diagonal sweeps along both axes).  Same class name, constructor keywords and
attributes as the reference script; the only user-visible difference is that the work arrays are
CUDA tensors (fields are device resident).

Run:  python diffusion/diff_2d_bc.py [N] [steps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import Base, Field, FieldBC, Integrator, grad  # noqa: E402
from pypde_b200.templates.hholtz import solverplan_hholtz2d_adi  # noqa: E402


class Diffusion2d(Integrator):
    CONFIG = {
        "bases": ("CD", "CN"),
        "shape": (20, 20),
        "kappa": 1.0,
        "tsave": 0.01,
        "dt": 0.2,
        "ndim": 2,
        "beta": 0.5,
    }

    def __init__(self, **kwargs):
        Integrator.__init__(self)
        self.__dict__.update(**self.CONFIG)
        self.__dict__.update(**kwargs)
        self.time = 0.0
        self.field = Field([Base(self.shape[0], self.bases[0]), Base(self.shape[1], self.bases[1])])
        self.setup_fieldbc()
        self.solver_from_template()
        self.init_field()
        self.field.save()
        self.rhs = torch.zeros(self.shape, dtype=torch.float64, device=self.field.vhat.device)
        self._fhat_cache = None

    def init_field(self):
        self.field.v[:] = 0
        self.field.forward()
        self.field.backward()

    def solver_from_template(self):
        self.solver = solverplan_hholtz2d_adi(self.field.xs, lam=self.dt * self.kappa * self.beta)

    def setup_fieldbc(self):
        """Lifting field of the wall values (diff_2d-bc.py:79-85)."""
        bc = np.zeros((2, self.shape[1]))
        bc[0, :] = np.cos(np.pi * self.field.y)
        self.fieldbc = FieldBC(self.field.xs, axis=0)
        self.fieldbc.add_bc(bc)

    @property
    def _fhat(self):
        """Forcing from the lifting field, computed once (diff_2d-bc.py:87-92)."""
        if self._fhat_cache is None:
            d2 = grad(self.fieldbc, deriv=(0, 2), return_field=True)
            self._fhat_cache = self.dt * self.kappa * d2.vhat
        return self._fhat_cache

    def update(self):
        """One theta-scheme step (diff_2d-bc.py:94-105), same operation order."""
        c = self.dt * self.kappa * (1.0 - self.beta)
        self.rhs.copy_(self._fhat)
        self.rhs += c * grad(self.field, deriv=(0, 2))
        self.rhs += c * grad(self.field, deriv=(2, 0))
        rhs = self.solver.solve_rhs(self.rhs)
        rhs += self.solver.solve_old(self.field.vhat)
        self.field.vhat = self.solver.solve_lhs(rhs)

    def total(self):
        """Physical field including the lifting (diff_2d-bc.py:134-135)."""
        self.field.backward()
        return self.field.v + self.fieldbc.v


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    D = Diffusion2d(shape=(n, n), dt=0.01, tsave=None, kappa=0.1, beta=0.5)
    for _ in range(steps):
        D.update()
        D.update_time()
    u = D.total()
    print("N=%d steps=%d  max|u|=%.12f  u(center)=%.12f" % (n, steps, float(u.abs().max()), float(u[n // 2, n // 2])))
