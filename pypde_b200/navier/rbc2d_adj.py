"""
Adjoint-descent iteration towards steady states of the 2-D Rayleigh-Benard model, device path of the
reference's navier/rbc2d_adj.py (NavierStokesAdjoint :15-356).

The iteration is built from the operators of the forward model and runs on the same sm_100a kernels through the
reference-facing API (grad / conv_term / galerkin_to_cheby / cheby_to_galerkin / SolverPlan), fields resident on
the GPU.  One stage (rbc2d_adj.py:321-356):

  1. residual of the forward model, (NS(state) - state) / dt, with the batched IMEX stepper      (:189-197)
  2. adjoint fields = residual smoothed by three non-singular eigen-Poisson solves, / nu, / kappa   (:199-205)
  3. explicit update of U and V with pressure gradient, convective and adjoint-convective terms and
     the residual, mapped back to the Galerkin space                                               (:207-262)
  4. pressure projection with the forward model's singular Poisson plan                              (:289-303)
  5. explicit update of T (convection of the adjoint temperature, residual, buoyancy of VA)          (:264-287)

Parity against the CPU oracle (oracle/pypde_port.py::RBC2DAdjoint, pinned bit for bit to the unmodified
reference by tests/golden/make_golden_adjoint.py) is asserted in tests/test_gpu_adjoint.py.
"""
import numpy as np
import torch

from .. import _cabi as C
from ..bases.spectralbase import Base, dealias_policy
from ..field import Field, MultiField
from ..field_operations import cheby_to_galerkin, conv_term, galerkin_to_cheby
from ..solver.integrator import Integrator
from ..templates.poisson import solverplan_poisson2d
from .rbc2d import NavierStokes
from .rbc2d_base import NavierStokesBase


class NavierStokesAdjoint(NavierStokesBase, Integrator):
    """Same constructor keywords as navier.rbc2d.NavierStokes (case, shape, ra, pr, dt, tsave, dealias, integrator,
    beta, aspect); `NS` is the embedded forward model, `T, U, V` the state, `TA, UA, VA` the adjoint fields."""

    avail_cases = ["rbc", "linear", "zero"]

    def __init__(self, case="rbc", dealias_grid="fft", **kwargs):
        if case not in self.avail_cases:
            raise ValueError("Specified case is not available: ", self.avail_cases)
        self.case = case
        Integrator.__init__(self)
        with dealias_policy(dealias_grid):
            NavierStokesBase.__init__(self, **kwargs)
            self.NS = NavierStokes(case=case, dealias_grid=dealias_grid, **self.CONFIG)
            N0, N1 = self.shape
            side = "CN" if case == "rbc" else "CD"

            def galerkin(kind0):
                return Field([Base(N0, kind0, dealias=3 / 2), Base(N1, "CD", dealias=3 / 2)])

            self.T, self.U, self.V = galerkin(side), galerkin("CD"), galerkin("CD")
            self.TA, self.UA, self.VA = galerkin(side), galerkin("CD"), galerkin("CD")
            self.P = Field([Base(N0, "CN"), Base(N1, "CN")])
            self.pres = Field([Base(N0, "CH"), Base(N1, "CH")])
        self.field = MultiField([self.T, self.U, self.V], ["temp", "ux", "uy"])
        self.setup_solver()
        # lifting temperature of the boundary conditions, in physical space on the product grid
        self.Tbc = self.NS.Tbc          # (its physical values Tbc.v were set by add_bc)
        self.temp_bc = self._physical(self.deriv_field, galerkin_to_cheby(self.Tbc.vhat, self.Tbc))
        self.rhs = torch.zeros(self.shape, dtype=torch.float64, device=C.device())

    # ------------------------------------------------------------------ set-up
    def setup_solver(self):
        ns = self.NS
        self.a, self.b, self.c, self.nstage = ns.a, ns.b, ns.c, ns.nstage
        self.solver_P = ns.solver_P
        self.nabla_U = solverplan_poisson2d(self.U.xs, singular=False, scale=self.scale)
        self.nabla_V = solverplan_poisson2d(self.V.xs, singular=False, scale=self.scale)
        self.nabla_T = solverplan_poisson2d(self.T.xs, singular=False, scale=self.scale)

    def reset_time(self):
        self.time = 0.0
        for f in self.field.fields:
            f.time = 0.0

    def set_temperature(self, amplitude=0.5):
        v = amplitude * np.sin(0.5 * np.pi * self.xx) * np.cos(0.5 * np.pi * self.yy)
        self.T.v = C.to_dev(v)
        self.T.forward()

    def save(self):
        for f in (self.T, self.P, self.U, self.V):
            f.save()

    # ------------------------------------------------------------------ building blocks
    def _physical(self, field, vhat):
        return (field.dealias if self.dealias else field).backward(vhat)

    def _to_spectral(self, v):
        return (self.deriv_field.dealias if self.dealias else self.deriv_field).forward(v)

    def conv(self, field, u, deriv):
        return conv_term(field, u, deriv=deriv, deriv_field=self.deriv_field, dealias=self.dealias, scale=self.scale)

    def _conv_adjoint(self, deriv, ux, uz, temp):
        """sum of the adjoint fields' derivative along `deriv` times the matching state component, with the
        lifting temperature added to the temperature (rbc2d_adj.py:162-187)"""
        acc = self.conv(self.UA, ux, deriv)
        acc += self.conv(self.VA, uz, deriv)
        acc += self.conv(self.TA, temp, deriv)
        acc += self.conv(self.TA, self.temp_bc, deriv)
        return self._to_spectral(acc)

    def conv_term_adj_ux(self, fieldx, fieldz, fieldT, ux, uz, temp, add_bc=None):
        return self._conv_adjoint((1, 0), ux, uz, temp)

    def conv_term_adj_uz(self, fieldx, fieldz, fieldT, ux, uz, temp, add_bc=None):
        return self._conv_adjoint((0, 1), ux, uz, temp)

    def update_NS(self):
        """Residual of the forward model and the smoothed adjoint fields."""
        ns, dt = self.NS, float(self.dt)
        pairs = ((ns.U, self.U), (ns.V, self.V), (ns.T, self.T))
        for model, mine in pairs:
            model.vhat[:] = mine.vhat
        ns.update()
        for model, mine in pairs:
            model.vhat[:] = (model.vhat - mine.vhat) / dt
        for plan, model, adj, coef in ((self.nabla_U, ns.U, self.UA, self.nu), (self.nabla_V, ns.V, self.VA, self.nu),
                                       (self.nabla_T, ns.T, self.TA, self.kappa)):
            smooth = plan.solve_lhs(plan.solve_rhs(galerkin_to_cheby(model.vhat, model)))
            adj.vhat[:] = smooth / float(coef)

    def _momentum(self, stage, deriv, state, adjoint, residual):
        a, b, c = float(self.a[stage]), float(self.b[stage]), float(self.c[stage])
        rhs = self.rhs
        rhs.zero_()
        rhs -= a * self.grad(self.pres, deriv=deriv)
        rhs += b * self.NS.conv_term(adjoint, self.ux, self.uz)
        rhs += b * self._conv_adjoint(deriv, self.ux, self.uz, self.temp)
        if c != 0:
            rhs += c * self.NS.conv_term(adjoint, self.ux_old, self.uz_old)
            rhs += c * self._conv_adjoint(deriv, self.ux_old, self.uz_old, self.temp_old)
        rhs += a * galerkin_to_cheby(residual.vhat, state)
        state.vhat += float(self.dt) * cheby_to_galerkin(rhs, state)

    def update_U(self, stage):
        self._momentum(stage, (1, 0), self.U, self.UA, self.NS.U)

    def update_V(self, stage):
        self._momentum(stage, (0, 1), self.V, self.VA, self.NS.V)

    def update_T(self, stage):
        a, b, c = float(self.a[stage]), float(self.b[stage]), float(self.c[stage])
        rhs = self.rhs
        rhs.zero_()
        rhs += b * self.NS.conv_term(self.TA, self.ux, self.uz)
        if c != 0:
            rhs += c * self.NS.conv_term(self.TA, self.ux_old, self.uz_old)
        rhs += a * galerkin_to_cheby(self.NS.T.vhat, self.T)
        rhs += a * galerkin_to_cheby(self.VA.vhat, self.VA)            # buoyancy
        self.T.vhat += float(self.dt) * cheby_to_galerkin(rhs, self.T)

    def update_P(self, div, singular=True):
        self.P.vhat[:] = self.solver_P.solve_lhs(self.solver_P.solve_rhs(div))
        if singular:
            self.P.vhat[0, 0] = 0

    def update_pres(self, div, stage):
        self.pres.vhat += galerkin_to_cheby(self.P.vhat, self.P) / float(self.dt * self.a[stage])

    def update_velocity(self, p, u, v, fac=1.0):
        u.vhat -= cheby_to_galerkin(self.grad(p, deriv=(1, 0)) * fac, u)
        v.vhat -= cheby_to_galerkin(self.grad(p, deriv=(0, 1)) * fac, v)

    def callback(self):
        self.eval_Nu()
        for label, u, v in (("|div|", self.U, self.V), ("|div residual|", self.NS.U, self.NS.V)):
            print("{:s} = {:4.2e}".format(label, float(torch.linalg.norm(self.NS.divergence_velocity(u, v)))))
        for label, f in ((" |U|", self.NS.U), (" |V|", self.NS.V), (" |T|", self.NS.T)):
            print("{:s} = {:5.2e}".format(label, float(torch.linalg.norm(f.vhat))))

    # ------------------------------------------------------------------ the iteration
    def update(self):
        self.ux_old = self.uz_old = self.temp_old = 0
        for rk in range(self.nstage):
            self.ux = self._physical(self.U, self.U.vhat)
            self.uz = self._physical(self.V, self.V.vhat)
            self.temp = self._physical(self.T, self.T.vhat)
            self.update_NS()
            self.update_U(stage=rk)
            self.update_V(stage=rk)
            div = self.NS.divergence_velocity(self.U, self.V)
            self.update_P(div)
            self.update_pres(div, stage=rk)
            self.update_velocity(self.P, self.U, self.V)
            self.update_T(stage=rk)
            self.ux_old, self.uz_old, self.temp_old = self.ux, self.uz, self.temp
