"""
2-D Rayleigh-Benard Navier-Stokes solver with the reference's interface
(navier/rbc2d.py:28-434): Python owns the IMEX (Euler / RK3) stage loop, every
field is a device-resident float64 tensor, every operator is an sm_100a kernel
behind the C ABI.

    NS = NavierStokes(case="rbc", shape=(64, 64), ra=1e5, pr=1.0, dt=0.01, tsave=None,
                      dealias=True, integrator="rk3", beta=1.0, aspect=1.0)
    NS.set_velocity(m=1, n=1, amplitude=0.2); NS.set_temperature(amplitude=0.2)
    NS.iterate(1.0); NS.eval_Nu()
"""
import time

import numpy as np
import torch

from .. import _cabi as C
from ..bases.spectralbase import Base, dealias_policy
from ..field import Field, FieldBC, MultiField
from ..field_operations import cheby_to_galerkin, convective_term, galerkin_to_cheby
from ..solver.integrator import Integrator
from .rbc2d_base import NavierStokesBase, NavierStokesSteadyState

# wall-clock accumulators of the reference (navier/rbc2d.py:16-25); kept as names, the
# meaningful timings are CUDA events in bench.py
TIME = TIME_U = TIME_V = TIME_P = TIME_T = TIME_Update = TIME_Divergence = TIME_FFT = TIME_Conv = 0


class NavierStokes(NavierStokesBase, NavierStokesSteadyState, Integrator):
    """
    rbc:    adiabatic side walls
    linear: isothermal side walls, linear temperature profile
    zero:   isothermal side walls, zero side-wall temperature
    """

    avail_cases = ["rbc", "linear", "zero"]

    def __init__(self, case="rbc", dealias_grid="fft", stepper="fast", graph=False, slab=False, **kwargs):
        """dealias_grid: "fft" (default) evaluates the 3/2-rule products on the next
        FFT-friendly Gauss-Lobatto grid >= 3N/2 (same truncated coefficients up to rounding);
        "reference" uses exactly int(3N/2) points like the reference.
        stepper: "fast" (default) runs the stage as fused axis passes (pass_stepper.PassStepper; grids with
        odd sizes fall back to "batched"), "batched" one batched launch per operator
        (fast_stepper.FastStepper), "reference" the operator-by-operator sequence of the reference
        (update_reference).
        graph: capture the whole time step in a CUDA graph (removes launch overhead on
        small grids).
        slab: distribute the step over the ranks of the default torch.distributed process group
        (one process per GPU, slab decomposition with NCCL all-to-all transposes, navier/slab.py);
        the fields of this object then hold the replicated initial state and are refreshed by
        sync_fields()."""
        if case not in self.avail_cases:
            raise ValueError("Specified case is not available: ", self.avail_cases)
        self.case = case
        self.dealias_grid = dealias_grid
        self._stepper_kind = stepper
        self._use_graph = graph
        self._slab = slab
        self._fast = None
        self._graph = None
        with dealias_policy(dealias_grid):
            self._construct(**kwargs)

    def _construct(self, **kwargs):
        NavierStokesBase.__init__(self, **kwargs)
        Integrator.__init__(self)

        side = "CN" if self.case == "rbc" else "CD"
        self.set_fieldbc = self.set_temp_fieldbc_zero if self.case == "zero" else self.set_temp_fieldbc_linear

        N0, N1 = self.shape
        self.T = Field([Base(N0, side, dealias=3 / 2), Base(N1, "CD", dealias=3 / 2)])
        self.U = Field([Base(N0, "CD", dealias=3 / 2), Base(N1, "CD", dealias=3 / 2)])
        self.V = Field([Base(N0, "CD", dealias=3 / 2), Base(N1, "CD", dealias=3 / 2)])
        self.P = Field([Base(N0, "CN"), Base(N1, "CN")])
        self.pres = Field([Base(N0, "CH"), Base(N1, "CH")])
        self.field = MultiField([self.T, self.U, self.V, self.pres], ["temp", "ux", "uy", "pres"])

        self.setup_solver()
        self.set_fieldbc()
        self.rhs = torch.zeros(self.shape, dtype=torch.float64, device=C.device())

    def reset(self, reset_time=True):
        """Call after ra / pr changed."""
        self.set_nu_kappa()
        self.setup_solver()
        if reset_time:
            self.reset_time()

    def reset_time(self):
        self.time = 0.0
        for field in self.field.fields:
            field.time = 0.0

    # -- initial conditions (rbc2d.py:122-133) ---------------------------------------------
    def set_temperature(self, amplitude=0.5, m=1):
        self.T.v = amplitude * np.sin(m * np.pi * self.xx) * np.cos(np.pi * self.yy)
        self.T.forward()

    def set_velocity(self, amplitude=0.5, m=1, n=1):
        x = (self.x - self.x[0]) / (self.x[-1] - self.x[0])
        y = (self.y - self.y[0]) / (self.y[-1] - self.y[0])
        xx, yy = np.meshgrid(x, y, indexing="ij")
        self.U.v = -amplitude * np.sin(m * np.pi * xx) * np.cos(n * np.pi * yy)
        self.V.v = amplitude * np.cos(m * np.pi * xx) * np.sin(n * np.pi * yy)
        self.U.forward()
        self.V.forward()

    # -- temperature boundary lift (rbc2d.py:135-178) --------------------------------------
    def _finish_fieldbc(self):
        self.dTbcdz2 = self.grad(self.Tbc, deriv=(0, 2))
        vhat = self.grad(self.Tbc, deriv=(0, 1))
        space = self.deriv_field.dealias if self.dealias else self.deriv_field
        self.dTbcdz1 = space.backward(vhat)
        self.Tbc_cheby = galerkin_to_cheby(self.Tbc.vhat, self.Tbc)
        self._fast = None          # the batched stepper snapshots Tbc_cheby / dTbcdz1 / dTbcdz2
        self._graph = None

    def set_temp_fieldbc_linear(self):
        bc = np.zeros((self.shape[0], 2))
        bc[:, 0], bc[:, 1] = 0.5, -0.5
        self.Tbc = FieldBC(self.T.xs, axis=1)
        self.Tbc.add_bc(bc)
        self._finish_fieldbc()

    def set_temp_fieldbc_zero(self):
        bc = np.zeros((2, self.shape[1]))
        bc[0, :] = transfer_function(0.5, 0, -0.5, self.y, k=0.02)
        bc[1, :] = bc[0, :]
        self.Tbc = FieldBC(self.T.xs, axis=0)
        self.Tbc.add_bc(bc)
        self._finish_fieldbc()

    # -- solver plans (rbc2d.py:180-211) --------------------------------------------------
    def _setup_solver_plans(self):
        from ..templates.hholtz import solverplan_hholtz2d_adi
        from ..templates.poisson import solverplan_poisson2d

        if self.integrator == "rk3":
            self.set_timestep_coefficients_rk3()
        else:
            self.set_timestep_coefficients_euler()
        self.solver_U, self.solver_V, self.solver_T = [], [], []
        for rk in range(self.nstage):
            lam_nu = self.dt * self.a[rk] * self.beta * self.nu
            lam_ka = self.dt * self.a[rk] * self.beta * self.kappa
            self.solver_U.append(solverplan_hholtz2d_adi(bases=self.U.xs, lam=lam_nu, scale=self.scale))
            self.solver_V.append(solverplan_hholtz2d_adi(bases=self.V.xs, lam=lam_nu, scale=self.scale))
            self.solver_T.append(solverplan_hholtz2d_adi(bases=self.T.xs, lam=lam_ka, scale=self.scale))
        self.solver_P = solverplan_poisson2d(self.P.xs, singular=True, scale=self.scale)

    # -- building blocks of a stage ---------------------------------------------------------
    def update_velocity(self, p, u, v, fac=1.0):
        """Pressure projection: u -= grad(p) mapped back to the Galerkin space (rbc2d.py:213-223)."""
        dpdx = self.grad(p, deriv=(1, 0))
        dpdz = self.grad(p, deriv=(0, 1))
        u.vhat -= cheby_to_galerkin(dpdx * fac, u)
        v.vhat -= cheby_to_galerkin(dpdz * fac, v)

    def divergence_velocity(self, u, v):
        return self.grad(u, deriv=(1, 0)) + self.grad(v, deriv=(0, 1))

    def conv_term(self, field, ux, uz, add_bc=None):
        return convective_term(field, ux, uz, deriv_field=self.deriv_field, add_bc=add_bc,
                               dealias=self.dealias, scale=self.scale)

    def _explicit_diffusion(self, rhs, field, coef, stage):
        if self.beta != 1.0:
            f = float(self.dt * self.a[stage] * (1 - self.beta) * coef)
            rhs += f * self.grad(field, deriv=(2, 0))
            rhs += f * self.grad(field, deriv=(0, 2))

    def _helmholtz(self, solver, rhs, field):
        rhs = solver.solve_rhs(rhs)
        rhs += solver.solve_old(field.vhat)
        field.vhat[:] = solver.solve_lhs(rhs)

    def update_U(self, stage):
        dpdx = self.grad(self.pres, deriv=(1, 0))
        rhs = float(-self.dt * self.a[stage]) * dpdx
        rhs -= float(self.dt * self.b[stage]) * self.conv_term(self.U, self.ux, self.uz)
        if self.c[stage] != 0:
            rhs -= float(self.dt * self.c[stage]) * self.conv_term(self.U, self.ux_old, self.uz_old)
        self._explicit_diffusion(rhs, self.U, self.nu, stage)
        self._helmholtz(self.solver_U[stage], rhs, self.U)

    def update_V(self, That, stage):
        dpdz = self.grad(self.pres, deriv=(0, 1))
        rhs = float(-self.dt * self.a[stage]) * dpdz
        rhs -= float(self.dt * self.b[stage]) * self.conv_term(self.V, self.ux, self.uz)
        if self.c[stage] != 0:
            rhs -= float(self.dt * self.c[stage]) * self.conv_term(self.V, self.ux_old, self.uz_old)
        rhs += float(self.dt * self.a[stage]) * That
        self._explicit_diffusion(rhs, self.V, self.nu, stage)
        self._helmholtz(self.solver_V[stage], rhs, self.V)

    def update_T(self, stage):
        rhs = float(-self.dt * self.b[stage]) * self.conv_term(self.T, self.ux, self.uz, add_bc=self.uz * self.dTbcdz1)
        if self.c[stage] != 0:
            rhs -= float(self.dt * self.c[stage]) * self.conv_term(self.T, self.ux_old, self.uz_old,
                                                            add_bc=self.uz_old * self.dTbcdz1)
        rhs += float(self.dt * self.a[stage] * self.kappa) * self.dTbcdz2
        self._explicit_diffusion(rhs, self.T, self.kappa, stage)
        self._helmholtz(self.solver_T[stage], rhs, self.T)

    def update_P(self, div, singular=True):
        rhs = self.solver_P.solve_rhs(div)
        self.P.vhat[:] = self.solver_P.solve_lhs(rhs)
        if singular:
            self.P.vhat[0, 0] = 0

    def update_pres(self, div, stage):
        self.pres.vhat -= float(1.0 * self.nu) * div * float(self.beta)
        self.pres.vhat += float(1.0 / (self.dt * self.a[stage])) * galerkin_to_cheby(self.P.vhat, self.P)

    def setup_solver(self):
        self._setup_solver_plans()
        self._fast = None          # tables changed: rebuild the batched stepper lazily
        self._graph = None

    def update(self):
        """One time step = nstage IMEX stages (rbc2d.py:396-434)."""
        if self._stepper_kind not in ("fast", "batched") or self.beta != 1.0:
            return self.update_reference()
        if self._fast is None:
            if self._slab:
                import torch.distributed as dist
                from .pass_slab_stepper import PassSlabStepper
                if self._stepper_kind == "fast" and PassSlabStepper.supported(self, dist.get_world_size()):
                    # peer-memory row passes, no NCCL on the data path; "yfirst": the 8-exchange schedule
                    import os
                    if os.environ.get("PDE_SLAB_SCHEDULE", "yfirst") == "yfirst":
                        from .pass_slab_stepper2 import PassSlabStepperY
                        self._fast = PassSlabStepperY(self)
                    else:
                        self._fast = PassSlabStepper(self)
                else:
                    from .slab_stepper import SlabStepper
                    self._fast = SlabStepper(self)       # NCCL all-to-all transposes ("batched")
            else:
                from .fast_stepper import FastStepper
                from .pass_stepper import PassStepper
                if self._stepper_kind == "fast" and PassStepper.supported(self):
                    self._fast = PassStepper(self)       # fused axis passes (even grids up to 4096)
                else:
                    self._fast = FastStepper(self)       # one launch per operator ("batched", or odd sizes)
        if not self._use_graph:
            for rk in range(self.nstage):
                self._fast.stage(rk)
        else:
            self._update_graph()
        self.ux, self.uz = self._fast.ux, self._fast.uz

    # -- parts of the reference class that are outside the time-step path (out of scope) ------
    def solve_stability(self, *args, **kwargs):
        raise NotImplementedError("NavierStokesStability (navier/rbc2d_base.py:389-452: dense host eigenproblems of "
                                  "pypde/stability) is not part of pypde_b200")

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting (pypde/plot) is not part of pypde_b200; use field.v.cpu().numpy()")

    def animate(self, *args, **kwargs):
        raise NotImplementedError("plotting (pypde/plot) is not part of pypde_b200; use field.V after save()")

    def sync_fields(self):
        """Slab mode: gather the distributed state into T, U, V, pres of every rank."""
        if self._slab and self._fast is not None:
            self._fast.gather()

    def close(self):
        """Slab mode: release the CUDA graph and the peer mappings (call on every rank before
        torch.distributed.destroy_process_group)."""
        import torch
        torch.cuda.synchronize()
        self._graph = None
        fs, self._fast = self._fast, None
        if fs is not None and hasattr(fs, "close"):
            fs.close()

    def _update_graph(self):
        fs = self._fast
        if self._slab:
            stale = not fs.bound
        else:
            cur = tuple(t.data_ptr() for t in (self.T.vhat, self.U.vhat, self.V.vhat, self.P.vhat, self.pres.vhat))
            stale = cur != fs.bound
        if self._graph is None or stale:
            fs.bind()
            for rk in range(self.nstage):      # warm-up outside the capture (plan tables, attributes, NCCL)
                fs.stage_calls[rk].run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for rk in range(self.nstage):
                    fs.stage_calls[rk].run()
            self._graph = g
            # the warm-up advanced the state by one step already
            return
        self._graph.replay()

    def update_reference(self):
        """One time step, operator by operator in the reference's order (rbc2d.py:396-434)."""
        self.ux_old, self.uz_old = 0, 0
        for rk in range(self.nstage):
            That = galerkin_to_cheby(self.T.vhat, self.T)
            That += self.Tbc_cheby
            if self.dealias:
                self.ux = self.U.dealias.backward(self.U.vhat)
                self.uz = self.V.dealias.backward(self.V.vhat)
            else:
                self.ux = self.U.backward(self.U.vhat)
                self.uz = self.V.backward(self.V.vhat)
            self.update_U(stage=rk)
            self.update_V(That, stage=rk)
            div = self.divergence_velocity(self.U, self.V)
            self.update_P(div)
            self.update_pres(div, stage=rk)
            self.update_velocity(self.P, self.U, self.V)
            self.update_T(stage=rk)
            self.ux_old, self.uz_old = self.ux, self.uz


def transfer_function(TL, TM, TR, x, k=0.01):
    """Smooth side-wall temperature profile (rbc2d.py:437-446)."""
    arr = np.zeros(x.shape)
    L = x[-1] - x[0]
    for i in range(x.size):
        xs = x[i] * 2.0 / L
        if xs < 0:
            arr[i] = -k * xs / (k + xs + 1) * (TL - TM) + TM
        else:
            arr[i] = k * xs / (k - xs + 1) * (TR - TM) + TM
    return arr
