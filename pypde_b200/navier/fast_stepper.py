"""
Batched execution of one IMEX stage of navier.rbc2d.NavierStokes (navier/rbc2d.py:396-434).

Same operator sequence as the reference-ordered `NavierStokes.update_reference()`, but
  * every operator is applied to all fields that need it in ONE launch (batched C-ABI entry
    points: pde_sweep / pde_to_cheb_multi / pde_banded_multi / pde_lincomb_multi / pde_dct1_multi),
  * all work arrays are allocated once, all launch descriptors are built once (`bind`), so a
    stage is a fixed list of ~45 C-ABI calls with no tensor allocation - which also makes the
    whole RK3 step capturable in a CUDA graph (`NavierStokes(graph=True)`),
  * the two convective terms of an RK3 stage (current and previous-stage velocities,
    rbc2d.py:260-266) are merged by linearity into one product with ub = b u + c u_old, so a
    stage needs 8 backward and 3 forward 2-D transforms instead of 14 + 6,
  * the T-equation (which only needs stage-start data, rbc2d.py:335-378) is solved together with
    U and V.
Differences to the reference are rounding-level only (re-association of linear operations);
parity is checked in tests/test_gpu_rbc.py against the CPU oracle.
"""
import ctypes

import numpy as np
import torch

from .. import _cabi as C
from .. import ops

DIFF, TDMA_FWD, TDMA_BWD, FDMA_FWD, FDMA_BWD = 0, 1, 2, 3, 4


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _ld(t):
    return t.stride(0) if t.dim() == 2 else 1


class _Calls:
    """A fixed list of (C function, argument tuple) launched on the current stream."""

    def __init__(self):
        self.calls = []
        self.labels = []    # one name per call (bench.py's per-kernel timing)
        self.keep = []      # ctypes arrays / tensors that must stay alive

    def add(self, fn, *args, label=None):
        self.calls.append((fn, args))
        self.labels.append(label or getattr(fn, "__name__", "call"))

    def run(self):
        st = C.stream()
        for fn, args in self.calls:
            rc = fn(*args, st)
            if rc != 0:
                C.check(rc)


class FastStepper:
    def __init__(self, ns):
        self.ns = ns
        if ns.beta != 1.0:
            raise ValueError("FastStepper covers beta == 1 (fully implicit diffusion)")
        self.dev = C.device()
        self.N0, self.N1 = ns.shape
        self.M0, self.M1 = self.N0 - 2, self.N1 - 2
        space = ns.deriv_field.dealias if ns.dealias else ns.deriv_field
        self.D0, self.D1 = space.shape_physical
        self.plan0 = ops.DctPlan.get(self.D0)
        self.plan1 = ops.DctPlan.get(self.D1)
        self._alloc()
        self._tables()
        self.bound = None
        self.stage_calls = None

    # ------------------------------------------------------------------ buffers
    def _new(self, *shape):
        """Work array.  Rows whose byte length is a multiple of 2 KB are padded by 64 bytes: the
        axis-0 kernels walk columns with the row pitch as stride, and a power-of-two pitch maps
        every row of a column strip onto the same few L2 slices / HBM channels."""
        if len(shape) == 2 and shape[1] % 256 == 0 and shape[0] > 1:
            return torch.zeros((shape[0], shape[1] + 8), dtype=torch.float64, device=self.dev)[:, : shape[1]]
        return torch.zeros(shape, dtype=torch.float64, device=self.dev)

    def _alloc(self):
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        self.c3 = [self._new(N0, M1) for _ in range(3)]       # Sx F
        self.d3 = [self._new(N0, M1) for _ in range(3)]       # dx Sx F / sx
        self.e6 = [self._new(N0, N1) for _ in range(6)]       # eU eV eT fU fV fT
        self.g3 = [self._new(N0, N1) for _ in range(3)]       # dz(e)/sz
        self.thc = self._new(N0, N1)
        self.X8 = [self._new(D0, N1) for _ in range(8)]
        # pde_conv_products indexes its 13 arrays flat over D0*D1: these stay contiguous (no pitch padding)
        flat = lambda: torch.zeros((D0, D1), dtype=torch.float64, device=self.dev)
        self.phys = [flat() for _ in range(6)]                # dxU dxV dxT dzU dzV dzT
        self.uw = [[flat(), flat()] for _ in range(2)]
        self.F3 = [self._new(D0, N1) for _ in range(3)]
        self.conv = [self._new(N0, N1) for _ in range(3)]
        self.dpdx, self.dpdz = self._new(N0, N1), self._new(N0, N1)
        self.rhs = [self._new(N0, N1) for _ in range(3)]
        self.gh = [self._new(M0, N1) for _ in range(3)]
        self.hh = [self._new(M0, M1) for _ in range(3)]
        self.rr = [self._new(M0, M1) for _ in range(3)]
        self.div = self._new(N0, N1)
        self.q = self._new(M0, N1)
        self.R = self._new(M0, M1)

    def _tables(self):
        ns = self.ns
        up = C.upload
        self.sx = {"T": ns.T.xs[0]._tables()[0], "U": ns.U.xs[0]._tables()[0], "V": ns.V.xs[0]._tables()[0],
                   "P": ns.P.xs[0]._tables()[0]}
        self.sy = {"T": ns.T.xs[1]._tables()[0], "U": ns.U.xs[1]._tables()[0], "V": ns.V.xs[1]._tables()[0],
                   "P": ns.P.xs[1]._tables()[0]}
        # from_cheb tables (s, a, den, w) + reciprocal of den, for the velocity spaces
        self.inv = {}
        for name, fld in (("U", ns.U), ("V", ns.V)):
            for ax in (0, 1):
                s, a, den, w = fld.xs[ax]._tables()
                self.inv[(name, ax)] = (s, a, den, w, (1.0 / den))
        # Helmholtz LU tables (+ reciprocal diagonal) per stage / field / axis
        self.lu = {}
        for rk in range(ns.nstage):
            for name, solver in (("U", ns.solver_U[rk]), ("V", ns.solver_V[rk]), ("T", ns.solver_T[rk])):
                for ax in (0, 1):
                    pl = solver.plan_for_lhs[ax]
                    self.lu[(rk, name, ax)] = tuple(pl._t) + (up(1.0 / pl.d),)
        self.tbc_cheby = C.to_dev(ns.Tbc_cheby).contiguous()
        self.dTbcdz2 = C.to_dev(ns.dTbcdz2).contiguous()
        self.dTbcdz1 = C.to_dev(ns.dTbcdz1).contiguous()

    # ------------------------------------------------------------------ job builders
    def _stencil(self, calls, axis, jobs):
        """jobs: (s, v, u) with v (.., M ..) -> u (.., n_out ..)"""
        arr = (C.StencilJob * len(jobs))()
        for k, (s, v, u) in enumerate(jobs):
            j = arr[k]
            j.s, j.v, j.ldv, j.M = _ptr(s), _ptr(v), _ld(v), v.shape[axis]
            j.u, j.ldu, j.n_out, j.batch = _ptr(u), _ld(u), u.shape[axis], u.shape[1 - axis]
            assert v.shape[1 - axis] == u.shape[1 - axis]
        calls.keep.append(arr)
        calls.add(C.lib().pde_to_cheb_multi, axis, len(jobs), arr)

    def _sweep(self, calls, op, axis, n, jobs):
        """jobs: dicts with in (list of tensors), out, tab (dict idx->tensor), flag, sc"""
        arr = (C.SweepJob * len(jobs))()
        for k, jb in enumerate(jobs):
            j = arr[k]
            for s, t in enumerate(jb["in"]):
                if t is not None:
                    j.inp[s], j.ldin[s] = _ptr(t), _ld(t)
            out = jb["out"]
            j.out, j.ldout = _ptr(out), _ld(out)
            for idx, t in jb.get("tab", {}).items():
                j.tab[idx] = _ptr(t)
            j.nseq = out.shape[1 - axis]
            j.flag = int(jb.get("flag", 0))
            j.sc = float(jb.get("sc", 1.0))
        calls.keep.append(arr)
        calls.add(C.lib().pde_sweep, op, axis, n, len(jobs), arr)

    def _diff(self, calls, axis, pairs, div):
        n = pairs[0][0].shape[axis]
        self._sweep(calls, DIFF, axis, n, [dict(**{"in": [c]}, out=d, flag=div != 1.0, sc=div) for c, d in pairs])

    def _band(self, calls, axis, jobs, accumulate=False):
        """jobs: (Band, x, y)"""
        arr = (C.BandJob * len(jobs))()
        for k, (band, x, y) in enumerate(jobs):
            j = arr[k]
            j.diags, j.ndiag = _ptr(band.diags), band.ndiag
            for d, o in enumerate(band.offsets):
                j.off[d] = o
            j.x, j.ldx, j.n_in = _ptr(x), _ld(x), band.n_in
            j.y, j.ldy, j.n_out = _ptr(y), _ld(y), band.n_out
            j.batch, j.accumulate = y.shape[1 - axis], int(accumulate)
            assert x.shape[axis] == band.n_in and y.shape[axis] == band.n_out
        calls.keep.append(arr)
        calls.add(C.lib().pde_banded_multi, axis, len(jobs), arr)

    def _lincomb(self, calls, jobs):
        """jobs: (y, [(coef, x), ...])"""
        arr = (C.LincombJob * len(jobs))()
        for k, (y, terms) in enumerate(jobs):
            j = arr[k]
            j.nterm = len(terms)
            for t, (coef, x) in enumerate(terms):
                j.x[t], j.ldx[t], j.coef[t] = _ptr(x), _ld(x), float(coef)
                assert tuple(x.shape) == tuple(y.shape)
            j.y, j.ldy, j.n0, j.n1 = _ptr(y), _ld(y), y.shape[0], y.shape[1]
        calls.keep.append(arr)
        calls.add(C.lib().pde_lincomb_multi, len(jobs), arr)

    def _dct(self, calls, plan, mode, axis, xs, ys):
        nj = len(xs)
        xa = (ctypes.c_void_p * nj)(*[_ptr(x) for x in xs])
        ya = (ctypes.c_void_p * nj)(*[_ptr(y) for y in ys])
        calls.keep += [xa, ya]
        x, y = xs[0], ys[0]
        calls.add(C.lib().pde_dct1_multi, plan.handle, mode, nj, xa, _ld(x), x.shape[axis], ya, _ld(y),
                  y.shape[axis], x.shape[1 - axis], axis)

    def _solve4(self, calls, rk, fields, xs, outs):
        """4-diagonal Helmholtz solves: x sweeps then y sweeps; the last sweep writes `outs`."""
        for ax in (0, 1):
            n = xs[0].shape[ax]
            fw, bw = [], []
            for name, x, out in zip(fields, xs, outs):
                l, d, u1, u2, rd = self.lu[(rk, name, ax)]
                fw.append(dict(**{"in": [x]}, out=x, tab={0: l}))
                bw.append(dict(**{"in": [x]}, out=(out if ax == 1 else x), tab={1: d, 2: u1, 3: u2, 4: rd}))
            self._sweep(calls, FDMA_FWD, ax, n, fw)
            self._sweep(calls, FDMA_BWD, ax, n, bw)

    def _from_cheb(self, calls, axis, names, us, vs):
        n = vs[0].shape[axis]
        fw, bw = [], []
        for name, u, v in zip(names, us, vs):
            s, a, den, w, rden = self.inv[(name, axis)]
            fw.append(dict(**{"in": [u, u]}, out=v, tab={0: s, 1: a, 2: den, 4: rden}))
            bw.append(dict(**{"in": [v]}, out=v, tab={3: w}))
        self._sweep(calls, TDMA_FWD, axis, n, fw)
        self._sweep(calls, TDMA_BWD, axis, n, bw)

    # ------------------------------------------------------------------ the stage
    def bind(self):
        """(Re)build the launch lists for the tensors currently held by the fields."""
        ns = self.ns
        T, U, V, P, pres = ns.T.vhat, ns.U.vhat, ns.V.vhat, ns.P.vhat, ns.pres.vhat
        for t in (T, U, V, P, pres):
            assert t.stride(1) == 1
        self.bound = tuple(t.data_ptr() for t in (T, U, V, P, pres))
        self.stage_calls = [self._build_stage(rk, T, U, V, P, pres) for rk in range(ns.nstage)]

    def _build_stage(self, rk, T, U, V, P, pres):
        ns, L = self.ns, C.lib()
        calls = _Calls()
        sx, sz = ns.scale
        dt, a, b, c = float(ns.dt), float(ns.a[rk]), float(ns.b[rk]), float(ns.c[rk])
        cU, cV, cT = self.c3
        dU, dV, dT = self.d3
        eU, eV, eT, fU, fV, fT = self.e6
        gU, gV, gT = self.g3
        new, old = self.uw[rk % 2], self.uw[(rk + 1) % 2]
        dxU, dxV, dxT, dzU, dzV, dzT = self.phys

        # -- spectral pre-processing: Chebyshev coefficients of u, w and of all first derivatives
        self._stencil(calls, 0, [(self.sx["U"], U, cU), (self.sx["V"], V, cV), (self.sx["T"], T, cT)])
        # (the pressure gradient of the right-hand sides only needs stage-start data: its two single-array
        # recurrences ride along with the batched ones instead of occupying one warp per SM on their own)
        self._diff(calls, 0, [(cU, dU), (cV, dV), (cT, dT), (pres, self.dpdx)], sx)
        self._stencil(calls, 1, [(self.sy["U"], cU, eU), (self.sy["V"], cV, eV), (self.sy["T"], cT, eT),
                                 (self.sy["U"], dU, fU), (self.sy["V"], dV, fV), (self.sy["T"], dT, fT)])
        self._diff(calls, 1, [(eU, gU), (eV, gV), (eT, gT), (pres, self.dpdz)], sz)
        self._lincomb(calls, [(self.thc, [(1.0, eT), (1.0, self.tbc_cheby)])])      # That (buoyancy)
        # -- 8 backward 2-D transforms onto the (dealiased) grid
        src = [eU, eV, fU, fV, fT, gU, gV, gT]
        dst = [new[0], new[1], dxU, dxV, dxT, dzU, dzV, dzT]
        self._dct(calls, self.plan0, ops.BWD, 0, src, self.X8)
        self._dct(calls, self.plan1, ops.BWD, 1, self.X8, dst)
        # -- products (both convective terms of the stage merged: ub = b u + c u_old)
        use_old = c != 0.0
        for t in list(new) + list(old) + list(self.phys) + [self.dTbcdz1]:
            assert t.is_contiguous() and tuple(t.shape) == (self.D0, self.D1)
        calls.add(L.pde_conv_products, self.D0 * self.D1, b, c, _ptr(new[0]), _ptr(new[1]),
                  _ptr(old[0]) if use_old else None, _ptr(old[1]) if use_old else None,
                  _ptr(dxU), _ptr(dzU), _ptr(dxV), _ptr(dzV), _ptr(dxT), _ptr(dzT), _ptr(self.dTbcdz1))
        # -- 3 forward transforms, truncated to N coefficients
        self._dct(calls, self.plan1, ops.FWD, 1, [dxU, dxV, dxT], [f[:, : self.N1] for f in self.F3])
        self._dct(calls, self.plan0, ops.FWD, 0, [f[:, : self.N1] for f in self.F3], [cv[: self.N0] for cv in self.conv])
        # -- right-hand sides (Chebyshev space)
        rU, rV, rT = self.rhs
        self._lincomb(calls, [
            (rU, [(-dt * a, self.dpdx), (-dt, self.conv[0])]),
            (rV, [(-dt * a, self.dpdz), (-dt, self.conv[1]), (dt * a, self.thc)]),
            (rT, [(-dt, self.conv[2]), (dt * a * ns.kappa, self.dTbcdz2)]),
        ])
        # -- Helmholtz: r = By Bx rhs + (By Sy)(Bx Sx) F, then the ADI solves
        solvers = (ns.solver_U[rk], ns.solver_V[rk], ns.solver_T[rk])
        state = (U, V, T)
        self._band(calls, 0, [(s.plan_for_rhs[0].band, r, g) for s, r, g in zip(solvers, self.rhs, self.gh)] +
                   [(s.plan_for_old[0].band, f, h) for s, f, h in zip(solvers, state, self.hh)])
        self._band(calls, 1, [(s.plan_for_rhs[1].band, g, r) for s, g, r in zip(solvers, self.gh, self.rr)])
        self._band(calls, 1, [(s.plan_for_old[1].band, h, r) for s, h, r in zip(solvers, self.hh, self.rr)],
                   accumulate=True)
        self._solve4(calls, rk, ("U", "V", "T"), self.rr, state)
        # -- divergence of the intermediate velocity
        self._stencil(calls, 0, [(self.sx["U"], U, cU), (self.sx["V"], V, cV)])
        self._diff(calls, 0, [(cU, dU)], sx)
        self._stencil(calls, 1, [(self.sy["U"], dU, eU), (self.sy["V"], cV, eV)])
        self._diff(calls, 1, [(eV, gV)], sz)
        self._lincomb(calls, [(self.div, [(1.0, eU), (1.0, gV)])])
        # -- pressure Poisson solve (eigen-decomposition along y)
        sp = ns.solver_P
        self._band(calls, 0, [(sp.plan_for_rhs[0].band, self.div, self.q)])
        Hy, Qy = sp.plan_for_rhs[1].dense, sp.plan_for_lhs[1].dense
        calls.add(L.pde_gemm_f64, 1, _ptr(self.q), _ld(self.q), _ptr(Hy), _ld(Hy), _ptr(self.R), _ld(self.R),
                  self.M0, self.M1, self.N1)
        calls.add(L.pde_poisson_solve, sp.plan_for_lhs[0]._plan.handle, _ptr(self.R), _ld(self.R))
        calls.add(L.pde_gemm_f64, 1, _ptr(self.R), _ld(self.R), _ptr(Qy), _ld(Qy), _ptr(P), _ld(P),
                  self.M0, self.M1, self.M1)
        zero = self._new(1, 1)
        calls.keep.append(zero)
        self._lincomb(calls, [(P[0:1, 0:1], [(1.0, zero)])])                          # P[0, 0] = 0
        # -- pressure update and velocity correction
        self._stencil(calls, 0, [(self.sx["P"], P, cT)])                             # Sx P
        self._diff(calls, 0, [(cT, dT)], sx)
        self._stencil(calls, 1, [(self.sy["P"], dT, fU), (self.sy["P"], cT, fV)])    # dpdx, SxSy P
        self._diff(calls, 1, [(fV, gT)], sz)                                         # dpdz
        self._lincomb(calls, [(pres, [(1.0, pres), (-(1.0 * ns.nu), self.div), (1.0 / (dt * a), fV)])])
        tU, tV = self.gh[0], self.gh[1]
        self._from_cheb(calls, 0, ("U", "V"), [fU, gT], [tU, tV])
        wU, wV = self.hh[0], self.hh[1]
        self._from_cheb(calls, 1, ("U", "V"), [tU, tV], [wU, wV])
        self._lincomb(calls, [(U, [(1.0, U), (-1.0, wU)]), (V, [(1.0, V), (-1.0, wV)])])
        calls.keep += [T, U, V, P, pres]
        return calls

    def stage(self, rk):
        ns = self.ns
        cur = tuple(t.data_ptr() for t in (ns.T.vhat, ns.U.vhat, ns.V.vhat, ns.P.vhat, ns.pres.vhat))
        if cur != self.bound:
            self.bind()
        self.stage_calls[rk].run()

    @property
    def ux(self):
        return self.uw[(self.ns.nstage - 1) % 2][0]

    @property
    def uz(self):
        return self.uw[(self.ns.nstage - 1) % 2][1]
