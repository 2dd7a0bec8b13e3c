"""
One IMEX stage of navier.rbc2d.NavierStokes (navier/rbc2d.py:396-434) as 8 fused axis passes + the
transforms + the two dense projections: 15 launches instead of the 39 of fast_stepper.FastStepper.

Every operator of the stage acts along one axis, so the stage is regrouped into passes that load each
sequence once, apply a chain of operators in shared memory (pde_pass_run, csrc/pass.cu) and store it once
("axis-pass model").  X = along axis 0 (per column), Y = along axis 1 (per row):

    PX1  F -> Sx F, dx Sx F / sx                         (F = U, V, T;  pres -> dpdx)
    PY2  -> eF = Sy Sx F, gF = dz eF / sz, fF = Sy dx Sx F / sx;
         rest_F = eF + (pressure gradient, buoyancy, boundary terms)
         8 backward 2-D DCTs, products, 3 forward 2-D DCTs    (dct_fft*.cu, batched.cu; unchanged)
    PX3  zF = Ax^-1 Bx (rest_F - dt conv_F)
    PY4  F* = Ay^-1 By zF;  y parts of div(U*, V*)
    PX5  div, q = Bx div
         R = q Hy^T                                          (gemm.cu)
    PX6  per-column Poisson solves (A + lam_i C) w_i = r_i
         P = W Qy^T
    PY7  P[0,0] = 0;  e1 = Sy P;  bU = Gy e1,  bV = Gy dz e1 / sz       (G = Chebyshev -> Galerkin map)
    PX8  U = U* - Gx dx Sx bU / sx,  V = V* - Gx Sx bV,  pres += Sx e1 / (dt a) - nu div

Identities used (exact for commuting tensor-product operators; rounding-level differences only, checked
against the CPU oracle in tests/test_gpu_rbc.py / test_gpu_large.py):
    Ay^-1 Ax^-1 [By Bx rhs + (By Sy)(Bx Sx) F]  =  [Ay^-1 By] [Ax^-1 Bx] (rhs + Sy Sx F)
    cheby_to_galerkin = Gy Gx = Gx Gy

The stepper advances a LIST of runs of one grid ("members": parameter-sweep ensembles, BASELINE.json
configs[4]); a single NavierStokes object is the one-member case.  Every launch carries the jobs of all members:
the passes take one job per member and array, the transforms / projections / products use the batched entry
points (device pointer arrays), so an ensemble step costs 15 launches per stage however many members it has.
"""
import ctypes

import torch

from .. import _cabi as C
from .. import ops
from .. import passes as PS
from .fast_stepper import FastStepper, _Calls, _ptr, _ld


class _Member:
    """Work arrays of one run."""


class PassStepper(FastStepper):
    @staticmethod
    def supported(ns):
        N0, N1 = ns.shape
        return ns.beta == 1.0 and N0 % 2 == 0 and N1 % 2 == 0 and 8 <= N0 <= 4096 and 8 <= N1 <= 4096

    def __init__(self, ns, members=None):
        """ns: the run (or the first member); members: all runs advanced together (same grid, dt, integrator)."""
        self.tables = PS.TableCache()
        self.runs = list(members) if members is not None else [ns]
        for m in self.runs:
            assert tuple(m.shape) == tuple(ns.shape) and m.dt == ns.dt and m.integrator == ns.integrator \
                and m.dealias == ns.dealias and m.case == ns.case and tuple(m.scale) == tuple(ns.scale), \
                "ensemble members share grid, time step, integrator and boundary case"
        FastStepper.__init__(self, ns)

    # ------------------------------------------------------------------ buffers
    def _stack(self, rows, cols, exact=False):
        """(members, rows, cols) views of ONE allocation (constant member stride).  Row passes move 16-byte units:
        even leading dimensions; rows whose byte length is a multiple of 2 KB get 64 bytes of padding like in
        FastStepper (column walks of the x-transforms)."""
        nm = len(self.runs)
        ld = cols + (cols & 1)      # 16-byte aligned rows: the DCT row kernel stages them with bulk copies
        if not exact:
            assert cols % 2 == 0
            ld = cols + 8 if cols % 256 == 0 else cols
        big = torch.zeros((nm, rows, ld), dtype=torch.float64, device=self.dev)
        return [big[m, :, :cols] for m in range(nm)]

    def _alloc(self):
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        self.mb = [_Member() for _ in self.runs]

        def give(name, rows, cols, count=0, exact=False):
            for k in range(max(count, 1)):
                for m, v in zip(self.mb, self._stack(rows, cols, exact)):
                    if count == 0:
                        setattr(m, name, v)
                    else:
                        if k == 0:
                            setattr(m, name, [])
                        getattr(m, name).append(v)
        for name in ("c3", "d3"):
            give(name, N0, M1, 3)
        for name in ("e3", "f3", "g3", "rest3", "conv"):
            give(name, N0, N1, 3)
        give("dpdx", N0, N1)
        # transforms: D x D arrays may have odd widths (3073): pitch rounded up to even, only the DCT / product kernels
        # touch them (the products run over the padded rows as one contiguous range: the pad column stays zero)
        give("X8", D0, N1, 8)
        give("phys", D0, D1, 6, exact=True)
        give("uwa", D0, D1, 2, exact=True)
        give("uwb", D0, D1, 2, exact=True)
        give("F3", D0, N1, 3)
        give("z3", M0, N1, 3)
        give("aU", N0, N1)          # M0 rows used; the rest stays zero (stencil-on-load)
        give("aV", N0, N1)
        give("div", N0, N1)
        give("q", M0, N1)
        give("R", M0, M1)
        give("e1", N0, N1)          # M0 rows used
        give("bU", N0, M1)
        give("bV", N0, M1)
        for m in self.mb:
            m.uw = [m.uwa, m.uwb]
        self.uw = self.mb[0].uw     # NavierStokes.ux / uz (FastStepper properties)

    def _tables(self):
        ns = self.ns
        self.tbc_cheby = C.to_dev(ns.Tbc_cheby).contiguous()
        self.dTbcdz2 = C.to_dev(ns.dTbcdz2).contiguous()
        d1 = C.to_dev(ns.dTbcdz1)
        self.dTbcdz1 = torch.zeros((self.D0, self.D1 + (self.D1 & 1)), dtype=torch.float64, device=self.dev)[:, : self.D1]
        self.dTbcdz1.copy_(d1)      # same pitch as the physical-space arrays
        pp = ns.solver_P.plan_for_lhs[0]
        self.ptab = PS.PoissonTables(pp._plan, PS.lg_for(self.M0))      # grid-only: shared by all members

    # ------------------------------------------------------------------ batched helpers
    def _ptr_array(self, calls, tensors):
        t = torch.tensor([x.data_ptr() for x in tensors], dtype=torch.int64, device=self.dev)
        calls.keep += [t] + list(tensors)
        return ctypes.c_void_p(t.data_ptr())

    def _dct_members(self, calls, plan, mode, axis, pick_x, pick_y):
        """one axis of a 2-D transform for all members: pick_x / pick_y map a member to its list of arrays"""
        if len(self.runs) == 1 and plan.algo != 1:
            return self._dct(calls, plan, mode, axis, pick_x(self.mb[0]), pick_y(self.mb[0]))
        # dense-matrix transforms (small grids): all arrays of all members in ONE batched GEMM launch
        xs = [x for m in self.mb for x in pick_x(m)]
        ys = [y for m in self.mb for y in pick_y(m)]
        x, y = xs[0], ys[0]
        if plan.algo != 1:          # FFT path (large grids): 8 arrays per launch
            for k in range(0, len(xs), 8):
                self._dct(calls, plan, mode, axis, xs[k:k + 8], ys[k:k + 8])
            return
        aligned = all(t.data_ptr() % 16 == 0 for t in xs + ys)
        calls.add(C.lib().pde_dct1_batched, plan.handle, mode, len(xs), self._ptr_array(calls, xs), _ld(x), x.shape[axis],
                  self._ptr_array(calls, ys), _ld(y), y.shape[axis], x.shape[1 - axis], axis, int(aligned))

    def _gemm_members(self, calls, pick_a, Bmat, pick_c, m, n, k):
        if len(self.runs) == 1:
            a, c = pick_a(self.mb[0]), pick_c(self.mb[0])
            return calls.add(C.lib().pde_gemm_f64, 1, _ptr(a), _ld(a), _ptr(Bmat), _ld(Bmat), _ptr(c), _ld(c), m, n, k)
        As, Cs = [pick_a(mm) for mm in self.mb], [pick_c(mm) for mm in self.mb]
        aligned = all(t.data_ptr() % 16 == 0 for t in As + Cs)
        calls.add(C.lib().pde_gemm_f64_batched, 1, None, self._ptr_array(calls, As), _ld(As[0]), _ptr(Bmat), None, _ld(Bmat),
                  self._ptr_array(calls, Cs), _ld(Cs[0]), m, n, k, len(As), int(aligned))

    # ------------------------------------------------------------------ the stage
    def _state_ptrs(self):
        return tuple(t.data_ptr() for r in self.runs for t in (r.T.vhat, r.U.vhat, r.V.vhat, r.P.vhat, r.pres.vhat))

    def bind(self):
        self.bound = self._state_ptrs()
        self.stage_calls = [self._build_stage(rk) for rk in range(self.ns.nstage)]

    def stage(self, rk):
        if self._state_ptrs() != self.bound:
            self.bind()
        self.stage_calls[rk].run()

    def _build_stage(self, rk, *unused):
        ns, Lb = self.ns, C.lib()
        calls = _Calls()
        N0, N1, M0, M1 = self.N0, self.N1, self.M0, self.M1
        sx, sz = ns.scale
        dt, a, b, c = float(ns.dt), float(ns.a[rk]), float(ns.b[rk]), float(ns.c[rk])
        names = ("U", "V", "T")
        pairs = list(zip(self.runs, self.mb))
        npass = [0]

        def add(L, label=None):
            L.finalize()
            calls.keep.append(L)
            fn, args = L.args()
            npass[0] += 1
            calls.add(fn, *args, label=label or "pass[P%s%d]" % ("Y" if L.layout else "X", npass[0]))

        def objs(r):
            fld = {"U": r.U, "V": r.V, "T": r.T}
            return ({k: fld[k].vhat for k in names}, {k: fld[k].xs[0] for k in names}, {k: fld[k].xs[1] for k in names},
                    {"U": r.solver_U[rk], "V": r.solver_V[rk], "T": r.solver_T[rk]})

        def arr(m):
            return (dict(zip(names, m.c3)), dict(zip(names, m.d3)), dict(zip(names, m.e3)), dict(zip(names, m.f3)),
                    dict(zip(names, m.g3)), dict(zip(names, m.rest3)), dict(zip(names, m.z3)), dict(zip(names, m.conv)))

        # ---- PX1
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            cF, dF = arr(m)[:2]
            for k in names:
                L.job(M1).load(F[k]).stencil(xb[k]).store(cF[k]).diff(sx).store(dF[k])
            L.job(N1).load(r.pres.vhat).diff(sx).store(m.dpdx)
        add(L)
        # ---- PY2: y stencils / derivatives of the 8 arrays to transform, and everything of the right-hand sides that
        # does not need the convective term (row passes run at ~1.6x the bandwidth of column passes:
        # tools/bench_pass.py): rest_F = Sy Sx F + explicit terms
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            cF, dF, eF, fF, gF, rest = arr(m)[:6]
            for k in names:
                p = L.job(N0).load(cF[k]).stencil(yb[k])
                if k != "T":
                    p.store(eF[k])
                p.diff(sz).store(gF[k])
                L.job(N0).load(dF[k]).stencil(yb[k]).store(fF[k])
            st = {k: self.tables.stencil_elem(yb[k]) for k in names}
            L.job(N0).load(cF["U"]).stencil(yb["U"]).axpy(-dt * a, m.dpdx).store(rest["U"])
            L.job(N0).load(r.pres.vhat).diff(sz).scale(-dt * a).axpy(1.0, cF["V"], stencil=st["V"]) \
                .axpy(dt * a, cF["T"], stencil=st["T"]).axpy(dt * a, self.tbc_cheby).store(rest["V"])
            L.job(N0).load(cF["T"]).stencil(yb["T"]).axpy(dt * a * r.kappa, self.dTbcdz2).store(rest["T"])
        add(L)

        # ---- transforms and products (both convective terms of the stage merged: ub = b u + c u_old)
        def src8(m):
            e, f, g = m.e3, m.f3, m.g3
            return [e[0], e[1], f[0], f[1], f[2], g[0], g[1], g[2]]

        def dst8(m):
            new = m.uw[rk % 2]
            return [new[0], new[1]] + list(m.phys)
        self._dct_members(calls, self.plan0, ops.BWD, 0, src8, lambda m: m.X8)
        self._dct_members(calls, self.plan1, ops.BWD, 1, lambda m: m.X8, dst8)
        use_old = c != 0.0
        m0 = self.mb[0]
        new0, old0 = m0.uw[rk % 2], m0.uw[(rk + 1) % 2]
        ldp = self.D1 + (self.D1 & 1)
        for t in list(new0) + list(old0) + list(m0.phys) + [self.dTbcdz1]:
            assert tuple(t.shape) == (self.D0, self.D1) and t.stride() == (ldp, 1)
        dxU, dxV, dxT, dzU, dzV, dzT = m0.phys
        pargs = (_ptr(new0[0]), _ptr(new0[1]), _ptr(old0[0]) if use_old else None, _ptr(old0[1]) if use_old else None,
                 _ptr(dxU), _ptr(dzU), _ptr(dxV), _ptr(dzV), _ptr(dxT), _ptr(dzT), _ptr(self.dTbcdz1))
        if len(self.runs) == 1:
            calls.add(Lb.pde_conv_products, self.D0 * ldp, b, c, *pargs)
        else:   # member m's arrays lie D0 * ldp elements after member m-1's (one allocation per array kind)
            calls.add(Lb.pde_conv_products_members, self.D0 * ldp, len(self.runs), self.D0 * ldp, b, c, *pargs)
        self._dct_members(calls, self.plan1, ops.FWD, 1, lambda m: list(m.phys[:3]), lambda m: [f[:, : N1] for f in m.F3])
        self._dct_members(calls, self.plan0, ops.FWD, 0, lambda m: [f[:, : N1] for f in m.F3],
                          lambda m: [cv[: N0] for cv in m.conv])
        # ---- PX3: z = Ax^-1 Bx (rest - dt conv)
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            rest, zF, conv = arr(m)[5:]
            for k in names:
                L.job(N1).lincomb([(1.0, rest[k]), (-dt, conv[k])]).band(solver[k].plan_for_rhs[0].band) \
                    .fdma(solver[k].plan_for_lhs[0]).store(zF[k])
        add(L)
        # ---- PY4: F* = Ay^-1 By z; y parts of the divergence
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            zF = arr(m)[6]
            for k in names:
                p = L.job(M0).load(zF[k]).band(solver[k].plan_for_rhs[1].band).fdma(solver[k].plan_for_lhs[1]).store(F[k])
                if k == "U":
                    p.stencil(yb["U"]).store(m.aU[:M0])
                elif k == "V":
                    p.stencil(yb["V"]).diff(sz).store(m.aV[:M0])
        add(L)
        # ---- PX5: div = dx Sx (Sy U) / sx + Sx (dz Sy V / sz); q = Bx div
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            p = L.job(N1).load(m.aU[:M0]).stencil(xb["U"]).diff(sx)
            p.axpy(1.0, m.aV, stencil=self.tables.stencil_elem(xb["V"])).store(m.div)
            p.band(r.solver_P.plan_for_rhs[0].band).store(m.q)
        add(L)
        # ---- pressure Poisson solve (eigen-decomposition along y)
        sp = ns.solver_P
        Hy, Qy = sp.plan_for_rhs[1].dense, sp.plan_for_lhs[1].dense
        self._gemm_members(calls, lambda m: m.q, Hy, lambda m: m.R, M0, M1, N1)
        L = PS.PassLaunch(PS.COL, M0, self.tables)
        for r, m in pairs:
            L.job(M1).load(m.R).poisson(self.ptab).store(m.R)
        add(L)
        Pm = {id(m): r.P.vhat for r, m in pairs}
        self._gemm_members(calls, lambda m: m.R, Qy, lambda m: Pm[id(m)], M0, M1, M1)
        # ---- PY7: P[0,0] = 0; e1 = Sy P; bU = Gy e1; bV = Gy dz e1 / sz
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            P, ybP = r.P.vhat, r.P.xs[1]
            L.job(M0).load(P).setz0(0).store(P, only_seq=0).stencil(ybP).store(m.e1[:M0]).diff(sz) \
                .from_cheb(yb["V"]).store(m.bV[:M0])
            L.job(M0).load(P).setz0(0).stencil(ybP).from_cheb(yb["U"]).store(m.bU[:M0])
        add(L)
        # ---- PX8: velocity projection and pressure update
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for r, m in pairs:
            F, xb, yb, solver = objs(r)
            xbP, pres = r.P.xs[0], r.pres.vhat
            L.job(M1).load(m.bU[:M0]).stencil(xbP).diff(sx).from_cheb(xb["U"]).axpy(1.0, F["U"], scale_buf=-1.0).store(F["U"])
            L.job(M1).load(m.bV[:M0]).stencil(xbP).from_cheb(xb["V"]).axpy(1.0, F["V"], scale_buf=-1.0).store(F["V"])
            L.job(N1).load(m.e1[:M0]).stencil(xbP).scale(1.0 / (dt * a)) \
                .lincomb([(1.0, pres), (-(1.0 * r.nu), m.div)], accumulate=True).store(pres)
        add(L)
        calls.keep += [t for r in self.runs for t in (r.T.vhat, r.U.vhat, r.V.vhat, r.P.vhat, r.pres.vhat)]
        return calls
