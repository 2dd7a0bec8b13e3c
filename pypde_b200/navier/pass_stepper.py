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
"""
import torch

from .. import _cabi as C
from .. import ops
from .. import passes as PS
from .fast_stepper import FastStepper, _Calls, _ptr, _ld


class PassStepper(FastStepper):
    @staticmethod
    def supported(ns):
        N0, N1 = ns.shape
        return ns.beta == 1.0 and N0 % 2 == 0 and N1 % 2 == 0 and 8 <= N0 <= 4096 and 8 <= N1 <= 4096

    def __init__(self, ns):
        self.tables = PS.TableCache()
        FastStepper.__init__(self, ns)

    def _new(self, *shape):
        """Row passes move 16-byte units: even leading dimensions, 16-byte aligned rows (torch allocations are
        256-byte aligned).  Rows whose byte length is a multiple of 2 KB are padded by 64 bytes like in
        FastStepper (column walks of the x-transforms)."""
        if len(shape) == 2 and shape[1] % 256 == 0 and shape[0] > 1:
            return torch.zeros((shape[0], shape[1] + 8), dtype=torch.float64, device=self.dev)[:, : shape[1]]
        assert len(shape) != 2 or shape[1] % 2 == 0 or shape[0] == 1
        return torch.zeros(shape, dtype=torch.float64, device=self.dev)

    def _alloc(self):
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        n = self._new
        self.c3 = [n(N0, M1) for _ in range(3)]
        self.d3 = [n(N0, M1) for _ in range(3)]
        self.e3 = [n(N0, N1) for _ in range(3)]
        self.f3 = [n(N0, N1) for _ in range(3)]
        self.g3 = [n(N0, N1) for _ in range(3)]
        self.dpdx = n(N0, N1)
        self.rest3 = [n(N0, N1) for _ in range(3)]
        # transforms: D x D arrays may have odd widths (3073): plain allocations, only the DCT kernels touch them
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=self.dev)
        self.X8 = [FastStepper._new(self, D0, N1) for _ in range(8)]
        self.phys = [z(D0, D1) for _ in range(6)]
        self.uw = [[z(D0, D1), z(D0, D1)] for _ in range(2)]
        self.F3 = [FastStepper._new(self, D0, N1) for _ in range(3)]
        self.conv = [n(N0, N1) for _ in range(3)]
        self.z3 = [n(M0, N1) for _ in range(3)]
        self.aU, self.aV = n(N0, N1), n(N0, N1)          # M0 rows used; the rest stays zero (stencil-on-load)
        self.div, self.q, self.R = n(N0, N1), n(M0, N1), n(M0, M1)
        self.e1, self.bU, self.bV = n(N0, N1), n(N0, M1), n(N0, M1)     # M0 rows used

    def _tables(self):
        FastStepper._tables(self)
        pp = self.ns.solver_P.plan_for_lhs[0]
        self.ptab = PS.PoissonTables(pp._plan, PS.lg_for(self.M0))

    # ------------------------------------------------------------------ the stage
    def _build_stage(self, rk, T, U, V, P, pres):
        ns, Lb = self.ns, C.lib()
        calls = _Calls()
        N0, N1, M0, M1 = self.N0, self.N1, self.M0, self.M1
        sx, sz = ns.scale
        dt, a, b, c = float(ns.dt), float(ns.a[rk]), float(ns.b[rk]), float(ns.c[rk])
        F = {"U": U, "V": V, "T": T}
        fld = {"U": ns.U, "V": ns.V, "T": ns.T}
        names = ("U", "V", "T")
        cF, dF = dict(zip(names, self.c3)), dict(zip(names, self.d3))
        eF, fF, gF = dict(zip(names, self.e3)), dict(zip(names, self.f3)), dict(zip(names, self.g3))
        zF, conv = dict(zip(names, self.z3)), dict(zip(names, self.conv))
        rest = dict(zip(names, self.rest3))
        xb = {k: fld[k].xs[0] for k in names}
        yb = {k: fld[k].xs[1] for k in names}
        xbP, ybP = ns.P.xs[0], ns.P.xs[1]
        solver = {"U": ns.solver_U[rk], "V": ns.solver_V[rk], "T": ns.solver_T[rk]}

        npass = [0]

        def add(L, label=None):
            L.finalize()
            calls.keep.append(L)
            fn, args = L.args()
            npass[0] += 1
            calls.add(fn, *args, label=label or "pass[P%s%d]" % ("Y" if L.layout else "X", npass[0]))

        # ---- PX1
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for k in names:
            L.job(M1).load(F[k]).stencil(xb[k]).store(cF[k]).diff(sx).store(dF[k])
        L.job(N1).load(pres).diff(sx).store(self.dpdx)
        add(L)
        # ---- PY2: y stencils / derivatives of the 8 arrays to transform, and everything of the right-hand sides that
        # does not need the convective term (row passes run at ~1.6x the bandwidth of column passes:
        # tools/bench_pass.py): rest_F = Sy Sx F + explicit terms
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in names:
            p = L.job(N0).load(cF[k]).stencil(yb[k])
            if k != "T":
                p.store(eF[k])
            p.diff(sz).store(gF[k])
            L.job(N0).load(dF[k]).stencil(yb[k]).store(fF[k])
        st = {k: self.tables.stencil_elem(yb[k]) for k in names}
        L.job(N0).load(cF["U"]).stencil(yb["U"]).axpy(-dt * a, self.dpdx).store(rest["U"])
        L.job(N0).load(pres).diff(sz).scale(-dt * a).axpy(1.0, cF["V"], stencil=st["V"]) \
            .axpy(dt * a, cF["T"], stencil=st["T"]).axpy(dt * a, self.tbc_cheby).store(rest["V"])
        L.job(N0).load(cF["T"]).stencil(yb["T"]).axpy(dt * a * ns.kappa, self.dTbcdz2).store(rest["T"])
        add(L)
        # ---- transforms and products (both convective terms of the stage merged: ub = b u + c u_old)
        new, old = self.uw[rk % 2], self.uw[(rk + 1) % 2]
        dxU, dxV, dxT, dzU, dzV, dzT = self.phys
        src = [eF["U"], eF["V"], fF["U"], fF["V"], fF["T"], gF["U"], gF["V"], gF["T"]]
        dst = [new[0], new[1], dxU, dxV, dxT, dzU, dzV, dzT]
        self._dct(calls, self.plan0, ops.BWD, 0, src, self.X8)
        self._dct(calls, self.plan1, ops.BWD, 1, self.X8, dst)
        use_old = c != 0.0
        for t in list(new) + list(old) + list(self.phys) + [self.dTbcdz1]:
            assert t.is_contiguous() and tuple(t.shape) == (self.D0, self.D1)
        calls.add(Lb.pde_conv_products, self.D0 * self.D1, b, c, _ptr(new[0]), _ptr(new[1]),
                  _ptr(old[0]) if use_old else None, _ptr(old[1]) if use_old else None,
                  _ptr(dxU), _ptr(dzU), _ptr(dxV), _ptr(dzV), _ptr(dxT), _ptr(dzT), _ptr(self.dTbcdz1))
        self._dct(calls, self.plan1, ops.FWD, 1, [dxU, dxV, dxT], [f[:, : N1] for f in self.F3])
        self._dct(calls, self.plan0, ops.FWD, 0, [f[:, : N1] for f in self.F3], [cv[: N0] for cv in self.conv])
        # ---- PX3: z = Ax^-1 Bx (rest - dt conv)
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for k in names:
            L.job(N1).lincomb([(1.0, rest[k]), (-dt, conv[k])]).band(solver[k].plan_for_rhs[0].band) \
                .fdma(solver[k].plan_for_lhs[0]).store(zF[k])
        add(L)
        # ---- PY4: F* = Ay^-1 By z; y parts of the divergence
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in names:
            p = L.job(M0).load(zF[k]).band(solver[k].plan_for_rhs[1].band).fdma(solver[k].plan_for_lhs[1]).store(F[k])
            if k == "U":
                p.stencil(yb["U"]).store(self.aU[:M0])
            elif k == "V":
                p.stencil(yb["V"]).diff(sz).store(self.aV[:M0])
        add(L)
        # ---- PX5: div = dx Sx (Sy U) / sx + Sx (dz Sy V / sz); q = Bx div
        sp = ns.solver_P
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        p = L.job(N1).load(self.aU[:M0]).stencil(xb["U"]).diff(sx)
        p.axpy(1.0, self.aV, stencil=self.tables.stencil_elem(xb["V"])).store(self.div)
        p.band(sp.plan_for_rhs[0].band).store(self.q)
        add(L)
        # ---- pressure Poisson solve (eigen-decomposition along y)
        Hy, Qy = sp.plan_for_rhs[1].dense, sp.plan_for_lhs[1].dense
        calls.add(Lb.pde_gemm_f64, 1, _ptr(self.q), _ld(self.q), _ptr(Hy), _ld(Hy), _ptr(self.R), _ld(self.R),
                  M0, M1, N1)
        L = PS.PassLaunch(PS.COL, M0, self.tables)
        L.job(M1).load(self.R).poisson(self.ptab).store(self.R)
        add(L)
        calls.add(Lb.pde_gemm_f64, 1, _ptr(self.R), _ld(self.R), _ptr(Qy), _ld(Qy), _ptr(P), _ld(P), M0, M1, M1)
        # ---- PY7: P[0,0] = 0; e1 = Sy P; bU = Gy e1; bV = Gy dz e1 / sz
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        L.job(M0).load(P).setz0(0).store(P, only_seq=0).stencil(ybP).store(self.e1[:M0]).diff(sz) \
            .from_cheb(yb["V"]).store(self.bV[:M0])
        L.job(M0).load(P).setz0(0).stencil(ybP).from_cheb(yb["U"]).store(self.bU[:M0])
        add(L)
        # ---- PX8: velocity projection and pressure update
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        L.job(M1).load(self.bU[:M0]).stencil(xbP).diff(sx).from_cheb(xb["U"]).axpy(1.0, U, scale_buf=-1.0).store(U)
        L.job(M1).load(self.bV[:M0]).stencil(xbP).from_cheb(xb["V"]).axpy(1.0, V, scale_buf=-1.0).store(V)
        L.job(N1).load(self.e1[:M0]).stencil(xbP).scale(1.0 / (dt * a)) \
            .lincomb([(1.0, pres), (-(1.0 * ns.nu), self.div)], accumulate=True).store(pres)
        add(L)
        calls.keep += [T, U, V, P, pres]
        return calls
