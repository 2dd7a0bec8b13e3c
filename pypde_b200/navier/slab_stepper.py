"""
Slab-decomposed IMEX stage (one process per GPU, NCCL all-to-all transposes): the same batched
sm_100a primitives as fast_stepper.FastStepper, applied to local slabs, with the commuting
axis operators regrouped so that a stage needs 10 distributed transposes (navier/slab.py).

    X layout: all rows, my block of columns  -> axis-0 operators (x stencils, x derivative, x-DCT,
              Helmholtz x sweeps, per-column Poisson solves, x part of the projection)
    Y layout: my block of rows, all columns  -> axis-1 operators (y stencils / derivative, y-DCT,
              products, Helmholtz y sweeps, the dense Hy / Qy projections, y part of the projection)

Regrouping used (all exact identities of commuting tensor-product operators; rounding-level
differences only, asserted <= 1e-12 against the CPU oracle in tests/dist_slab_check.py):
    Ay^-1 Ax^-1 [By Bx rhs + (By Sy)(Bx Sx) F]  =  Ay^-1 [ By (Ax^-1 Bx rhs) + (By Sy)(Ax^-1 Bx Sx F) ]
    div                                         =  dx Sx (Sy U)/sx + Sx (dy Sy V)/sz
    U -= Gy Gx dx Sx Sy P / sx                  =  Gx dx Sx (Gy Sy P) / sx        (same for V)
"""
import torch
import torch.distributed as dist

from .. import _cabi as C
from .. import ops
from .fast_stepper import FastStepper, _Calls, _ptr, _ld, FDMA_FWD, FDMA_BWD, TDMA_FWD, TDMA_BWD
from .slab import SlabComm, partition


class SlabStepper(FastStepper):
    def __init__(self, ns, group=None):
        self.comm = SlabComm(group)
        self.P, self.r = self.comm.size, self.comm.rank
        FastStepper.__init__(self, ns)

    # ------------------------------------------------------------------ layout helpers
    def _new(self, *shape):
        # bundles are exchanged as flat buffers: always contiguous (no pitch padding)
        return torch.zeros(shape, dtype=torch.float64, device=self.dev)

    def _alloc(self):
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        P, r = self.P, self.r
        self.rp, self.cp, self.dp = partition(N0, P), partition(N1, P), partition(D0, P)
        self.r0, self.N0r = self.rp[r]
        self.c0, self.N1c = self.cp[r]
        self.d0, self.D0r = self.dp[r]
        self.M1c = max(0, min(self.c0 + self.N1c, M1) - self.c0)      # my valid Galerkin columns
        self.M0r = max(0, min(self.r0 + self.N0r, M0) - self.r0)      # my valid Galerkin rows
        if min(self.M1c, self.M0r, self.D0r) <= 0:
            raise ValueError("grid too small for %d slabs" % P)
        n = self._new
        W, Wy = self.N1c, N1
        self.A_X, self.A_Y = n(N0, 10 * W), n(self.N0r, 10 * Wy)
        self.B_Y, self.B_X = n(self.N0r, 10 * Wy), n(N0, 10 * W)
        self.C_X, self.C_Y = n(D0, 8 * W), n(self.D0r, 8 * Wy)
        self.F_Y, self.F_X = n(self.D0r, 3 * Wy), n(D0, 3 * W)
        self.G_X, self.G_Y = n(N0, 3 * W), n(self.N0r, 3 * Wy)
        self.S_Y, self.S_X = n(self.N0r, 5 * Wy), n(N0, 5 * W)
        self.Q_X, self.Q_Y = n(N0, W), n(self.N0r, Wy)
        self.R_Y, self.R_X = n(self.N0r, Wy), n(N0, W)
        self.W_Y, self.P_Y = n(self.N0r, Wy), n(self.N0r, Wy)
        self.Z_Y, self.Z_X = n(self.N0r, 3 * Wy), n(N0, 3 * W)
        self.phys = [n(self.D0r, D1) for _ in range(6)]
        self.uw = [[n(self.D0r, D1), n(self.D0r, D1)] for _ in range(2)]
        self.conv = [n(N0, W) for _ in range(3)]
        self.rhs = [n(N0, W) for _ in range(3)]
        self.dpdx, self.divx = n(N0, W), n(N0, W)
        self.tx = [n(N0, W) for _ in range(4)]
        self.ty = [n(self.N0r, Wy) for _ in range(3)]

    def XV(self, b, k, nrows, ncols):
        return b[:nrows, k * self.N1c: k * self.N1c + ncols]

    def YV(self, b, k, nrows, ncols):
        return b[:nrows, k * self.N1: k * self.N1 + ncols]

    def _tables(self):
        FastStepper._tables(self)
        ns = self.ns
        self.tbc_Y = self.tbc_cheby[self.r0:self.r0 + self.N0r].contiguous()
        self.dTbcdz2_X = self.dTbcdz2[:, self.c0:self.c0 + self.N1c].contiguous()
        self.dTbcdz1_Y = self.dTbcdz1[self.d0:self.d0 + self.D0r].contiguous()
        pp = ns.solver_P.plan_for_lhs[0]
        self.poisson_local = ops.PoissonPlan(pp._Ad, pp._Cd, pp.alpha[self.c0:self.c0 + self.M1c], pp.singular)

    # ------------------------------------------------------------------ state movement
    def scatter(self):
        """Global fields (replicated on every rank) -> local X-layout slabs."""
        ns, c0 = self.ns, self.c0
        for k, f in enumerate((ns.T, ns.U, ns.V)):
            self.XV(self.S_X, k, self.M0, self.M1c).copy_(f.vhat[:, c0:c0 + self.M1c])
        self.XV(self.A_X, 9, self.N0, self.N1c).copy_(ns.pres.vhat[:, c0:c0 + self.N1c])

    def gather(self):
        """Local slabs -> global fields on every rank (diagnostics, parity tests, I/O)."""
        ns = self.ns
        wmax = max(w for _, w in self.cp)
        for k, f in enumerate((ns.T, ns.U, ns.V, ns.pres)):
            rows = self.M0 if k < 3 else self.N0
            cols_total = self.M1 if k < 3 else self.N1
            loc = torch.zeros((rows, wmax), dtype=torch.float64, device=self.dev)
            src = self.XV(self.S_X, k, rows, self.M1c) if k < 3 else self.XV(self.A_X, 9, rows, self.N1c)
            loc[:, : src.shape[1]].copy_(src)
            parts = [torch.empty_like(loc) for _ in range(self.P)]
            dist.all_gather(parts, loc, group=self.comm.group)
            for s, (o, w) in enumerate(self.cp):
                w = max(0, min(o + w, cols_total) - o)
                if w:
                    f.vhat[:, o:o + w].copy_(parts[s][:, :w])

    # ------------------------------------------------------------------ the stage
    def bind(self):
        self.scatter()
        self.bound = True
        self.stage_calls = [self._build_stage(rk) for rk in range(self.ns.nstage)]

    def _x2y(self, calls, xb, yb, K, rows):
        def go(_stream):
            self.comm.x2y(xb, yb, K, rows, self.N1)
            return 0
        go.__name__ = "slab_transpose_x2y"
        calls.add(go)

    def _y2x(self, calls, yb, xb, K, rows):
        def go(_stream):
            self.comm.y2x(yb, xb, K, rows, self.N1)
            return 0
        go.__name__ = "slab_transpose_y2x"
        calls.add(go)

    def _solve_axis(self, calls, rk, ax, fields, xs):
        """in-place 4-diagonal solves along one axis for several fields"""
        n = xs[0].shape[ax]
        fw, bw = [], []
        for name, x in zip(fields, xs):
            l, d, u1, u2, rd = self.lu[(rk, name, ax)]
            fw.append(dict(**{"in": [x]}, out=x, tab={0: l}))
            bw.append(dict(**{"in": [x]}, out=x, tab={1: d, 2: u1, 3: u2, 4: rd}))
        self._sweep(calls, FDMA_FWD, ax, n, fw)
        self._sweep(calls, FDMA_BWD, ax, n, bw)

    def _build_stage(self, rk):
        ns, L = self.ns, C.lib()
        calls = _Calls()
        XV, YV = self.XV, self.YV
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        N0r, N1c, M0r, M1c, D0r = self.N0r, self.N1c, self.M0r, self.M1c, self.D0r
        sx, sz = ns.scale
        dt, a, b, c = float(ns.dt), float(ns.a[rk]), float(ns.b[rk]), float(ns.c[rk])
        names = ("U", "V", "T")
        slot = {"T": 0, "U": 1, "V": 2}                    # state slots in S_X / S_Y
        solvers = {"U": ns.solver_U[rk], "V": ns.solver_V[rk], "T": ns.solver_T[rk]}
        state = {f: XV(self.S_X, slot[f], M0, M1c) for f in names}
        pres = XV(self.A_X, 9, N0, N1c)

        # ---- X1: x stencils / derivatives of the state, "old" Helmholtz term with its x solve
        cs = {f: XV(self.A_X, k, N0, M1c) for k, f in enumerate(names)}
        ds = {f: XV(self.A_X, 3 + k, N0, M1c) for k, f in enumerate(names)}
        hs = {f: XV(self.A_X, 6 + k, M0, M1c) for k, f in enumerate(names)}
        self._stencil(calls, 0, [(self.sx[f], state[f], cs[f]) for f in names])
        self._diff(calls, 0, [(cs[f], ds[f]) for f in names], sx)
        self._band(calls, 0, [(solvers[f].plan_for_old[0].band, state[f], hs[f]) for f in names])
        self._solve_axis(calls, rk, 0, names, [hs[f] for f in names])
        self._x2y(calls, self.A_X, self.A_Y, 10, N0)
        # ---- Y2: y stencils / derivatives, buoyancy term, dp/dz
        eU, eV = YV(self.B_Y, 0, N0r, N1), YV(self.B_Y, 1, N0r, N1)
        fU, fV, fT = (YV(self.B_Y, k, N0r, N1) for k in (2, 3, 4))
        gU, gV, gT = (YV(self.B_Y, k, N0r, N1) for k in (5, 6, 7))
        thc, dpdz = YV(self.B_Y, 8, N0r, N1), YV(self.B_Y, 9, N0r, N1)
        eT = self.ty[0]
        ay = lambda k: YV(self.A_Y, k, N0r, M1)
        self._stencil(calls, 1, [(self.sy["U"], ay(0), eU), (self.sy["V"], ay(1), eV), (self.sy["T"], ay(2), eT),
                                 (self.sy["U"], ay(3), fU), (self.sy["V"], ay(4), fV), (self.sy["T"], ay(5), fT)])
        self._diff(calls, 1, [(eU, gU), (eV, gV), (eT, gT), (YV(self.A_Y, 9, N0r, N1), dpdz)], sz)
        self._lincomb(calls, [(thc, [(1.0, eT), (1.0, self.tbc_Y)])])
        self._y2x(calls, self.B_Y, self.B_X, 10, N0)
        # ---- X3: backward x-DCT of the 8 coefficient arrays
        self._dct(calls, self.plan0, ops.BWD, 0, [XV(self.B_X, k, N0, N1c) for k in range(8)],
                  [XV(self.C_X, k, D0, N1c) for k in range(8)])
        self._x2y(calls, self.C_X, self.C_Y, 8, D0)
        # ---- Y4: backward y-DCT, products, forward y-DCT
        new, old = self.uw[rk % 2], self.uw[(rk + 1) % 2]
        dxU, dxV, dxT, dzU, dzV, dzT = self.phys
        self._dct(calls, self.plan1, ops.BWD, 1, [YV(self.C_Y, k, D0r, N1) for k in range(8)],
                  [new[0], new[1], dxU, dxV, dxT, dzU, dzV, dzT])
        use_old = c != 0.0
        calls.add(L.pde_conv_products, D0r * D1, b, c, _ptr(new[0]), _ptr(new[1]),
                  _ptr(old[0]) if use_old else None, _ptr(old[1]) if use_old else None,
                  _ptr(dxU), _ptr(dzU), _ptr(dxV), _ptr(dzV), _ptr(dxT), _ptr(dzT), _ptr(self.dTbcdz1_Y))
        self._dct(calls, self.plan1, ops.FWD, 1, [dxU, dxV, dxT], [YV(self.F_Y, k, D0r, N1) for k in range(3)])
        self._y2x(calls, self.F_Y, self.F_X, 3, D0)
        # ---- X5: forward x-DCT, right-hand sides, Bx and the x solves
        self._dct(calls, self.plan0, ops.FWD, 0, [XV(self.F_X, k, D0, N1c) for k in range(3)], self.conv)
        self._diff(calls, 0, [(pres, self.dpdx)], sx)
        rU, rV, rT = self.rhs
        self._lincomb(calls, [
            (rU, [(-dt * a, self.dpdx), (-dt, self.conv[0])]),
            (rV, [(-dt * a, XV(self.B_X, 9, N0, N1c)), (-dt, self.conv[1]), (dt * a, XV(self.B_X, 8, N0, N1c))]),
            (rT, [(-dt, self.conv[2]), (dt * a * ns.kappa, self.dTbcdz2_X)]),
        ])
        gs = {f: XV(self.G_X, slot[f], M0, N1c) for f in names}
        rr = {"U": rU, "V": rV, "T": rT}
        self._band(calls, 0, [(solvers[f].plan_for_rhs[0].band, rr[f], gs[f]) for f in names])
        self._solve_axis(calls, rk, 0, names, [gs[f] for f in names])
        self._x2y(calls, self.G_X, self.G_Y, 3, N0)
        # ---- Y6: By, (By Sy), y solves -> new fields; y parts of the divergence
        gy = {f: YV(self.G_Y, slot[f], M0r, N1) for f in names}
        hy = {f: YV(self.A_Y, 6 + k, M0r, M1) for k, f in enumerate(names)}
        ry = {f: YV(self.S_Y, slot[f], M0r, M1) for f in names}
        self._band(calls, 1, [(solvers[f].plan_for_rhs[1].band, gy[f], ry[f]) for f in names])
        self._band(calls, 1, [(solvers[f].plan_for_old[1].band, hy[f], ry[f]) for f in names], accumulate=True)
        self._solve_axis(calls, rk, 1, names, [ry[f] for f in names])
        aU, aV = YV(self.S_Y, 3, M0r, N1), YV(self.S_Y, 4, M0r, N1)
        tV = self.ty[1][:M0r]
        self._stencil(calls, 1, [(self.sy["U"], ry["U"], aU), (self.sy["V"], ry["V"], tV)])
        self._diff(calls, 1, [(tV, aV)], sz)
        self._y2x(calls, self.S_Y, self.S_X, 5, N0)
        # ---- X7: divergence, Bx
        t1, t2, t3, t4 = self.tx
        self._stencil(calls, 0, [(self.sx["U"], XV(self.S_X, 3, M0, N1c), t1), (self.sx["V"], XV(self.S_X, 4, M0, N1c), t3)])
        self._diff(calls, 0, [(t1, t2)], sx)
        self._lincomb(calls, [(self.divx, [(1.0, t2), (1.0, t3)])])
        sp = ns.solver_P
        self._band(calls, 0, [(sp.plan_for_rhs[0].band, self.divx, XV(self.Q_X, 0, M0, N1c))])
        self._x2y(calls, self.Q_X, self.Q_Y, 1, N0)
        # ---- Y8 / X9 / Y10: Hy projection, per-column Poisson sweeps, Qy projection
        Hy, Qy = sp.plan_for_rhs[1].dense, sp.plan_for_lhs[1].dense
        calls.add(L.pde_gemm_f64, 1, _ptr(self.Q_Y), _ld(self.Q_Y), _ptr(Hy), _ld(Hy), _ptr(self.R_Y), _ld(self.R_Y),
                  M0r, M1, N1)
        self._y2x(calls, self.R_Y, self.R_X, 1, N0)
        Rx = XV(self.R_X, 0, M0, M1c)
        calls.add(L.pde_poisson_solve, self.poisson_local.handle, _ptr(Rx), _ld(Rx))
        self._x2y(calls, self.R_X, self.W_Y, 1, N0)
        calls.add(L.pde_gemm_f64, 1, _ptr(self.W_Y), _ld(self.W_Y), _ptr(Qy), _ld(Qy), _ptr(self.P_Y), _ld(self.P_Y),
                  M0r, M1, M1)
        if self.r0 == 0:
            zero = self._new(1, 1)
            calls.keep.append(zero)
            self._lincomb(calls, [(self.P_Y[0:1, 0:1], [(1.0, zero)])])
        Py = YV(self.P_Y, 0, M0r, M1)
        bU, bV, e1 = YV(self.Z_Y, 0, M0r, M1), YV(self.Z_Y, 1, M0r, M1), YV(self.Z_Y, 2, M0r, N1)
        e2 = self.ty[2][:M0r]
        self._stencil(calls, 1, [(self.sy["P"], Py, e1)])
        self._diff(calls, 1, [(e1, e2)], sz)
        self._from_cheb(calls, 1, ("U", "V"), [e1, e2], [bU, bV])
        self._y2x(calls, self.Z_Y, self.Z_X, 3, N0)
        # ---- X11: x part of the projection, pressure update
        bUx, bVx, e1x = XV(self.Z_X, 0, M0, M1c), XV(self.Z_X, 1, M0, M1c), XV(self.Z_X, 2, M0, N1c)
        u1, u2, u3, u4 = (t[:, :M1c] for t in self.tx)
        self._stencil(calls, 0, [(self.sx["P"], bUx, u1), (self.sx["P"], bVx, u3)])
        self._diff(calls, 0, [(u1, u2)], sx)
        cU, cV = self.rhs[0][:M0, :M1c], self.rhs[1][:M0, :M1c]
        self._from_cheb(calls, 0, ("U", "V"), [u2, u3], [cU, cV])
        self._lincomb(calls, [(state["U"], [(1.0, state["U"]), (-1.0, cU)]), (state["V"], [(1.0, state["V"]), (-1.0, cV)])])
        pe = self.conv[0]
        self._stencil(calls, 0, [(self.sx["P"], e1x, pe)])
        self._lincomb(calls, [(pres, [(1.0, pres), (-(1.0 * ns.nu), self.divx), (1.0 / (dt * a), pe)])])
        return calls

    def stage(self, rk):
        if not self.bound:
            self.bind()
        self.stage_calls[rk].run()

    def local_state(self):
        """The tensors that hold this rank's part of the state (T, U, V slabs; pres slab)."""
        return [self.XV(self.S_X, k, self.M0, self.M1c) for k in range(3)] + [self.XV(self.A_X, 9, self.N0, self.N1c)]
