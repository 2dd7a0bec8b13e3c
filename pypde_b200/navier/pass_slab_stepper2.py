"""
Slab-decomposed IMEX stage, y-first schedule: the 8-exchange form of SURVEY.md §8(e).

pass_slab_stepper.PassSlabStepper keeps the operator order of the single-GPU stepper (x-transform first) and needs
10 exchanges per stage.  Here every two-axis operator is applied y-first, so that the row passes, the y-transforms
and the y half of the Helmholtz problems all run on the SAME row slab without an exchange in between:

    PX1   (X) F -> Sx F, dx Sx F / sx; pres -> dpdx                      push c, d, pres, dpdx          | barrier
    PY2   (Y) y stencils / derivatives (8 rows per field set), right-hand-side parts rest_F             (local)
          (Y) 8 backward y-DCTs  (N0r x N1 -> N0r x D1)                  push 8 (N x D)                 | barrier
          (X) 8 backward x-DCTs, products, 3 forward x-DCTs              push 3 (N x D)                 | barrier
          (Y) 3 forward y-DCTs; PYh: h = Ay^-1 By (rest - dt conv)       pushed by the pass itself      | barrier
    PXh   (X) F* = Ax^-1 Bx h -> state; x parts of div(U*, V*)           push 2                         | barrier
    PYd   (Y) div; R' = div Hy^T                                         push R', div                   | barrier
    PX6   (X) q = Bx R'; per-column Poisson solves                       push W                         | barrier
          (Y) P = W Qy^T; PY7: projection parts                          pushed by the pass itself      | barrier
    PX8   (X) velocity projection, pressure update

8 barriers and 35.5 array-exchanges (in units of one N x N array) per stage instead of 10 and 47.5.
Same identities as pass_stepper.py (commuting tensor-product operators) plus Bx (div Hy^T) = (Bx div) Hy^T.
"""
import torch
import torch.distributed as dist

from .. import _cabi as C
from .. import ops
from .. import passes as PS
from .fast_stepper import FastStepper, _Calls, _ptr, _ld
from .pass_slab_stepper import PassSlabStepper
from .peer import PeerMem, SlabLayout, partition


class PassSlabStepperY(PassSlabStepper):
    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        lay = self.lay = SlabLayout(N0, N1, D0, D1, self.P, self.r)
        self.dcp = partition(D1, self.P, 4)                 # columns of the X layout of y-physical arrays
        self.dc0, self.Wd = self.dcp[self.r]
        if min(w for _, w in self.dcp) <= 0:
            raise ValueError("grid too small for %d slabs" % self.P)
        self.Wdmax = max(w for _, w in self.dcp)
        self.Wdmax += self.Wdmax & 1
        self.D1p = D1 + (D1 & 1)
        W = lay.Wmax
        off = PeerMem.HEADER
        self.xoff, self.yoff = {}, {}
        self.xld, self.yld = {}, {}

        def arr(table, ldt, name, rows, ld):
            nonlocal off
            table[name] = (off, rows)
            ldt[name] = ld
            off += ((rows * ld * 8 + 255) // 256) * 256
        N0rmax = max(n_ for _, n_ in lay.rp)
        for k in "TUV":
            for pre in ("S", "c", "d", "h"):
                arr(self.xoff, self.xld, pre + k, N0, W)
        for name in ("pres", "dpdx", "aUx", "aVx", "div", "R", "e1", "bU", "bV"):
            arr(self.xoff, self.xld, name, N0, W)
        for k in range(8):
            arr(self.xoff, self.xld, "ZX8_%d" % k, N0, self.Wdmax)
        for name in ("y_cU", "y_cV", "y_cT", "y_dU", "y_dV", "y_dT", "y_pres", "y_dpdx", "y_aUx", "y_aVx", "WY"):
            arr(self.yoff, self.yld, name, N0rmax, N1)
        for k in range(3):
            arr(self.yoff, self.yld, "F3Y_%d" % k, N0rmax, self.D1p)
        self.mem = PeerMem(off, self.group)
        self.X = {name: self.mem.view(o, (rows, self.xld[name])) for name, (o, rows) in self.xoff.items()}
        self.Yp = {name: self.mem.view(o, (rows, self.yld[name])) for name, (o, rows) in self.yoff.items()}
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=self.dev)
        N0r = lay.N0r
        self.eY = [z(N0r, N1) for _ in range(8)]                     # eU eV fU fV fT gU gV gT
        self.restY = [z(N0r, N1) for _ in range(3)]
        self.Z8 = [z(N0r, self.D1p) for _ in range(8)]
        self.phys = [z(D0, self.Wdmax) for _ in range(6)]
        self.uw = [[z(D0, self.Wdmax), z(D0, self.Wdmax)] for _ in range(2)]
        self.F3X = [z(N0, self.Wdmax) for _ in range(3)]
        self.convY = [z(N0r, N1) for _ in range(3)]
        self.divY, self.RY, self.PY = z(N0r, N1), z(N0r, M1), z(N0r, M1)

    def _tables(self):
        FastStepper._tables(self)
        ns, lay = self.ns, self.lay
        r0, N0r = lay.r0, lay.N0r
        self.tbc_Y = self.tbc_cheby[r0:r0 + N0r].contiguous()
        self.dTbcdz2_Y = self.dTbcdz2[r0:r0 + N0r].contiguous()
        t = torch.zeros((self.D0, self.Wdmax), dtype=torch.float64, device=self.dev)
        t[:, :self.Wd] = self.dTbcdz1[:, self.dc0:self.dc0 + self.Wd]
        self.dTbcdz1_X = t
        pp = ns.solver_P.plan_for_lhs[0]
        self.poisson_local = ops.PoissonPlan(pp._Ad, pp._Cd, pp.alpha[lay.c0:lay.c0 + lay.M1c], pp.singular)
        self.ptab = PS.PoissonTables(self.poisson_local, PS.lg_for(self.N0))     # PX6 runs with the N0-long launch

    # ------------------------------------------------------------------ operands
    def xr(self, name, ncols, row0):
        """rows of an X-layout array (column partition of N1) seen from the Y layout"""
        off, _ = self.xoff[name]
        ld = self.xld[name]
        ptrs = [b + off + 8 * row0 * ld for b in self.mem.base]
        return PS.Operand(ptrs, [ld] * self.P, self.lay.col_starts(ncols), ncols, keep=(self.mem,))

    def push_x2y(self, L, xname, yname, parts, ncols, colpart=None, src=None):
        """row pieces (my columns) of a local X-layout array -> the Y-layout array of the rank that owns the row"""
        colpart = colpart or self.lay.cp
        c0, w = colpart[self.r]
        w = max(0, min(c0 + w, ncols) - c0)
        yo, _ = self.yoff[yname]
        ld = self.yld[yname]
        srcarr = self.X[xname] if src is None else src
        for s, (row0, nrows) in enumerate(parts):
            if nrows <= 0 or w <= 0:
                continue
            dst = PS.Operand([self.mem.base[s] + yo + 8 * c0], [ld], [0, w], w, keep=(self.mem,))
            L.job(nrows).load(srcarr[row0:row0 + nrows, :w]).store(dst)

    def yseg(self, yname, parts, nrows_total):
        """COL-layout operand: column q of my slab, rows split over their owners -- the Y-layout array `yname` of every
        rank, at my column offset.  A column pass that stores through it writes 64-byte row pieces (8 columns) straight
        into the peers' row slabs (PDE_SLAB_COLPUSH=1; the default is a local store + a row pass that pushes 2 KB pieces)."""
        yo, _ = self.yoff[yname]
        ld = self.yld[yname]
        ptrs, starts = [], []
        for s, (row0, nrows) in enumerate(parts):
            if nrows <= 0:
                continue
            ptrs.append(self.mem.base[s] + yo + 8 * self.lay.c0)
            starts.append(row0)
        starts.append(nrows_total)
        return PS.Operand(ptrs, [ld] * len(ptrs), starts, nrows_total, keep=(self.mem,))

    def push_y2x_wide(self, L, src, xname, nrows, row0):
        """rows of a local Y-layout array with D1 columns -> the X-layout arrays (column partition dcp) of all ranks;
        rows longer than 4096 elements' worth of one pass are cut at column 2048"""
        xo, _ = self.xoff[xname]
        ld = self.xld[xname]
        D1 = self.D1
        for h0 in range(0, D1, 2048):
            h1 = min(D1, h0 + 2048)
            ptrs, starts = [], []
            for s, (c0, w) in enumerate(self.dcp):
                lo, hi = max(c0, h0), min(c0 + w, h1)
                if hi <= lo:
                    continue
                ptrs.append(self.mem.base[s] + xo + 8 * (row0 * ld + (lo - c0)))
                starts.append(lo - h0)
            starts.append(h1 - h0)
            dst = PS.Operand(ptrs, [ld] * len(ptrs), starts, h1 - h0, keep=(self.mem,))
            L.job(nrows).load(src[:nrows, h0:h1]).store(dst)

    # ------------------------------------------------------------------ state movement
    def scatter(self):
        ns, lay = self.ns, self.lay
        for k, f in (("T", ns.T), ("U", ns.U), ("V", ns.V)):
            self.xl("S" + k, self.M0, lay.M1c).copy_(f.vhat[:, lay.c0:lay.c0 + lay.M1c])
        self.xl("pres", self.N0, lay.W).copy_(ns.pres.vhat[:, lay.c0:lay.c0 + lay.W])
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    # ------------------------------------------------------------------ the stage
    def _build_stage(self, rk):
        ns, Lb, lay = self.ns, C.lib(), self.lay
        calls = _Calls()
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        W, M1c, N0r, M0r, r0 = lay.W, lay.M1c, lay.N0r, lay.M0r, lay.r0
        Wd = self.Wd
        sx, sz = ns.scale
        dt, a, b, c = float(ns.dt), float(ns.a[rk]), float(ns.b[rk]), float(ns.c[rk])
        names = ("U", "V", "T")
        fld = {"U": ns.U, "V": ns.V, "T": ns.T}
        xb = {k: fld[k].xs[0] for k in names}
        yb = {k: fld[k].xs[1] for k in names}
        xbP, ybP = ns.P.xs[0], ns.P.xs[1]
        solver = {"U": ns.solver_U[rk], "V": ns.solver_V[rk], "T": ns.solver_T[rk]}
        xl, xr = self.xl, self.xr
        state = {k: xl("S" + k, M0, M1c) for k in names}
        pres = xl("pres", N0, W)
        mparts = [(o, max(0, min(o + n_, M0) - o)) for o, n_ in lay.rp]      # Galerkin rows per rank
        yl = lambda name, rows, cols: self.Yp[name][:rows, :cols]
        npass = [0]

        def add(L, label=None):
            L.finalize()
            calls.keep.append(L)
            fn, args = L.args()
            npass[0] += 1
            calls.add(fn, *args, label=label or "pass[P%s%d]" % ("Y" if L.layout else "X", npass[0]))

        def barrier():
            fn, args = self.mem.barrier_args()
            calls.add(fn, *args, label="peer_barrier")

        # ---- PX1 (local): F -> Sx F, dx Sx F / sx; pres -> dpdx; pushed to the row owners
        import os
        colpush = os.environ.get("PDE_SLAB_COLPUSH", "0") == "1"
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        if colpush:
            for k in names:
                L.job(M1c).load(state[k]).stencil(xb[k]).store(self.yseg("y_c" + k, lay.rp, N0)).diff(sx) \
                    .store(self.yseg("y_d" + k, lay.rp, N0))
            L.job(W).load(pres).store(self.yseg("y_pres", lay.rp, N0)).diff(sx).store(self.yseg("y_dpdx", lay.rp, N0))
            add(L)
        else:
            for k in names:
                L.job(M1c).load(state[k]).stencil(xb[k]).store(xl("c" + k, N0, M1c)).diff(sx).store(xl("d" + k, N0, M1c))
            L.job(W).load(pres).diff(sx).store(xl("dpdx", N0, W))
            add(L)
            L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
            for k in names:
                self.push_x2y(L, "c" + k, "y_c" + k, lay.rp, M1)
                self.push_x2y(L, "d" + k, "y_d" + k, lay.rp, M1)
            self.push_x2y(L, "pres", "y_pres", lay.rp, N1)
            self.push_x2y(L, "dpdx", "y_dpdx", lay.rp, N1)
            add(L, "exchange[push c, d, pres, dpdx]")
        barrier()
        # ---- PY2 (local rows): y stencils / derivatives, right-hand-side parts
        eU, eV, fU, fV, fT, gU, gV, gT = self.eY
        eF, fF, gF = {"U": eU, "V": eV}, {"U": fU, "V": fV, "T": fT}, {"U": gU, "V": gV, "T": gT}
        rest = dict(zip(names, self.restY))
        cY = {k: yl("y_c" + k, N0r, M1) for k in names}
        dY = {k: yl("y_d" + k, N0r, M1) for k in names}
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in names:
            p = L.job(N0r, r0).load(cY[k]).stencil(yb[k])
            if k != "T":
                p.store(eF[k])
            p.diff(sz).store(gF[k])
            L.job(N0r, r0).load(dY[k]).stencil(yb[k]).store(fF[k])
        st = {k: self.tables.stencil_elem(yb[k]) for k in names}
        L.job(N0r, r0).load(cY["U"]).stencil(yb["U"]).axpy(-dt * a, yl("y_dpdx", N0r, N1)).store(rest["U"])
        L.job(N0r, r0).load(yl("y_pres", N0r, N1)).diff(sz).scale(-dt * a).axpy(1.0, cY["V"], stencil=st["V"]) \
            .axpy(dt * a, cY["T"], stencil=st["T"]).axpy(dt * a, self.tbc_Y).store(rest["V"])
        L.job(N0r, r0).load(cY["T"]).stencil(yb["T"]).axpy(dt * a * ns.kappa, self.dTbcdz2_Y).store(rest["T"])
        add(L)
        # ---- backward y-DCT on my rows, pushed to the column owners
        self._dct(calls, self.plan1, ops.BWD, 1, self.eY, [zz[:, :D1] for zz in self.Z8])
        L = PS.PassLaunch(PS.ROW, min(D1, 2048), self.tables)
        for k in range(8):
            self.push_y2x_wide(L, self.Z8[k], "ZX8_%d" % k, N0r, r0)
        add(L, "exchange[push 8 (N x D)]")
        barrier()
        # ---- backward x-DCT, products, forward x-DCT on my columns
        new, old = self.uw[rk % 2], self.uw[(rk + 1) % 2]
        dxU, dxV, dxT, dzU, dzV, dzT = self.phys
        dst = [new[0], new[1], dxU, dxV, dxT, dzU, dzV, dzT]
        self._dct(calls, self.plan0, ops.BWD, 0, [xl("ZX8_%d" % k, N0, Wd) for k in range(8)], [t[:, :Wd] for t in dst])
        use_old = c != 0.0
        calls.add(Lb.pde_conv_products, D0 * self.Wdmax, b, c, _ptr(new[0]), _ptr(new[1]),
                  _ptr(old[0]) if use_old else None, _ptr(old[1]) if use_old else None,
                  _ptr(dxU), _ptr(dzU), _ptr(dxV), _ptr(dzV), _ptr(dxT), _ptr(dzT), _ptr(self.dTbcdz1_X))
        self._dct(calls, self.plan0, ops.FWD, 0, [t[:, :Wd] for t in (dxU, dxV, dxT)], [t[:, :Wd] for t in self.F3X])
        L = PS.PassLaunch(PS.ROW, self.Wdmax, self.tables)
        for k in range(3):
            self.push_x2y(L, None, "F3Y_%d" % k, lay.rp, D1, colpart=self.dcp, src=self.F3X[k])
        add(L, "exchange[push 3 (N x D)]")
        barrier()
        # ---- forward y-DCT; PYh: h = Ay^-1 By (rest - dt conv), pushed to the column owners
        self._dct(calls, self.plan1, ops.FWD, 1, [yl("F3Y_%d" % k, N0r, D1) for k in range(3)], self.convY)
        conv = dict(zip(names, self.convY))
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in names:
            L.job(N0r, r0).lincomb([(1.0, rest[k]), (-dt, conv[k])]).band(solver[k].plan_for_rhs[1].band) \
                .fdma(solver[k].plan_for_lhs[1]).store(xr("h" + k, M1, r0))
        add(L)
        barrier()
        # ---- PXh (local): F* = Ax^-1 Bx h -> state; x parts of the divergence
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for k in names:
            p = L.job(M1c).load(xl("h" + k, N0, M1c)).band(solver[k].plan_for_rhs[0].band) \
                .fdma(solver[k].plan_for_lhs[0]).store(state[k])
            if k == "U":
                p.stencil(xb["U"]).diff(sx).store(self.yseg("y_aUx", lay.rp, N0) if colpush else xl("aUx", N0, M1c))
            elif k == "V":
                p.stencil(xb["V"]).store(self.yseg("y_aVx", lay.rp, N0) if colpush else xl("aVx", N0, M1c))
        add(L)
        if not colpush:
            L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
            self.push_x2y(L, "aUx", "y_aUx", lay.rp, M1)
            self.push_x2y(L, "aVx", "y_aVx", lay.rp, M1)
            add(L, "exchange[push div parts]")
        barrier()
        # ---- PYd: div = Sy (dx Sx U / sx) + dz Sy (Sx V) / sz; R' = div Hy^T; both pushed
        sp = ns.solver_P
        Hy, Qy = sp.plan_for_rhs[1].dense, sp.plan_for_lhs[1].dense
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        L.job(N0r, r0).load(yl("y_aVx", N0r, M1)).stencil(yb["V"]).diff(sz) \
            .axpy(1.0, yl("y_aUx", N0r, M1), stencil=st["U"]).store(self.divY)
        add(L)
        calls.add(Lb.pde_gemm_f64, 1, _ptr(self.divY), _ld(self.divY), _ptr(Hy), _ld(Hy), _ptr(self.RY), _ld(self.RY),
                  N0r, M1, N1)
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        L.job(N0r, r0).load(self.RY).store(xr("R", M1, r0))
        L.job(N0r, r0).load(self.divY).store(xr("div", N1, r0))
        add(L, "exchange[push R, div]")
        barrier()
        # ---- PX6 (local): q = Bx R'; per-column Poisson solves; pushed to the row owners
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        p = L.job(M1c).load(xl("R", N0, M1c)).band(sp.plan_for_rhs[0].band).poisson(self.ptab)
        p.store(self.yseg("WY", mparts, M0) if colpush else xl("R", M0, M1c))
        add(L)
        if not colpush:
            L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
            self.push_x2y(L, "R", "WY", mparts, M1)
            add(L, "exchange[push W]")
        barrier()
        # ---- P = W Qy^T on my rows; PY7: P[0,0] = 0, e1 = Sy P, bU = Gy e1, bV = Gy dz e1 / sz (pushed)
        WY = yl("WY", M0r, M1)
        calls.add(Lb.pde_gemm_f64, 1, _ptr(WY), _ld(WY), _ptr(Qy), _ld(Qy), _ptr(self.PY), _ld(self.PY), M0r, M1, M1)
        Py = self.PY[:M0r]
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        L.job(M0r, r0).load(Py).setz0(0).store(Py, only_seq=0).stencil(ybP).store(xr("e1", N1, r0)).diff(sz) \
            .from_cheb(yb["V"]).store(xr("bV", M1, r0))
        L.job(M0r, r0).load(Py).setz0(0).stencil(ybP).from_cheb(yb["U"]).store(xr("bU", M1, r0))
        add(L)
        barrier()
        # ---- PX8 (local): velocity projection and pressure update
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        L.job(M1c).load(xl("bU", M0, M1c)).stencil(xbP).diff(sx).from_cheb(xb["U"]) \
            .axpy(1.0, state["U"], scale_buf=-1.0).store(state["U"])
        L.job(M1c).load(xl("bV", M0, M1c)).stencil(xbP).from_cheb(xb["V"]).axpy(1.0, state["V"], scale_buf=-1.0) \
            .store(state["V"])
        L.job(W).load(xl("e1", M0, W)).stencil(xbP).scale(1.0 / (dt * a)) \
            .lincomb([(1.0, pres), (-(1.0 * ns.nu), xl("div", N0, W))], accumulate=True).store(pres)
        add(L)
        return calls
