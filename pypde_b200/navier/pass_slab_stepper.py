"""
Slab-decomposed IMEX stage on the fused axis passes (one process per GPU, one node).

Same pass structure as pass_stepper.PassStepper.  Axis-0 passes (X) run on the local column slab
(all rows, my columns), axis-1 passes (Y) on the local row slab.  There is no all-to-all and no
pack / unpack kernel: every X-layout array lives in peer-mapped memory (navier/peer.py), and the
distributed transposes ARE the loads and stores of the row passes --

    pull:  LOAD of a row pass reads the row's pieces straight out of the peers' column slabs,
    push:  STORE of a row pass writes the pieces of its result row into the peers' column slabs

(2 KB contiguous per piece at 2048^2 on 8 GPUs) -- over NVLink, overlapped with the recurrences of the
other rows in flight.  The transforms and the dense projections keep their own kernels; their inputs /
outputs cross with a two-instruction row pass (load -> store).  A device-side barrier (pde_peer_barrier,
one tiny launch) separates producers and consumers on different GPUs; the whole step is one CUDA graph.

Reference: navier/rbc2d.py:396-434 (the stage), with the regrouping documented in pass_stepper.py.
"""
import torch
import torch.distributed as dist

from .. import _cabi as C
from .. import ops
from .. import passes as PS
from .fast_stepper import FastStepper, _Calls, _ptr, _ld
from .peer import PeerMem, SlabLayout, row_segments


class PassSlabStepper(FastStepper):
    @staticmethod
    def supported(ns, world):
        N0, N1 = ns.shape
        return ns.beta == 1.0 and N0 % 2 == 0 and N1 % 4 == 0 and N0 <= 4096 and N1 <= 4096 and world <= PS.MAX_SEG

    def __init__(self, ns, group=None):
        self.group = group
        self.P, self.r = dist.get_world_size(group), dist.get_rank(group)
        self.tables = PS.TableCache()
        self.bound = False
        FastStepper.__init__(self, ns)

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        lay = self.lay = SlabLayout(N0, N1, D0, D1, self.P, self.r)
        W = lay.Wmax
        # X-layout arrays in peer memory, identical offsets on every rank: name -> (offset, rows)
        self.xoff, off = {}, PeerMem.HEADER

        def xarr(name, rows):
            nonlocal off
            self.xoff[name] = (off, rows)
            off += ((rows * W * 8 + 255) // 256) * 256
        for k in "TUV":
            for pre in ("S", "c", "d", "e", "f", "g", "z", "conv"):
                xarr(pre + k, N0)
        for name in ("pres", "dpdx", "dpdz", "aU", "aV", "div", "q", "R", "e1", "bU", "bV"):
            xarr(name, N0)
        for k in range(8):
            xarr("X8_%d" % k, D0)
        for k in range(3):
            xarr("F3_%d" % k, D0)
        # Y-layout arrays that are filled by the OTHER ranks (pushed row pieces): name -> (offset, rows)
        self.yoff = {}
        D0rmax, N0rmax = max(n_ for _, n_ in lay.dp), max(n_ for _, n_ in lay.rp)

        def yarr(name, rows):
            nonlocal off
            self.yoff[name] = (off, rows)
            off += ((rows * N1 * 8 + 255) // 256) * 256
        for k in range(8):
            yarr("Y8_%d" % k, D0rmax)
        yarr("qY", N0rmax)
        yarr("WY", N0rmax)
        for name in ("cU", "cV", "cT", "dU", "dV", "dT", "pres", "zU", "zV", "zT"):
            yarr("y_" + name, N0rmax)
        self.mem = PeerMem(off, self.group)
        self.X = {name: self.mem.view(o, (rows, W)) for name, (o, rows) in self.xoff.items()}
        self.Yp = {name: self.mem.view(o, (rows, N1)) for name, (o, rows) in self.yoff.items()}
        # Y-layout arrays (local)
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=self.dev)
        self.Y8 = [self.Yp["Y8_%d" % k][:lay.D0r] for k in range(8)]
        self.phys = [z(lay.D0r, D1) for _ in range(6)]
        self.uw = [[z(lay.D0r, D1), z(lay.D0r, D1)] for _ in range(2)]
        self.F3y = [z(lay.D0r, N1) for _ in range(3)]
        self.qY, self.WY = self.Yp["qY"][:lay.N0r], self.Yp["WY"][:lay.N0r, :M1]
        self.RY, self.PY = z(lay.N0r, M1), z(lay.N0r, M1)

    def _tables(self):
        FastStepper._tables(self)
        ns, lay = self.ns, self.lay
        c0, W = lay.c0, lay.W
        self.tbc_X = self.tbc_cheby[:, c0:c0 + W].contiguous()
        self.dTbcdz2_X = self.dTbcdz2[:, c0:c0 + W].contiguous()
        self.dTbcdz1_Y = self.dTbcdz1[lay.d0:lay.d0 + lay.D0r].contiguous()
        pp = ns.solver_P.plan_for_lhs[0]
        self.poisson_local = ops.PoissonPlan(pp._Ad, pp._Cd, pp.alpha[c0:c0 + lay.M1c], pp.singular)
        self.ptab = PS.PoissonTables(self.poisson_local, PS.lg_for(self.M0))

    # ------------------------------------------------------------------ operands
    def xl(self, name, rows, cols):
        """local view of an X-layout array"""
        return self.X[name][:rows, :cols]

    def xr(self, name, ncols, row0):
        """the rows of an X-layout array seen from the Y layout: one segment per rank (peer mappings)"""
        off, _ = self.xoff[name]
        ptrs, lds, starts = row_segments(self.mem.base, off, self.lay.Wmax, row0, self.lay.col_starts(ncols), ncols)
        return PS.Operand(ptrs, lds, starts, ncols, keep=(self.mem,))

    def push_rows(self, L, xname, yname, parts, ncols):
        """Jobs of a row pass over the LOCAL column slab of X-layout array `xname` that write every row piece
        (my columns) into the Y-layout array `yname` of the rank that owns the row: one job per destination,
        2 KB contiguous remote stores instead of remote loads (posted writes: ~1.7x the pull rate measured)."""
        lay = self.lay
        w = lay.ncols_local(ncols)
        yo, _ = self.yoff[yname]
        for s, (row0, nrows) in enumerate(parts):
            if nrows <= 0 or w <= 0:
                continue
            src = self.X[xname][row0:row0 + nrows, :w]
            dst = PS.Operand([self.mem.base[s] + yo + 8 * lay.c0], [self.N1], [0, w], w, keep=(self.mem,))
            L.job(nrows).load(src).store(dst)

    # ------------------------------------------------------------------ state movement
    def scatter(self):
        """Global fields (replicated on every rank) -> local X-layout slabs."""
        ns, lay = self.ns, self.lay
        for k, f in (("T", ns.T), ("U", ns.U), ("V", ns.V)):
            self.xl("S" + k, self.M0, lay.M1c).copy_(f.vhat[:, lay.c0:lay.c0 + lay.M1c])
        self.xl("pres", self.N0, lay.W).copy_(ns.pres.vhat[:, lay.c0:lay.c0 + lay.W])
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def gather(self):
        """Local slabs -> global fields on every rank (diagnostics, parity tests, I/O)."""
        ns, lay = self.ns, self.lay
        for name, f, rows, total in (("ST", ns.T, self.M0, self.M1), ("SU", ns.U, self.M0, self.M1),
                                     ("SV", ns.V, self.M0, self.M1), ("pres", ns.pres, self.N0, self.N1)):
            loc = self.X[name][:rows].contiguous()
            parts = [torch.empty_like(loc) for _ in range(self.P)]
            dist.all_gather(parts, loc, group=self.group)
            for s, (o, w) in enumerate(lay.cp):
                w = max(0, min(o + w, total) - o)
                if w:
                    f.vhat[:, o:o + w].copy_(parts[s][:, :w])

    def local_state(self):
        lay = self.lay
        return [self.xl("S" + k, self.M0, lay.M1c) for k in "TUV"] + [self.xl("pres", self.N0, lay.W)]

    # ------------------------------------------------------------------ the stage
    def bind(self):
        self.scatter()
        self.bound = True
        self.stage_calls = [self._build_stage(rk) for rk in range(self.ns.nstage)]

    def stage(self, rk):
        if not self.bound:
            self.bind()
        self.stage_calls[rk].run()

    def _build_stage(self, rk):
        ns, Lb, lay = self.ns, C.lib(), self.lay
        calls = _Calls()
        N0, N1, M0, M1, D0, D1 = self.N0, self.N1, self.M0, self.M1, self.D0, self.D1
        W, M1c, N0r, M0r, D0r, r0, d0 = lay.W, lay.M1c, lay.N0r, lay.M0r, lay.D0r, lay.r0, lay.d0
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

        # ---- PX1 (local): F -> Sx F, dx Sx F / sx; pres -> dpdx
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        for k in names:
            L.job(M1c).load(state[k]).stencil(xb[k]).store(xl("c" + k, N0, M1c)).diff(sx).store(xl("d" + k, N0, M1c))
        L.job(W).load(pres).diff(sx).store(xl("dpdx", N0, W))
        add(L)
        # ---- rows of c, d, pres -> their owners (pushed), then PY2 on local rows; its results are pushed back
        L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
        for k in names:
            self.push_rows(L, "c" + k, "y_c" + k, lay.rp, M1)
            self.push_rows(L, "d" + k, "y_d" + k, lay.rp, M1)
        self.push_rows(L, "pres", "y_pres", lay.rp, N1)
        add(L, "exchange[push c, d, pres]")
        barrier()
        yl = lambda name, rows, cols: self.Yp["y_" + name][:rows, :cols]
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in names:
            L.job(N0r, r0).load(yl("c" + k, N0r, M1)).stencil(yb[k]).store(xr("e" + k, N1, r0)).diff(sz) \
                .store(xr("g" + k, N1, r0))
            L.job(N0r, r0).load(yl("d" + k, N0r, M1)).stencil(yb[k]).store(xr("f" + k, N1, r0))
        L.job(N0r, r0).load(yl("pres", N0r, N1)).diff(sz).store(xr("dpdz", N1, r0))
        add(L)
        barrier()
        # ---- backward x-DCT of the 8 coefficient arrays (local), rows of the result pulled into the Y layout
        src = ["eU", "eV", "fU", "fV", "fT", "gU", "gV", "gT"]
        self._dct(calls, self.plan0, ops.BWD, 0, [xl(s, N0, W) for s in src], [xl("X8_%d" % k, D0, W) for k in range(8)])
        L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
        for k in range(8):
            self.push_rows(L, "X8_%d" % k, "Y8_%d" % k, lay.dp, N1)
        add(L, "exchange[push 8 (D x N)]")
        barrier()
        # ---- backward y-DCT, products, forward y-DCT (local rows)
        new, old = self.uw[rk % 2], self.uw[(rk + 1) % 2]
        dxU, dxV, dxT, dzU, dzV, dzT = self.phys
        self._dct(calls, self.plan1, ops.BWD, 1, self.Y8, [new[0], new[1], dxU, dxV, dxT, dzU, dzV, dzT])
        use_old = c != 0.0
        calls.add(Lb.pde_conv_products, D0r * D1, b, c, _ptr(new[0]), _ptr(new[1]),
                  _ptr(old[0]) if use_old else None, _ptr(old[1]) if use_old else None,
                  _ptr(dxU), _ptr(dzU), _ptr(dxV), _ptr(dzV), _ptr(dxT), _ptr(dzT), _ptr(self.dTbcdz1_Y))
        self._dct(calls, self.plan1, ops.FWD, 1, [dxU, dxV, dxT], self.F3y)
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in range(3):
            L.job(D0r, d0).load(self.F3y[k]).store(xr("F3_%d" % k, N1, d0))
        add(L, "exchange[push 3 (D x N)]")
        barrier()
        # ---- forward x-DCT (local), PX3: z = Ax^-1 Bx (Sy Sx F + rhs)
        conv = {k: xl("conv" + k, N0, W) for k in names}
        self._dct(calls, self.plan0, ops.FWD, 0, [xl("F3_%d" % k, D0, W) for k in range(3)],
                  [conv["U"], conv["V"], conv["T"]])
        e = {k: xl("e" + k, N0, W) for k in names}
        z = {k: xl("z" + k, M0, W) for k in names}
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        p = L.job(W).lincomb([(1.0, e["U"]), (-dt * a, xl("dpdx", N0, W)), (-dt, conv["U"])])
        p.band(solver["U"].plan_for_rhs[0].band).fdma(solver["U"].plan_for_lhs[0]).store(z["U"])
        p = L.job(W).lincomb([(1.0, e["V"]), (-dt * a, xl("dpdz", N0, W)), (-dt, conv["V"]), (dt * a, e["T"]),
                              (dt * a, self.tbc_X)])
        p.band(solver["V"].plan_for_rhs[0].band).fdma(solver["V"].plan_for_lhs[0]).store(z["V"])
        p = L.job(W).lincomb([(1.0, e["T"]), (-dt, conv["T"]), (dt * a * ns.kappa, self.dTbcdz2_X)])
        p.band(solver["T"].plan_for_rhs[0].band).fdma(solver["T"].plan_for_lhs[0]).store(z["T"])
        add(L)
        L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
        for k in names:
            self.push_rows(L, "z" + k, "y_z" + k, mparts, N1)
        add(L, "exchange[push z]")
        barrier()
        # ---- PY4: F* = Ay^-1 By z -> pushed into the state slabs; y parts of the divergence
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        for k in names:
            p = L.job(M0r, r0).load(yl("z" + k, M0r, N1)).band(solver[k].plan_for_rhs[1].band) \
                .fdma(solver[k].plan_for_lhs[1]).store(xr("S" + k, M1, r0))
            if k == "U":
                p.stencil(yb["U"]).store(xr("aU", N1, r0))
            elif k == "V":
                p.stencil(yb["V"]).diff(sz).store(xr("aV", N1, r0))
        add(L)
        barrier()
        # ---- PX5 (local): div, q = Bx div
        sp = ns.solver_P
        L = PS.PassLaunch(PS.COL, N0, self.tables)
        p = L.job(W).load(xl("aU", M0, W)).stencil(xb["U"]).diff(sx)
        p.axpy(1.0, xl("aV", N0, W), stencil=self.tables.stencil_elem(xb["V"])).store(xl("div", N0, W))
        p.band(sp.plan_for_rhs[0].band).store(xl("q", M0, W))
        add(L)
        # ---- R = q Hy^T on my rows
        Hy, Qy = sp.plan_for_rhs[1].dense, sp.plan_for_lhs[1].dense
        L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
        self.push_rows(L, "q", "qY", mparts, N1)
        add(L, "exchange[push q]")
        barrier()
        calls.add(Lb.pde_gemm_f64, 1, _ptr(self.qY), _ld(self.qY), _ptr(Hy), _ld(Hy), _ptr(self.RY), _ld(self.RY),
                  M0r, M1, N1)
        L = PS.PassLaunch(PS.ROW, N1, self.tables)
        L.job(M0r, r0).load(self.RY).store(xr("R", M1, r0))
        add(L, "exchange[push R]")
        barrier()
        # ---- PX6 (local): per-column Poisson solves
        Rx = xl("R", M0, M1c)
        L = PS.PassLaunch(PS.COL, M0, self.tables)
        L.job(M1c).load(Rx).poisson(self.ptab).store(Rx)
        add(L)
        # ---- P = W Qy^T on my rows; PY7: P[0,0] = 0, e1 = Sy P, bU = Gy e1, bV = Gy dz e1 / sz (pushed)
        L = PS.PassLaunch(PS.ROW, lay.Wmax, self.tables)
        self.push_rows(L, "R", "WY", mparts, M1)
        add(L, "exchange[push W]")
        barrier()
        calls.add(Lb.pde_gemm_f64, 1, _ptr(self.WY), _ld(self.WY), _ptr(Qy), _ld(Qy), _ptr(self.PY), _ld(self.PY),
                  M0r, M1, M1)
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

    def check(self):
        self.mem.check()

    def close(self):
        self.mem.close()
