"""
Host side of the fused axis passes (pde_pass_run, include/pypde_b200.h): builds the per-job programs
(device arrays of pde_pass_ins) and the coefficient tables in the layouts the kernel expects.

One `PassLaunch` = one kernel launch = one axis pass of SURVEY.md §8(d): every job is a 2-D array whose
sequences (rows for axis 1, columns for axis 0) run the same small program of 1-D operators

    load -> stencil / derivative / banded product / Thomas solve / axpy ... -> store

with the operators of the reference they replace:
    stencil        GalerkinChebyshev.to_chebyshev           bases/chebyshev.py:287-293
    stencil_t+tdma GalerkinChebyshev.from_chebyshev         bases/chebyshev.py:295-337 -> tdma.f90:55-106
    diff           differentiate_cheby.diff_*               bases/fortran/differentiate_cheby.f90:28-53
    band           PlanRHS.solve (banded B)                 solver/plans.py:54-74
    fdma           Plan_fdma.solve                          solver/linalg/fortran/fdma.f90:1-98
    poisson        Plan_Poisson.solve (per-column tables)   fdma.f90:146-195
"""
import ctypes

import numpy as np
import torch

from . import _cabi as C

MAX_SEG = 8
MAX_TERMS = 6
MAX_INS = 16
SLOTS = 4
COL, ROW = 0, 1
LOAD, STORE, AXPY, SCALE, SETZ0, POINT, DIFF, REC1, REC2, TABLES, LINCOMB = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11
F_DESC, F_PERSEQ, F_SCALED, F_STENCIL, F_ONLY_SEQ, F_ACCUM, F_BULK = 1, 2, 4, 8, 16, 32, 64
import os as _os
BULK_STORES = _os.environ.get("PDE_PASS_BULK", "1") == "1"      # ROW stores through the TMA bulk-copy engine


class PassIns(ctypes.Structure):
    """pde_pass_ins (include/pypde_b200.h)"""
    _fields_ = [("op", ctypes.c_int), ("n", ctypes.c_int), ("flags", ctypes.c_int), ("nseg", ctypes.c_int),
                ("f0", ctypes.c_double), ("f1", ctypes.c_double), ("coef", ctypes.c_double * MAX_TERMS),
                ("p", ctypes.c_void_p * MAX_SEG), ("ld", ctypes.c_long * MAX_SEG),
                ("start", ctypes.c_int * (MAX_SEG + 1)), ("off", ctypes.c_int * 4), ("slot", ctypes.c_int * 3)]


class PassJob(ctypes.Structure):
    """pde_pass_job"""
    _fields_ = [("prog", ctypes.c_void_p), ("nins", ctypes.c_int), ("nseq", ctypes.c_int), ("seq0", ctypes.c_int),
                ("pad_", ctypes.c_int)]


def lg_for(n):
    """log2 of the units per lane for sequences of up to n elements (NUP = 32 << lg units of 2 elements)."""
    nu = (int(n) + 1) // 2
    segu = max(1, -(-nu // 32))
    lg = int(np.ceil(np.log2(segu))) if segu > 1 else 0
    if lg > 6:
        raise ValueError("fused axis passes support sequences up to 4096 elements (got %d)" % n)
    return lg


# ------------------------------------------------------------------ operands
class Operand:
    """Elements [0, n) of every sequence of a 2-D array, possibly split along the sequence into segments with
    their own base pointer / leading dimension (peer slabs)."""

    def __init__(self, ptrs, lds, starts, n, keep=()):
        assert 1 <= len(ptrs) <= MAX_SEG and len(ptrs) == len(lds) and len(starts) == len(ptrs) + 1
        self.ptrs, self.lds, self.starts, self.n, self.keep = list(ptrs), list(lds), list(starts), int(n), keep

    @staticmethod
    def of(t, layout):
        """A local tensor: ROW: sequences = rows, COL: sequences = columns."""
        assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1)
        n = t.shape[1] if layout == ROW else t.shape[0]
        return Operand([t.data_ptr()], [t.stride(0)], [0, n], n, keep=(t,))

    def check(self, layout):
        if layout == ROW:
            for p, ld, s in zip(self.ptrs, self.lds, self.starts):
                assert p % 16 == 0 and ld % 2 == 0 and s % 2 == 0, \
                    "row passes need 16-byte aligned rows (even leading dimensions and segment starts)"

    def fill(self, ins):
        ins.nseg, ins.n = len(self.ptrs), self.n
        for s, (p, ld) in enumerate(zip(self.ptrs, self.lds)):
            ins.p[s], ins.ld[s] = p, ld
        for s, st in enumerate(self.starts):
            ins.start[s] = st


# ------------------------------------------------------------------ tables
def _pad(t, n):
    out = np.zeros(n)
    t = np.asarray(t, dtype=float).ravel()
    out[: min(n, t.size)] = t[:n]
    return out


def unit_table(elem, lg):
    """Element table -> plain unit table (2 NUP doubles, zero padded) on the device."""
    return C.upload(_pad(elem, 2 * (32 << lg)))


def seg_order(t, lg):
    """(..., 2 NUP) element tables (torch or numpy) -> segment order (..., SEGU, 32, 2): the unit of (lane, j)
    at [j, lane]."""
    segu = 1 << lg
    shp = tuple(t.shape[:-1])
    if isinstance(t, torch.Tensor):
        return t.reshape(shp + (32, segu, 2)).transpose(-3, -2).contiguous()
    return np.ascontiguousarray(np.swapaxes(t.reshape(shp + (32, segu, 2)), -3, -2))


def seg_table(elem, lg):
    return C.upload(seg_order(_pad(elem, 2 * (32 << lg)), lg))


def shift2(t, n=None):
    """out[i] = t[i-2] (out[0] = out[1] = 0)."""
    t = np.asarray(t, dtype=float).ravel()
    out = np.zeros((t.size + 2) if n is None else n)
    m = min(t.size, out.size - 2)
    out[2:2 + m] = t[:m]
    return out


class TableCache:
    """Device tables of the operators, keyed by (object identity, kind, lg)."""

    def __init__(self):
        self._c = {}
        self._keep = []

    def get(self, key, make):
        if key not in self._c:
            self._c[key] = make()
        return self._c[key]

    def stencil(self, base, lg):
        """taps of u = S v: u_i = v_i + s_{i-2} v_{i-2}"""
        def make():
            s = base._tables_host()[0]
            return unit_table(shift2(s), lg)
        return self.get((base.N, base.id, "S", lg), make)

    def stencil_elem(self, base):
        """element table st_i = s_{i-2} for AXPY with the STENCIL flag (any length >= N)"""
        def make():
            s = base._tables_host()[0]
            return C.upload(_pad(shift2(s), base.N + 8))
        return self.get((base.N, base.id, "Se"), make)

    def stencil_t(self, base, lg):
        """taps of S^T u: y_i = u_i + s_i u_{i+2}"""
        return self.get((base.N, base.id, "ST", lg), lambda: unit_table(base._tables_host()[0], lg))

    def tdma(self, base, lg):
        """(T0, T1) of the forward sweep g_i = (rhs_i - a_{i-2} g_{i-2}) / den_i and T1 of the back substitution
        x_i = g_i - w_i x_{i+2} (tdma.f90:82-104), reciprocal-scaled."""
        def make():
            s, a, den, w = base._tables_host()
            rden = 1.0 / den
            t1 = shift2(a, den.size) * rden
            return seg_table(rden, lg), seg_table(t1, lg), seg_table(w[: max(den.size - 2, 0)], lg)
        return self.get((base.N, base.id, "tdma", lg), make)

    def fdma(self, plan, lg):
        """Plan_fdma (after FDMA_LU): forward x_i -= l_{i-2} x_{i-2}; backward x_i = (x_i - u1_i x_{i+2} -
        u2_i x_{i+4}) / d_i with reciprocal-scaled tables (fdma.f90:26-36)."""
        def make():
            n = plan.d.size
            rd = 1.0 / plan.d
            return (seg_table(shift2(plan.l, n), lg), seg_table(rd, lg), seg_table(_pad(plan.u1, n) * rd, lg),
                    seg_table(_pad(plan.u2, n) * rd, lg))
        self._keep.append(plan)             # id() keys must not be recycled
        return self.get((id(plan), "fdma", lg), make)

    def band(self, band, lg):
        """unit tables of the diagonals of a Band (even offsets)"""
        def make():
            return [unit_table(band.host[d], lg) for d in range(band.ndiag)]
        # keyed by content: the B matrices depend on the grid only (ensemble members share them)
        return self.get((band.n_out, band.n_in, tuple(band.offsets), hash(band.host.tobytes()), "band", lg), make)


# ------------------------------------------------------------------ programs
class Program:
    """The instruction list of one job (one 2-D array); all operators act on the job's sequences."""

    def __init__(self, launch, nseq, seq0=0):
        self.L, self.nseq, self.seq0, self.ins = launch, int(nseq), int(seq0), []
        self.slots = []             # shared-memory table slots of this job: list of device tables

    def _new(self, op):
        i = PassIns()
        i.op, i.nseg = op, 1
        for k in range(3):
            i.slot[k] = -1
        self.ins.append(i)
        return i

    def _slot(self, tab):
        """shared-memory slot of a (job-wide) recurrence table, or -1 when the four slots are taken"""
        for k, t in enumerate(self.slots):
            if t is tab:
                return k
        if len(self.slots) < SLOTS:
            self.slots.append(tab)
            return len(self.slots) - 1
        return -1

    def finish(self):
        """instruction list with the TABLES prologue"""
        out = []
        if self.slots:
            i = PassIns()
            i.op, i.n, i.nseg = TABLES, len(self.slots), 1
            for k, t in enumerate(self.slots):      # positional: table k -> slot k
                i.p[k] = t.data_ptr()
            for k in range(3):
                i.slot[k] = -1
            out.append(i)
        out += self.ins
        if len(out) > MAX_INS:
            raise ValueError("axis-pass program too long (%d > %d instructions)" % (len(out), MAX_INS))
        return out

    def _opnd(self, x):
        if isinstance(x, Operand):
            o = x
        else:
            o = Operand.of(x, self.L.layout)
            nseq = x.shape[0] if self.L.layout == ROW else x.shape[1]
            assert nseq >= self.nseq, "operand has fewer sequences than the job"
        o.check(self.L.layout)
        assert o.n <= 2 * self.L.NUP
        self.L.keep.append(o.keep)
        return o

    # -- data movement
    def load(self, x):
        self._opnd(x).fill(self._new(LOAD))
        return self

    def store(self, x, only_seq=None):
        i = self._new(STORE)
        self._opnd(x).fill(i)
        if BULK_STORES and self.L.layout == ROW and only_seq is None:
            i.flags |= F_BULK
        if only_seq is not None:
            i.flags |= F_ONLY_SEQ
            i.off[0] = int(only_seq)
        return self

    def axpy(self, coef, x, scale_buf=None, stencil=None):
        """buffer (times scale_buf) += coef * x; stencil: element table st (st_i = s_{i-2}): x_i + st_i x_{i-2}"""
        i = self._new(AXPY)
        self._opnd(x).fill(i)
        i.f0 = float(coef)
        if scale_buf is not None:
            i.flags |= F_SCALED
            i.f1 = float(scale_buf)
        if stencil is not None:
            i.flags |= F_STENCIL
            i.p[MAX_SEG - 1] = stencil.data_ptr()
            assert i.nseg < MAX_SEG
            self.L.keep.append(stencil)
        return self

    def scale(self, f):
        self._new(SCALE).f0 = float(f)
        return self

    def setz0(self, seq):
        self._new(SETZ0).off[0] = int(seq)
        return self

    # -- operators
    def point(self, taps):
        """taps: list of (unit offset, unit table or None)"""
        offs = [o for o, _ in taps]
        assert 1 <= len(taps) <= 4 and all(-1 <= o <= 2 for o in offs)
        assert all(o >= 0 for o in offs) or all(o <= 0 for o in offs)
        i = self._new(POINT)
        i.n = len(taps)
        for t, (o, tab) in enumerate(taps):
            i.off[t] = o
            i.p[t] = tab.data_ptr() if tab is not None else None
            if tab is not None:
                assert tab.numel() >= 2 * self.L.NUP
                self.L.keep.append(tab)
        return self

    def diff(self, div=1.0):
        self._new(DIFF).f0 = 1.0 / float(div)
        return self

    def lincomb(self, terms, accumulate=False):
        """buffer = [buffer +] sum coef * x over (coef, tensor) terms (local, unsegmented operands)"""
        assert 1 <= len(terms) <= MAX_TERMS
        i = self._new(LINCOMB)
        i.nseg = len(terms)
        i.flags = F_ACCUM if accumulate else 0
        for k, (coef, x) in enumerate(terms):
            o = self._opnd(x)
            assert len(o.ptrs) == 1
            i.p[k], i.ld[k], i.start[k], i.coef[k] = o.ptrs[0], o.lds[0], o.n, float(coef)
        return self

    def _rec(self, op, tabs, desc, perseq):
        i = self._new(op)
        i.flags = (F_DESC if desc else 0) | (F_PERSEQ if perseq else 0)
        # a recurrence runs either entirely from shared slots or entirely from global tables
        slots = [-1] * len(tabs)
        if not perseq:
            before = list(self.slots)
            slots = [self._slot(t) if t is not None else -1 for t in tabs]
            if any(s < 0 and t is not None for s, t in zip(slots, tabs)):
                self.slots = before
                slots = [-1] * len(tabs)
        for k, t in enumerate(tabs):
            i.slot[k] = slots[k]
            i.p[k] = t.data_ptr() if t is not None else None
            if perseq:
                i.ld[k] = int(perseq)
        self.L.keep += list(tabs)
        return self

    def rec1(self, t0, t1, desc=False, perseq=None):
        return self._rec(REC1, (t0, t1), desc, perseq)

    def rec2(self, t0, t1, t2, perseq=None):
        return self._rec(REC2, (t0, t1, t2), True, perseq)

    # -- composites on the framework's objects
    def stencil(self, base):
        """Galerkin -> Chebyshev coefficients (u = S v)"""
        return self.point([(0, None), (-1, self.L.tables.stencil(base, self.L.lg))])

    def from_cheb(self, base):
        """Chebyshev -> Galerkin coefficients: S^T product, then the offset-2 tridiagonal solve"""
        t0, t1, w = self.L.tables.tdma(base, self.L.lg)
        self.point([(0, None), (1, self.L.tables.stencil_t(base, self.L.lg))])
        self.rec1(t0, t1)
        return self.rec1(None, w, desc=True)

    def band(self, band):
        assert all(o % 2 == 0 for o in band.offsets), "banded operators of the Chebyshev bases couple equal parities"
        tabs = self.L.tables.band(band, self.L.lg)
        return self.point([(o // 2, t) for o, t in zip(band.offsets, tabs)])

    def fdma(self, plan):
        lt, rd, u1, u2 = self.L.tables.fdma(plan, self.L.lg)
        self.rec1(None, lt)
        return self.rec2(rd, u1, u2)

    def poisson(self, ptab):
        """ptab: PoissonTables (per-column tables in segment order)"""
        assert ptab.lg == self.L.lg
        self.rec1(None, ptab.lt, perseq=ptab.stride)
        return self.rec2(ptab.rd, ptab.u1, ptab.u2, perseq=ptab.stride)


class PoissonTables:
    """Per-column tables of Plan_Poisson ((A + lam_i C) x_i = b_i, fdma.f90:146-195) in the layout of the fused
    passes: built once from the plan's device factorisation (same operation order as init_fdma)."""

    def __init__(self, plan, lg):
        n, m = plan.n, plan.m
        self.lg = lg
        nup2 = 2 * (32 << lg)
        raw = []
        for which in range(5):
            t = torch.empty((n, m), dtype=torch.float64, device=C.device())
            C.check(C.lib().pde_poisson_plan_export(plan.handle, which, C.p(t)))
            raw.append(t)
        l, d, u1, u2, rd = raw
        offc = torch.empty((m,), dtype=torch.int32, device=C.device())
        C.check(C.lib().pde_poisson_plan_export(plan.handle, 5, ctypes.c_void_p(offc.data_ptr())))
        rd = rd.clone()
        rd[0, offc != 0] = 0.0                       # singular column: x_0 = 0 (fdma.f90:173-185)

        def lay(t, shift=0):
            full = torch.zeros((m, nup2), dtype=torch.float64, device=t.device)
            full[:, shift:shift + n - shift] = t[: n - shift].transpose(0, 1)
            return seg_order(full, lg)
        self.lt = lay(l, 2)                          # x_i -= l_{i-2} x_{i-2}
        self.rd = lay(rd)
        self.u1 = lay(u1 * rd)
        self.u2 = lay(u2 * rd)
        self.stride = nup2


class PassLaunch:
    """One launch of pde_pass_run: jobs with their programs, finalised into device arrays."""

    def __init__(self, layout, n, tables):
        self.layout, self.lg, self.tables = layout, lg_for(n), tables
        self.NUP = 32 << self.lg
        self.jobs, self.keep = [], []
        self._dev = None

    def job(self, nseq, seq0=0):
        p = Program(self, nseq, seq0)
        self.jobs.append(p)
        return p

    def finalize(self):
        nj = len(self.jobs)
        progs = [j.finish() for j in self.jobs]
        nins = sum(len(pr) for pr in progs)
        isz, jsz = ctypes.sizeof(PassIns), ctypes.sizeof(PassJob)
        buf = torch.empty(((jsz * nj + 15) // 16 * 16 + isz * nins,), dtype=torch.uint8, device=C.device())
        base = buf.data_ptr()
        ioff = (jsz * nj + 15) // 16 * 16
        jobs = (PassJob * nj)()
        ins = (PassIns * max(nins, 1))()
        k = 0
        for j, (pr, ins_list) in enumerate(zip(self.jobs, progs)):
            jobs[j].prog, jobs[j].nins, jobs[j].nseq, jobs[j].seq0 = base + ioff + isz * k, len(ins_list), pr.nseq, pr.seq0
            for i in ins_list:
                ctypes.memmove(ctypes.byref(ins[k]), ctypes.byref(i), isz)
                k += 1
        host = bytes(jobs) + b"\0" * (ioff - jsz * nj) + bytes(ins)[: isz * nins]
        buf.copy_(torch.frombuffer(bytearray(host), dtype=torch.uint8))
        self._dev = buf
        self.max_nseq = max(j.nseq for j in self.jobs)
        return self

    def args(self):
        """(C function, argument tuple) for fast_stepper._Calls"""
        if self._dev is None:
            self.finalize()
        return (C.lib().pde_pass_run, (self.layout, self.lg, len(self.jobs), self.max_nseq,
                                       ctypes.c_void_p(self._dev.data_ptr())))

    def run(self):
        fn, a = self.args()
        C.check(fn(*a, C.stream()))
