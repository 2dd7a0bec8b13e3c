"""GPU: the fused axis passes (pde_pass_run, pypde_b200/csrc/pass.cu) operator by operator against the CPU
oracle (oracle/kernels.py = C restatement of the reference's Fortran, oracle/pypde_port.Basis).

The chain-split recurrences re-associate across lane segments and use reciprocal-scaled tables, so the
comparison is <= 1e-13 relative, not bit-exact; both layouts (axis 0 = COL, axis 1 = ROW),
ragged batches, lengths 64 ... 3070, operands split into segments like the slabs of peer GPUs."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-13
SIZES = [(64, 5), (130, 37), (512, 64), (2046, 41), (2048, 12), (3070, 9)]


def _mods():
    import torch
    from pypde_b200 import _cabi as C
    from pypde_b200 import passes as PS
    return torch, C, PS


def _arr(rng, n, nseq, axis):
    """(host array with the sequences along `axis`)"""
    a = rng.standard_normal((n, nseq) if axis == 0 else (nseq, n))
    return a


def _run(PS, axis, n, build, nseq):
    L = PS.PassLaunch(axis, n, PS.TableCache())
    build(L.job(nseq))
    L.run()


def _dev(torch, a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def _along0(fn, a, axis):
    """apply an axis-0 oracle function along `axis`"""
    return fn(a) if axis == 0 else np.ascontiguousarray(fn(np.ascontiguousarray(a.T)).T)


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("n,nseq", SIZES)
@pytest.mark.parametrize("kind", ["CD", "CN"])
def test_stencil_diff_from_cheb(axis, n, nseq, kind):
    """S v, d/dx (S v) / scale and the inverse map (S^T S)^-1 S^T u in one pass each."""
    torch, C, PS = _mods()
    from oracle import kernels as K
    from oracle import pypde_port as P
    from pypde_b200.bases.spectralbase import Base
    rng = np.random.default_rng(n + axis)
    ob, b = P.Basis(n, kind), Base(n, kind)
    M = n - 2
    v = _arr(rng, M, nseq, axis)
    want_u = _along0(ob.to_cheb, v, axis)
    want_d = _along0(lambda x: K.diff_2d(ob.to_cheb(x)), v, axis) / 0.5
    dv = _dev(torch, v)
    shape = (n, nseq) if axis == 0 else (nseq, n)
    du, dd = torch.zeros(shape, dtype=torch.float64, device="cuda"), torch.zeros(shape, dtype=torch.float64, device="cuda")
    _run(PS, axis, n, lambda p: p.load(dv).stencil(b).store(du).diff(0.5).store(dd), nseq)
    assert rel_l2(du.cpu().numpy(), want_u) < 1e-15
    assert rel_l2(dd.cpu().numpy(), want_d) < TOL
    # inverse map
    u = _arr(rng, n, nseq, axis)
    want_v = _along0(ob.from_cheb, u, axis)
    dvv = torch.zeros(tuple(v.shape), dtype=torch.float64, device="cuda")
    uu = _dev(torch, u)
    _run(PS, axis, n, lambda p: p.load(uu).from_cheb(b).store(dvv), nseq)
    assert rel_l2(dvv.cpu().numpy(), want_v) < TOL


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("n,nseq", SIZES)
def test_band_and_fdma(axis, n, nseq):
    """Helmholtz along one axis: x = A^-1 B r with the reference's plans (templates/hholtz.py:60-91)."""
    torch, C, PS = _mods()
    from oracle import kernels as K
    from pypde_b200.bases.spectralbase import Base
    from pypde_b200.templates.hholtz import solverplan_hholtz2d_adi
    rng = np.random.default_rng(7 * n + axis)
    sol = solverplan_hholtz2d_adi([Base(n, "CD"), Base(n, "CN")], lam=3e-4, scale=(0.5, 0.5))
    for which in (0, 1):
        rhs_plan, lhs_plan = sol.plan_for_rhs[which], sol.plan_for_lhs[which]
        band = rhs_plan.band
        r = _arr(rng, n, nseq, axis)
        # oracle: banded product along the axis, then the Fortran fdma sweeps
        A = np.zeros((band.n_out, band.n_in))
        for d, o in enumerate(band.offsets):
            for row in range(band.n_out):
                if 0 <= row + o < band.n_in:
                    A[row, row + o] = band.host[d, row]
        br = _along0(lambda x: A @ x, r, axis)
        want_b = br.copy()
        x = np.ascontiguousarray(br)
        K.solve_fdma_2d(lhs_plan.l, lhs_plan.d, lhs_plan.u1, lhs_plan.u2, x, axis)
        dr = _dev(torch, r)
        out_b = torch.zeros(tuple(br.shape), dtype=torch.float64, device="cuda")
        out_x = torch.zeros(tuple(br.shape), dtype=torch.float64, device="cuda")
        _run(PS, axis, n, lambda p: p.load(dr).band(band).store(out_b).fdma(lhs_plan).store(out_x), nseq)
        assert rel_l2(out_b.cpu().numpy(), want_b) < 1e-15
        assert rel_l2(out_x.cpu().numpy(), x) < TOL, (which, rel_l2(out_x.cpu().numpy(), x))


@pytest.mark.parametrize("n,m", [(62, 62), (126, 70), (510, 510), (2046, 300)])
def test_poisson_columns(n, m):
    """(A + lam_i C) x_i = b_i per column with the per-column tables in segment order (fdma.f90:146-195),
    including the singular column (lam = 0: row / column 0 dropped, x_0 = 0)."""
    torch, C, PS = _mods()
    from oracle import kernels as K
    from oracle import pypde_port as P
    from pypde_b200.bases.spectralbase import Base
    from pypde_b200.templates.poisson import solverplan_poisson2d
    N = n + 2
    sol = solverplan_poisson2d([Base(N, "CN"), Base(N, "CN")], singular=True, scale=(0.5, 0.5))
    plan = sol.plan_for_lhs[0]
    lam = np.asarray(plan.alpha, dtype=float)[:m]
    from pypde_b200 import ops
    sub = ops.PoissonPlan(plan._Ad, plan._Cd, lam, plan.singular)
    o = P.PoissonEig([P.Basis(N, "CN"), P.Basis(N, "CN")], singular=True, scale=(0.5, 0.5))
    rng = np.random.default_rng(n)
    b = rng.standard_normal((n, m))
    want = b.copy()
    Ax = o.Ax.toarray() if hasattr(o.Ax, "toarray") else np.asarray(o.Ax)
    Cx = o.Cx.toarray() if hasattr(o.Cx, "toarray") else np.asarray(o.Cx)
    K.solve_fdma_type2(Ax, Cx, o.wy[:m], want, 0, True)
    assert np.array_equal(o.wy[:m], lam)
    lg = PS.lg_for(n)
    tab = PS.PoissonTables(sub, lg)
    x = _dev(torch, b)
    _run(PS, 0, n, lambda p: p.load(x).poisson(tab).store(x), m)
    got = x.cpu().numpy()
    assert got[0, 0] == 0.0
    # column-wise: every system has its own conditioning
    err = np.linalg.norm(got - want, axis=0) / np.linalg.norm(want, axis=0)
    assert err.max() < 1e-11 and np.median(err) < 1e-14, (err.max(), np.median(err))


@pytest.mark.parametrize("axis", [0, 1])
def test_axpy_scale_setz_segments(axis):
    """data movement: axpy (scaled, stencil-on-load), scale, setz0, store of one sequence, and operands split
    into segments along the sequence (what the peer slabs of the distributed transposes look like)."""
    torch, C, PS = _mods()
    from oracle import pypde_port as P
    from pypde_b200.bases.spectralbase import Base
    n, nseq = 260, 11
    rng = np.random.default_rng(3)
    a, g, h = (_arr(rng, n, nseq, axis) for _ in range(3))
    b = Base(n, "CN")
    ob = P.Basis(n, "CN")
    gm = _arr(rng, n - 2, nseq, axis)
    want = 0.25 * (a * 2.0 + 3.0 * g) - 1.5 * h + _along0(ob.to_cheb, gm, axis)
    da, dg, dh, dgm = (_dev(torch, t) for t in (a, g, h, gm))
    out = torch.zeros_like(da)
    L = PS.PassLaunch(axis, n, PS.TableCache())
    st = L.tables.stencil_elem(b)
    # the stencil-on-load operand has n - 2 valid entries; its image has n
    gm_pad = torch.zeros_like(da)
    if axis == 0:
        gm_pad[: n - 2] = dgm
    else:
        gm_pad[:, : n - 2] = dgm
    L.job(nseq).load(da).scale(2.0).axpy(3.0, dg).axpy(-1.5, dh, scale_buf=0.25).axpy(1.0, gm_pad, stencil=st).store(out)
    L.run()
    assert rel_l2(out.cpu().numpy(), want) < 1e-15
    # segments: the same array presented as three pieces with their own base pointers
    cuts = [0, 64, 200, n]
    pieces = []
    for s in range(3):
        sl = (slice(cuts[s], cuts[s + 1]), slice(None)) if axis == 0 else (slice(None), slice(cuts[s], cuts[s + 1]))
        pieces.append(da[sl].contiguous())
    opnd = PS.Operand([p.data_ptr() for p in pieces], [p.stride(0) for p in pieces], cuts, n, keep=tuple(pieces))
    outs = [torch.zeros_like(p) for p in pieces]
    oop = PS.Operand([p.data_ptr() for p in outs], [p.stride(0) for p in outs], cuts, n, keep=tuple(outs))
    L = PS.PassLaunch(axis, n, PS.TableCache())
    L.job(nseq, seq0=100).load(opnd).setz0(103).store(oop).scale(-1.0).store(out, only_seq=105)
    L.run()
    got = torch.cat(outs, dim=axis).cpu().numpy()
    ref = a.copy()
    if axis == 0:
        ref[0, 3] = 0.0
    else:
        ref[3, 0] = 0.0
    assert np.array_equal(got, ref)
    o2 = out.cpu().numpy()
    seq5 = o2[:, 5] if axis == 0 else o2[5]
    assert np.array_equal(seq5, -(ref[:, 5] if axis == 0 else ref[5]))
    other = o2[:, 4] if axis == 0 else o2[4]
    assert rel_l2(other, (want[:, 4] if axis == 0 else want[4])) < 1e-15       # untouched by the one-sequence store


def test_many_jobs_one_launch():
    """One launch, 40 jobs with different programs (the ensemble / multi-field form)."""
    torch, C, PS = _mods()
    from oracle import kernels as K
    n, nseq = 128, 9
    rng = np.random.default_rng(11)
    L = PS.PassLaunch(1, n, PS.TableCache())
    ins, outs = [], []
    for j in range(40):
        a = rng.standard_normal((nseq, n))
        d, o = _dev(torch, a), torch.zeros((nseq, n), dtype=torch.float64, device="cuda")
        ins.append(a)
        outs.append(o)
        p = L.job(nseq).load(d)
        if j % 2:
            p.diff(0.5)
        p.scale(float(j)).store(o)
    L.run()
    for j in range(40):
        want = ins[j]
        if j % 2:
            want = np.ascontiguousarray(K.diff_2d(np.ascontiguousarray(want.T)).T) / 0.5
        assert rel_l2(outs[j].cpu().numpy(), want * j) < TOL


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("n,nseq", [(66, 7), (2048, 19)])
def test_lincomb(axis, n, nseq):
    """buffer = sum_k coef_k G_k with operands of different valid lengths, then accumulated onto the buffer."""
    torch, C, PS = _mods()
    rng = np.random.default_rng(n + axis)
    hs = [_arr(rng, n - 2 * (k % 2), nseq, axis) for k in range(5)]
    ds = [_dev(torch, h) for h in hs]
    coef = [1.0, -0.25, 3.0, 1e-3, -7.5]
    shape = (n, nseq) if axis == 0 else (nseq, n)
    out1 = torch.zeros(shape, dtype=torch.float64, device="cuda")
    out2 = torch.zeros(shape, dtype=torch.float64, device="cuda")
    _run(PS, axis, n, lambda p: p.lincomb(list(zip(coef, ds))).store(out1)
         .lincomb([(2.0, ds[0]), (-1.0, ds[2])], accumulate=True).store(out2), nseq)
    want = np.zeros(shape)
    for cf, h in zip(coef, hs):
        if axis == 0:
            want[: h.shape[0]] += cf * h
        else:
            want[:, : h.shape[1]] += cf * h
    assert rel_l2(out1.cpu().numpy(), want) < 1e-15
    want2 = want.copy()
    want2 += 2.0 * hs[0] - hs[2]
    assert rel_l2(out2.cpu().numpy(), want2) < 1e-15


def test_batched_gemm_dct_products():
    """pde_gemm_f64_batched / pde_dct1_batched / pde_conv_products_members (one launch for many members) against
    the per-array entry points and NumPy."""
    import ctypes
    torch, C, PS = _mods()
    from pypde_b200 import ops
    rng = np.random.default_rng(2)
    nb, m, n, k = 7, 50, 38, 44
    A = [torch.as_tensor(rng.standard_normal((m, k)), device="cuda") for _ in range(nb)]
    B = torch.as_tensor(rng.standard_normal((n, k)), device="cuda")
    Cs = [torch.zeros((m, n), dtype=torch.float64, device="cuda") for _ in range(nb)]
    pa = torch.tensor([a.data_ptr() for a in A], dtype=torch.int64, device="cuda")
    pc = torch.tensor([c.data_ptr() for c in Cs], dtype=torch.int64, device="cuda")
    C.check(C.lib().pde_gemm_f64_batched(1, None, ctypes.c_void_p(pa.data_ptr()), k, C.p(B), None, k,
                                         ctypes.c_void_p(pc.data_ptr()), n, m, n, k, nb, 1, C.stream()))
    for a, c in zip(A, Cs):
        assert rel_l2(c.cpu().numpy(), a.cpu().numpy() @ B.cpu().numpy().T) < 1e-14
        assert torch.equal(c, ops.gemm(a, B, transB=True))
    # dense DCT, both axes, truncated / padded
    L = 97
    plan = ops.DctPlan.get(L)
    assert plan.algo == 1
    for axis in (0, 1):
        xs = [torch.as_tensor(rng.standard_normal((64, 30) if axis == 0 else (30, 64)), device="cuda") for _ in range(5)]
        ys = [torch.zeros((L, 30) if axis == 0 else (30, L), dtype=torch.float64, device="cuda") for _ in range(5)]
        px = torch.tensor([x.data_ptr() for x in xs], dtype=torch.int64, device="cuda")
        py = torch.tensor([y.data_ptr() for y in ys], dtype=torch.int64, device="cuda")
        C.check(C.lib().pde_dct1_batched(plan.handle, ops.BWD, 5, ctypes.c_void_p(px.data_ptr()), xs[0].stride(0), 64,
                                         ctypes.c_void_p(py.data_ptr()), ys[0].stride(0), L, 30, axis, 1, C.stream()))
        for x, y in zip(xs, ys):
            assert torch.equal(y, ops.dct1(plan, ops.BWD, x, axis=axis, n_out=L))
    # products of 3 members with a shared boundary term
    nm, npts = 3, 1000
    arrs = [torch.as_tensor(rng.standard_normal((nm, npts)), device="cuda") for _ in range(10)]
    u, w, uo, wo, dxU, dzU, dxV, dzV, dxT, dzT = arrs
    dTbc = torch.as_tensor(rng.standard_normal(npts), device="cuda")
    ref = [t.clone() for t in (dxU, dxV, dxT)]
    for mm in range(nm):
        C.check(C.lib().pde_conv_products(npts, 0.4, -0.3, C.p(u[mm]), C.p(w[mm]), C.p(uo[mm]), C.p(wo[mm]), C.p(ref[0][mm]),
                                          C.p(dzU[mm]), C.p(ref[1][mm]), C.p(dzV[mm]), C.p(ref[2][mm]), C.p(dzT[mm]),
                                          C.p(dTbc), C.stream()))
    C.check(C.lib().pde_conv_products_members(npts, nm, npts, 0.4, -0.3, C.p(u), C.p(w), C.p(uo), C.p(wo), C.p(dxU), C.p(dzU),
                                              C.p(dxV), C.p(dzV), C.p(dxT), C.p(dzT), C.p(dTbc), C.stream()))
    for a, b in zip((dxU, dxV, dxT), ref):
        assert torch.equal(a, b)
