"""GPU parity tests: every C-ABI primitive (called through the reference-facing Python
classes) against the golden fixtures and the CPU oracle on the same seeded inputs.

Tolerances: recurrence / banded kernels are compiled without FMA contraction and must be
BIT-EXACT; the DCT-I (different summation order than pocketfft) and the dense contractions
must agree to <= 1e-13 relative (L2), well inside the 1e-12 budget of the north star."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

DCT_TOL = 2e-14


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda")


def T(a, dev):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def H(t):
    return t.detach().cpu().numpy()


def test_library_reports_blackwell():
    from pypde_b200 import _cabi
    import ctypes
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _cabi.check(_cabi.lib().pde_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)))
    assert sm.value > 0 and ma.value >= 10, "built for sm_100a only"


@pytest.mark.parametrize("N", [20, 33, 64])
@pytest.mark.parametrize("kind", ["CH", "CD", "CN"])
def test_basis_against_golden(golden_prim, dev, kind, N):
    from pypde_b200 import Base
    g = golden_prim
    b = Base(N, kind)
    key = "%s%d" % (kind, N)
    f, c = T(g[key + "_f"], dev), T(g[key + "_c"], dev)
    assert rel_l2(H(b.forward_fft(f)), g[key + "_forward"]) < DCT_TOL
    assert rel_l2(H(b.backward_fft(c)), g[key + "_backward"]) < DCT_TOL
    for order in (1, 2):
        assert np.array_equal(H(b.derivative(c, order)), g[key + "_deriv%d" % order])
    if kind != "CH":
        assert np.array_equal(H(b.to_chebyshev(c)), g[key + "_to_cheb"])
        assert np.array_equal(H(b.from_chebyshev(T(g[key + "_u"], dev))), g[key + "_from_cheb"])
    # NumPy in -> NumPy out (drop-in behaviour of examples/transform1d.py)
    out = b.forward_fft(g[key + "_f"])
    assert isinstance(out, np.ndarray) and rel_l2(out, g[key + "_forward"]) < DCT_TOL


@pytest.mark.parametrize("L", [2, 3, 5, 17, 64, 96, 128, 192, 257])
def test_raw_dct1_against_scipy_golden(golden_prim, dev, L):
    from pypde_b200 import ops
    g = golden_prim
    plan = ops.DctPlan.get(L)
    y = ops.dct1(plan, ops.RAW, T(g["dct1_%d_x" % L], dev), axis=0)
    assert rel_l2(H(y), g["dct1_%d_y" % L]) < DCT_TOL


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("kind", ["CH", "CD", "CN"])
def test_axis1_equals_axis0_transposed(dev, kind, axis):
    """axis-1 (row) kernels against the oracle applied to the transposed problem; ragged sizes."""
    from pypde_b200 import Base
    from oracle import pypde_port as P
    rng = np.random.default_rng(11)
    N, nb = 45, 37
    b, o = Base(N, kind), P.Basis(N, kind)
    shp = lambda n: (n, nb) if axis == 0 else (nb, n)
    tr = (lambda a: a) if axis == 0 else (lambda a: a.T)
    f, c, u = rng.standard_normal(shp(N)), rng.standard_normal(shp(b.M)), rng.standard_normal(shp(N))
    assert rel_l2(H(b.forward_fft(T(f, dev), axis=axis)), tr(o.forward(tr(f)))) < DCT_TOL
    assert rel_l2(H(b.backward_fft(T(c, dev), axis=axis)), tr(o.backward(tr(c).copy()))) < DCT_TOL
    for order in (1, 2):
        assert np.array_equal(H(b.derivative(T(c, dev), order, axis=axis, div=0.75 ** order)),
                              tr(o.deriv(tr(c), order) / 0.75 ** order))
    if kind != "CH":
        assert np.array_equal(H(b.to_chebyshev(T(c, dev), axis=axis)), tr(o.to_cheb(tr(c))))
        assert np.array_equal(H(b.from_chebyshev(T(u, dev), axis=axis)), tr(o.from_cheb(tr(u))))


@pytest.mark.parametrize("axis", [0, 1])
def test_banded_solvers_bit_exact(golden_prim, dev, axis):
    from pypde_b200 import PlanLHS
    g = golden_prim
    b = g["fdma_b%d" % axis]
    plan = PlanLHS(g["fdma_A"], ndim=2, axis=axis, method="fdma")
    x = T(b, dev)
    out = plan.solve(x)
    assert out.data_ptr() == x.data_ptr(), "fdma solves in place like the reference"
    assert np.array_equal(H(x), g["fdma_x%d" % axis])
    A2 = np.diag(g["twodma_d"]) + np.diag(g["twodma_u"], 2)
    x = PlanLHS(A2, ndim=2, axis=axis, method="twodma").solve(T(b, dev))
    assert np.array_equal(H(x), g["twodma_x%d" % axis])
    # 1-D
    b1 = np.ascontiguousarray(b[:, 0] if axis == 0 else b[0, :])
    x1 = PlanLHS(g["fdma_A"], ndim=1, axis=0, method="fdma").solve(T(b1, dev))
    assert np.array_equal(H(x1), g["fdma_x%d" % axis][:, 0] if axis == 0 else g["fdma_x%d" % axis][0, :])
    # numpy in -> solved in place on the caller's array
    bn = b.copy()
    plan.solve(bn)
    assert np.array_equal(bn, g["fdma_x%d" % axis])


@pytest.mark.parametrize("n,batch", [(5, 1), (6, 3), (64, 1000), (2046, 40), (3070, 9)])
@pytest.mark.parametrize("axis", [0, 1])
def test_fdma_sizes_vs_oracle(dev, n, batch, axis):
    """Minimum, ragged and maximum (row-tile limit) sizes of the 4-diagonal solve."""
    from pypde_b200 import ops, _cabi as C
    from oracle import kernels as K, pypde_port as P
    rng = np.random.default_rng(n * 7 + batch)
    A = np.zeros((n, n))
    for off in (-2, 0, 2, 4):
        if n - abs(off) > 0:
            A += np.diag(rng.standard_normal(n - abs(off)) * 0.3 + (3.0 if off == 0 else 0.0), off)
    l, d, u1, u2 = P.fdma_lu(A)
    b = rng.standard_normal((n, batch) if axis == 0 else (batch, n))
    ref = K.solve_fdma_2d(l, d, u1, u2, b.copy(), axis)
    tabs = [C.upload(t if t.size else np.zeros(1)) for t in (l, d, u1, u2)]
    x = ops.fdma_solve(*tabs, T(b, dev), axis=axis)
    assert np.array_equal(H(x), ref)


def test_banded_product_and_dense_product(dev):
    from pypde_b200 import PlanRHS
    import scipy.sparse as sp
    rng = np.random.default_rng(2)
    n_out, n_in, nb = 38, 40, 17
    A = sp.diags([rng.standard_normal(n_out), rng.standard_normal(n_out), rng.standard_normal(n_out - 2)],
                 [0, 2, 4], shape=(n_out, n_in), format="csr")
    for axis in (0, 1):
        b = rng.standard_normal((n_in, nb) if axis == 0 else (nb, n_in))
        ref = A @ b if axis == 0 else (A @ b.T).T
        assert np.array_equal(H(PlanRHS(A, ndim=2, axis=axis).solve(T(b, dev))), ref)
        D = rng.standard_normal((n_out, n_in))          # dense -> tensor-core contraction
        ref = D @ b if axis == 0 else b @ D.T
        assert rel_l2(H(PlanRHS(D, ndim=2, axis=axis).solve(T(b, dev))), ref) < 1e-14


# the last five sizes select, on 148 SMs, the 128x56, 128x48, 128x64, 64x64 tile shapes and the unaligned
# (scalar cp.async) path of the 128x56 one (pde::gemm_f64 picks the shape that fills whole waves)
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (128, 64, 16), (129, 65, 17), (300, 257, 511), (62, 62, 64),
                                   (1874, 1874, 33), (1538, 1538, 40), (2050, 2050, 20), (1090, 1090, 24),
                                   (1875, 1877, 9)])
@pytest.mark.parametrize("transB", [False, True])
def test_gemm_f64(dev, m, n, k, transB):
    import torch
    from pypde_b200 import ops
    rng = np.random.default_rng(m + n + k)
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((n, k) if transB else (k, n))
    ref = A @ (B.T if transB else B)
    assert rel_l2(H(ops.gemm(T(A, dev), T(B, dev), transB)), ref) < 1e-14
    # fp64 torch reference of the same op (cuBLAS)
    tref = T(A, dev) @ (T(B, dev).T if transB else T(B, dev))
    assert rel_l2(H(ops.gemm(T(A, dev), T(B, dev), transB)), H(tref)) < 1e-14


def test_fields_against_golden(golden_fs, dev):
    from pypde_b200 import Base, Field, FieldBC, grad, galerkin_to_cheby, cheby_to_galerkin
    g = golden_fs
    N0, N1 = 40, 20
    for kx, ky in (("CD", "CN"), ("CH", "CH"), ("CN", "CD")):
        fld = Field([Base(N0, kx, dealias=3 / 2), Base(N1, ky, dealias=3 / 2)])
        key = "f2d_%s%s" % (kx, ky)
        assert rel_l2(H(fld.forward(T(g[key + "_v"], dev))), g[key + "_fwd"]) < DCT_TOL
        assert rel_l2(H(fld.backward(T(g[key + "_vhat"], dev))), g[key + "_bwd"]) < DCT_TOL
        bwd = fld.dealias.backward(T(g[key + "_vhat"], dev))
        assert tuple(bwd.shape) == (60, 30)
        assert rel_l2(H(bwd), g[key + "_dbwd"]) < DCT_TOL
        fwd = fld.dealias.forward(T(g[key + "_dbwd"] * g[key + "_dbwd"], dev))
        assert tuple(fwd.shape) == tuple(g[key + "_dfwd"].shape)
        assert rel_l2(H(fwd), g[key + "_dfwd"]) < 10 * DCT_TOL
        fld.vhat = g[key + "_vhat"]
        for deriv in ((1, 0), (0, 1), (2, 0), (1, 1)):
            assert np.array_equal(H(grad(fld, deriv, scale=(0.75, 0.5))), g[key + "_grad%d%d" % deriv])
        if kx != "CH":
            assert np.array_equal(H(galerkin_to_cheby(T(g[key + "_vhat"], dev), fld)), g[key + "_g2c"])
            assert np.array_equal(H(cheby_to_galerkin(T(g[key + "_cheb"], dev), fld)), g[key + "_c2g"])
    for axis in (0, 1):
        fb = FieldBC([Base(N0, "CD"), Base(N1, "CN")], axis=axis)
        fb.add_bc(g["fbc%d_bc" % axis])
        assert rel_l2(H(fb.v), g["fbc%d_v" % axis]) < DCT_TOL
        assert rel_l2(H(fb.vhat), g["fbc%d_vhat" % axis]) < 10 * DCT_TOL


def test_solver_templates_against_golden(golden_fs, dev):
    from pypde_b200 import Base
    from pypde_b200.templates.hholtz import solverplan_hholtz1d, solverplan_hholtz2d_adi
    from pypde_b200.templates.poisson import solverplan_poisson1d, solverplan_poisson2d
    g = golden_fs
    N0, N1 = 50, 40
    for kx, ky in (("CD", "CN"), ("CN", "CN"), ("CD", "CD")):
        key = kx + ky
        bases = [Base(N0, kx), Base(N1, ky)]
        s = solverplan_hholtz2d_adi(bases, lam=0.013, scale=(0.75, 0.5))
        r = s.solve_rhs(T(g["hh_" + key + "_rhs"], dev))
        r += s.solve_old(T(g["hh_" + key + "_old"], dev))
        assert np.array_equal(H(s.solve_lhs(r)), g["hh_" + key + "_x"]), "Helmholtz ADI is bit-exact"
        p = solverplan_poisson2d(bases, singular=(key == "CNCN"), scale=(0.75, 0.5))
        x = p.solve_lhs(p.solve_rhs(T(g["hh_" + key + "_rhs"], dev)))
        assert rel_l2(H(x), g["po_" + key + "_x"]) < 1e-9      # golden from another host's LAPACK
    f, uo = g["t1d_f"], g["t1d_old"]
    for kind in ("CD", "CN"):
        s = solverplan_hholtz1d([Base(50, kind)], lam=0.02)
        r = s.solve_rhs(T(f, dev))
        r += s.solve_old(T(uo, dev))
        assert np.array_equal(H(s.solve_lhs(r)), g["hh1_%s_x" % kind])
        p = solverplan_poisson1d([Base(50, kind)], singular=(kind == "CN"))
        assert np.array_equal(H(p.solve_lhs(p.solve_rhs(T(f, dev)))), g["po1_%s_x" % kind])


@pytest.mark.parametrize("kinds", [("CN", "CN"), ("CD", "CD")])
def test_poisson_against_oracle_same_host(dev, kinds):
    """Same-host comparison (identical LAPACK setup on both sides): tight tolerance."""
    from pypde_b200 import Base
    from pypde_b200.templates.poisson import solverplan_poisson2d
    from oracle import pypde_port as P
    rng = np.random.default_rng(4)
    N0, N1 = 66, 50
    rhs = rng.standard_normal((N0, N1))
    sing = kinds == ("CN", "CN")
    o = P.PoissonEig([P.Basis(N0, kinds[0]), P.Basis(N1, kinds[1])], singular=sing, scale=(0.5, 0.5))
    ref = o.solve_lhs(o.solve_rhs(rhs))
    p = solverplan_poisson2d([Base(N0, kinds[0]), Base(N1, kinds[1])], singular=sing, scale=(0.5, 0.5))
    x = p.solve_lhs(p.solve_rhs(T(rhs, dev)))
    assert rel_l2(H(x), ref) < 1e-12


def test_poisson_analytic(dev):
    """reference test/test_poisson2d.py: cos(pi x/2) cos(pi y/2) forcing, rtol 1e-3."""
    from pypde_b200 import Base, Field
    from pypde_b200.templates.poisson import solverplan_poisson2d
    N0, N1 = 50, 40
    fld = Field([Base(N0, "CD"), Base(N1, "CD")])
    xx, yy = np.meshgrid(fld.x, fld.y, indexing="ij")
    arg = np.pi / 2
    sol = np.cos(arg * xx) * np.cos(arg * yy)
    f = -2 * arg ** 2 * sol
    ch = Field([Base(N0, "CH"), Base(N1, "CH")])
    fhat = ch.forward(T(f, dev))
    p = solverplan_poisson2d(fld.xs, singular=False)
    fld.vhat = p.solve_lhs(p.solve_rhs(fhat))
    fld.backward()
    assert np.allclose(H(fld.v), sol, rtol=1e-3, atol=1e-6)


def test_transform_roundtrip_large(dev):
    """Size-independent property at transform-sweep sizes: backward(forward(f)) == f."""
    import torch
    from pypde_b200 import Base
    for N, nb in ((512, 300), (1024, 64)):
        b = Base(N, "CH")
        f = torch.randn((N, nb), dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(0))
        back = b.backward_fft(b.forward_fft(f))
        assert float(torch.linalg.norm(back - f) / torch.linalg.norm(f)) < 1e-13


@pytest.mark.parametrize("L", [3, 5, 7, 13, 17, 97, 129, 161, 193, 769, 1537, 3073])
@pytest.mark.parametrize("axis", [0, 1])
def test_fft_dct_vs_oracle(dev, L, axis):
    """Shared-memory FFT DCT-I (algo 2) against scipy's pocketfft through the oracle, all three
    modes, with zero padding (n_in < L) and truncation (n_out < L); ragged batch sizes."""
    from pypde_b200 import ops
    from oracle import pypde_port as P
    rng = np.random.default_rng(L)
    plan = ops.DctPlan(L, algo=2)
    assert plan.algo == 2
    nb = 37 if L < 1000 else 11
    o = P.Basis(L, "CH")
    tr = (lambda a: a) if axis == 0 else (lambda a: np.ascontiguousarray(a.T))
    x = rng.standard_normal((L, nb))
    from scipy.fftpack import dctn
    tol = 5e-15 * max(1.0, np.log2(L))
    got = ops.dct1(plan, ops.RAW, T(tr(x), dev), axis=axis)
    assert rel_l2(tr(H(got)), dctn(x, type=1, axes=(0,))) < tol
    got = ops.dct1(plan, ops.FWD, T(tr(x), dev), axis=axis)
    assert rel_l2(tr(H(got)), o.forward(x)) < tol
    got = ops.dct1(plan, ops.BWD, T(tr(x), dev), axis=axis)
    assert rel_l2(tr(H(got)), o.backward(x.copy())) < tol
    if L >= 7:
        n_in, n_out = (2 * L) // 3, (2 * L) // 3 + 1
        xp = np.zeros((L, nb))
        xp[:n_in] = x[:n_in]
        got = ops.dct1(plan, ops.BWD, T(tr(x[:n_in]), dev), axis=axis)
        assert rel_l2(tr(H(got)), o.backward(xp.copy())) < tol
        got = ops.dct1(plan, ops.FWD, T(tr(x), dev), axis=axis, n_out=n_out)
        assert tuple(tr(H(got)).shape) == (n_out, nb)
        assert rel_l2(tr(H(got)), o.forward(x)[:n_out]) < tol


@pytest.mark.parametrize("L", [2049, 3073, 4097])
def test_fft_dct_rows_persistent_tma(dev, L):
    """Axis-1 transforms of rows with an even pitch (16-byte aligned rows) run in the persistent kernel whose input
    rows are staged by the bulk-copy engine (k_dct_row_tma): all modes, odd and even input lengths (the odd tail
    element travels separately), zero padding, truncation, and more rows than resident CTAs (3 x 148)."""
    import torch
    from pypde_b200 import ops
    from oracle import pypde_port as P
    rng = np.random.default_rng(L)
    plan = ops.DctPlan(L, algo=2)
    o = P.Basis(L, "CH")
    tol = 5e-15 * np.log2(L)
    nb = 1000
    x = rng.standard_normal((nb, L))
    for n_in in (L, (2 * L) // 3, (2 * L) // 3 + 1, 2, 1):
        ld = n_in + 2
        ld += ld & 1                                          # an even pitch > n_in
        buf = torch.full((nb, ld), float("nan"), dtype=torch.float64, device=dev)     # pad columns must never be read
        buf[:, :n_in] = T(x[:, :n_in], dev)
        xin = buf[:, :n_in]
        assert xin.stride(0) % 2 == 0 and xin.data_ptr() % 16 == 0
        xp = np.zeros((L, nb))
        xp[:n_in] = x[:, :n_in].T
        for mode, ref in ((ops.BWD, lambda a: o.backward(a.copy())), (ops.FWD, o.forward)):
            if mode == ops.FWD and n_in != L:
                continue
            for n_out in (L, (2 * L) // 3 + 1):
                got = H(ops.dct1(plan, mode, xin, axis=1, n_out=n_out))
                assert got.shape == (nb, n_out)
                assert rel_l2(got.T, ref(xp)[:n_out]) < tol, (L, n_in, mode, n_out)


def test_fft_dct_roundtrip_full_size(dev):
    """BASELINE-size property: backward(forward(f)) == f on the 3073-point dealias grid."""
    import torch
    from pypde_b200 import ops
    plan = ops.DctPlan.get(3073)
    assert plan.algo == 2
    f = torch.randn((3073, 2048), dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(1))
    for axis in (0, 1):
        g = f if axis == 0 else f.T.contiguous()
        back = ops.dct1(plan, ops.BWD, ops.dct1(plan, ops.FWD, g, axis=axis), axis=axis)
        assert float(torch.linalg.norm(back - g) / torch.linalg.norm(g)) < 1e-14


@pytest.mark.parametrize("L", [128, 200, 256, 512, 1000, 1024, 2048, 4096])
@pytest.mark.parametrize("axis", [0, 1])
def test_bluestein_dct_vs_oracle(dev, L, axis):
    """Chirp-z DCT-I (algo 3) for lengths whose L-1 is odd / has large prime factors (the Base(N, "CH")
    sizes with N a power of two: 127 prime, 2047 = 23*89, 4095 = 3^2*5*7*13) against pocketfft."""
    from pypde_b200 import ops
    from oracle import pypde_port as P
    from scipy.fftpack import dctn
    rng = np.random.default_rng(L + 1)
    plan = ops.DctPlan(L, algo=3)
    assert plan.algo == 3
    nb = 19 if L <= 1024 else 7
    o = P.Basis(L, "CH")
    tr = (lambda a: a) if axis == 0 else (lambda a: np.ascontiguousarray(a.T))
    x = rng.standard_normal((L, nb))
    tol = 2e-14 * max(1.0, np.log2(L))
    got = ops.dct1(plan, ops.RAW, T(tr(x), dev), axis=axis)
    assert rel_l2(tr(H(got)), dctn(x, type=1, axes=(0,))) < tol
    got = ops.dct1(plan, ops.FWD, T(tr(x), dev), axis=axis)
    assert rel_l2(tr(H(got)), o.forward(x)) < tol
    got = ops.dct1(plan, ops.BWD, T(tr(x), dev), axis=axis)
    assert rel_l2(tr(H(got)), o.backward(x.copy())) < tol
    n_in, n_out = (2 * L) // 3, (2 * L) // 3 + 1
    xp = np.zeros((L, nb))
    xp[:n_in] = x[:n_in]
    got = ops.dct1(plan, ops.BWD, T(tr(x[:n_in]), dev), axis=axis, n_out=n_out)
    assert rel_l2(tr(H(got)), o.backward(xp.copy())[:n_out]) < tol


def test_auto_algo_selection(dev):
    from pypde_b200 import ops
    assert ops.DctPlan(64).algo == 1          # short and odd L-1: dense DMMA matrix
    assert ops.DctPlan(97).algo == 1          # short: dense wins (1.8 TB/s vs 0.8 TB/s measured)
    assert ops.DctPlan(769).algo == 2         # L-1 = 768: shared-memory FFT
    assert ops.DctPlan(3073).algo == 2
    assert ops.DctPlan(2048).algo == 3        # L-1 = 2047 = 23 * 89: Bluestein
    assert ops.DctPlan(6144).algo == 1        # L-1 = 6143 prime, M = 16384 does not fit: dense fallback


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("scales", [(0.5, 0.5, 0.5), (0.25, 2.0, 1.0), (0.5, 0.75, 1.0)])
def test_batched_diff_sweep_bit_exact(dev, axis, scales):
    """pde_sweep(DIFF) with several jobs of different widths in one launch (the stepper's batched form).
    Power-of-two scales take the multiply-by-reciprocal instantiation (x / 2^k == x * 2^-k), any other
    scale the correctly rounded division sequence: both must equal diff_2d(c) / scale bit for bit
    (differentiate_cheby.f90:28-53, field_operations.py:40-45)."""
    import torch
    from pypde_b200 import _cabi as C
    from oracle import kernels as K
    rng = np.random.default_rng(5 + axis)
    n = 203                                  # several interior chunks + ragged edges
    widths = (70, 33, 1)
    arr = (C.SweepJob * len(widths))()
    keep, refs = [], []
    for k, (w, sc) in enumerate(zip(widths, scales)):
        c = rng.standard_normal((n, w) if axis == 0 else (w, n))
        ref = K.diff_2d(c if axis == 0 else np.ascontiguousarray(c.T))
        ref = ref if axis == 0 else ref.T
        refs.append(ref / sc if sc != 1.0 else ref)
        x, y = T(c, dev), torch.full(c.shape, np.nan, dtype=torch.float64, device=dev)
        keep += [x, y]
        j = arr[k]
        j.inp[0], j.ldin[0], j.out, j.ldout = x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0)
        j.nseq, j.flag, j.sc = w, int(sc != 1.0), sc
    C.check(C.lib().pde_sweep(0, axis, n, len(widths), arr, C.stream()))
    torch.cuda.synchronize()
    for k, ref in enumerate(refs):
        assert np.array_equal(H(keep[2 * k + 1]), ref), "job %d" % k


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("offsets", [(0, 2, 4), (-2, 0, 2, 4), (0,), (-1, 0, 3)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_banded_multi_bit_exact(dev, axis, offsets, accumulate):
    """pde_banded_multi (the stepper's batched PlanRHS products, solver/plans.py:54-74, matrix.py:48-53): jobs of
    different shapes in one launch.  Even sizes with diagonals off0, off0+2, ... take the strip kernels, the
    (-1, 0, 3) pattern and the odd-sized job the generic kernel; all must equal the NumPy expression
    sum_d diag_d * x[row + off_d] (products and sums rounded separately, ascending d) bit for bit."""
    import torch
    from pypde_b200 import _cabi as C
    rng = np.random.default_rng(17 + axis)
    if offsets == (-2, 0, 2, 4):
        shapes = [(70, 70, 130), (34, 34, 64), (6, 6, 2)]          # strip kernels (square B S operators)
    elif offsets == (0, 2, 4):
        shapes = [(70, 72, 130), (34, 36, 64), (2, 4, 258)]        # strip kernels (B(2,2): n_in = n_out + 2)
    else:
        shapes = [(70, 72, 130), (34, 36, 64), (33, 35, 7)]        # generic kernel (odd job / other pattern)
    arr = (C.BandJob * len(shapes))()
    keep, refs = [], []
    for k, (n_out, n_in, batch) in enumerate(shapes):
        diags = rng.standard_normal((len(offsets), n_out))
        diags[:, min(3, n_out - 1)] = 0.0                       # zero coefficients are skipped
        x = rng.standard_normal((n_in, batch) if axis == 0 else (batch, n_in))
        y0 = rng.standard_normal((n_out, batch) if axis == 0 else (batch, n_out))
        xs = x if axis == 0 else x.T
        acc = np.zeros((n_out, batch))
        for d, off in enumerate(offsets):
            for r in range(n_out):
                c = r + off
                if 0 <= c < n_in and diags[d, r] != 0.0:
                    acc[r] = acc[r] + diags[d, r] * xs[c]
        ref = acc if axis == 0 else acc.T
        refs.append(y0 + ref if accumulate else ref)
        dx, dy, dd = T(x, dev), T(y0, dev), T(diags, dev)
        keep += [dx, dy, dd]
        j = arr[k]
        j.diags, j.ndiag = dd.data_ptr(), len(offsets)
        for d, off in enumerate(offsets):
            j.off[d] = off
        j.x, j.ldx, j.n_in = dx.data_ptr(), dx.stride(0), n_in
        j.y, j.ldy, j.n_out = dy.data_ptr(), dy.stride(0), n_out
        j.batch, j.accumulate = batch, int(accumulate)
    C.check(C.lib().pde_banded_multi(axis, len(shapes), arr, C.stream()))
    torch.cuda.synchronize()
    for k, ref in enumerate(refs):
        assert np.array_equal(H(keep[3 * k + 1]), ref), "job %d" % k


def _sweep_jobs(C, jobs):
    """jobs: dicts with in (list), out, tab {slot: tensor}, nseq, flag, sc -> ctypes array (kept alive by the caller)"""
    arr = (C.SweepJob * len(jobs))()
    for k, jb in enumerate(jobs):
        j = arr[k]
        for s, t in enumerate(jb["in"]):
            j.inp[s], j.ldin[s] = t.data_ptr(), t.stride(0)
        j.out, j.ldout = jb["out"].data_ptr(), jb["out"].stride(0)
        for slot, t in jb.get("tab", {}).items():
            j.tab[slot] = t.data_ptr()
        j.nseq, j.flag, j.sc = jb["nseq"], int(jb.get("flag", 0)), float(jb.get("sc", 1.0))
    return arr


@pytest.mark.parametrize("n", [128, 256, 330, 2046])
def test_tiled_row_sweeps_bit_exact(dev, n):
    """Axis-1 sweeps at even lengths >= 128 take the tiled kernel (k_sweep_tile: 32 rows per warp, 64-element tiles,
    both parity chains per thread).  Every operator of the stepper, batched over jobs of ragged widths, against the
    oracle's Fortran restatements: differentiate_cheby.f90:28-53, tdma.f90:55-106 (+ S^T, chebyshev.py:327),
    fdma.f90:40-98.  Bit for bit."""
    import torch
    from pypde_b200 import Base, _cabi as C
    from oracle import kernels as K, pypde_port as P
    rng = np.random.default_rng(n)
    widths = (70, 33, 1) if n < 2000 else (40, 9)
    DIFF, TDMA_FWD, TDMA_BWD, FDMA_FWD, FDMA_BWD = 0, 1, 2, 3, 4
    run = lambda op, nn, jobs: C.check(C.lib().pde_sweep(op, 1, nn, len(jobs), _sweep_jobs(C, jobs), C.stream()))
    # derivative, scale 1/2 (power of two) -> the reference's diff_2d(c) / 0.5
    cs = [rng.standard_normal((w, n)) for w in widths]
    xs = [T(c, dev) for c in cs]
    ys = [torch.full(c.shape, np.nan, dtype=torch.float64, device=dev) for c in cs]
    run(DIFF, n, [dict(**{"in": [x]}, out=y, nseq=x.shape[0], flag=1, sc=0.5) for x, y in zip(xs, ys)])
    for c, y in zip(cs, ys):
        assert np.array_equal(H(y), K.diff_2d(np.ascontiguousarray(c.T)).T / 0.5)
    # from_chebyshev = S^T + offset-2 tridiagonal solve: n + 2 Chebyshev coefficients -> n Galerkin coefficients
    for kind in ("CD", "CN"):
        b, o = Base(n + 2, kind), P.Basis(n + 2, kind)
        s, a, den, w = b._tables()
        rden = 1.0 / den
        us = [rng.standard_normal((wd, n + 2)) for wd in widths]
        du = [T(u, dev) for u in us]
        dv = [torch.full((u.shape[0], n), np.nan, dtype=torch.float64, device=dev) for u in us]
        run(TDMA_FWD, n, [dict(**{"in": [u, u]}, out=v, nseq=u.shape[0], tab={0: s, 1: a, 2: den, 4: rden})
                          for u, v in zip(du, dv)])
        run(TDMA_BWD, n, [dict(**{"in": [v]}, out=v, nseq=v.shape[0], tab={3: w}) for v in dv])
        # oracle sweep (tdma.f90) on the product's own diagonals: at N = 514 / 2048 the reference's dense BLAS
        # product S^T S (chebyshev.py:339-345) rounds 1 + s_k^2 twice for one or two k, depending on the host's
        # BLAS blocking; the product rounds it once for every k -- this test pins the kernels
        l2, dd, u2 = b._init_stencil_inv()
        sd = b.stencil_diag().copy()
        sd[np.abs(sd) < 1e-12] = 0
        for u, v in zip(us, dv):
            rhs = u[:, :n] + sd[:n] * u[:, 2:]
            assert np.array_equal(H(v), K.solve_tdma_2d(l2, dd, u2, np.ascontiguousarray(rhs.T), 2).T), kind
            if n <= 330:
                assert np.array_equal(H(v), o.from_cheb(np.ascontiguousarray(u.T)).T), kind
    # 4-diagonal solve, forward in place, back substitution into a second array (the stepper's last sweep)
    A = np.zeros((n, n))
    for off in (-2, 0, 2, 4):
        A += np.diag(rng.standard_normal(n - abs(off)) * 0.3 + (3.0 if off == 0 else 0.0), off)
    l, d, u1, u2 = P.fdma_lu(A)
    tl, td, tu1, tu2, trd = (C.upload(t) for t in (l, d, u1, u2, 1.0 / d))
    bs = [rng.standard_normal((wd, n)) for wd in widths]
    dx = [T(bb, dev) for bb in bs]
    dy = [torch.full(bb.shape, np.nan, dtype=torch.float64, device=dev) for bb in bs]
    run(FDMA_FWD, n, [dict(**{"in": [x]}, out=x, nseq=x.shape[0], tab={0: tl}) for x in dx])
    run(FDMA_BWD, n, [dict(**{"in": [x]}, out=y, nseq=x.shape[0], tab={1: td, 2: tu1, 3: tu2, 4: trd})
                      for x, y in zip(dx, dy)])
    torch.cuda.synchronize()
    for bb, y in zip(bs, dy):
        assert np.array_equal(H(y), K.solve_fdma_2d(l, d, u1, u2, bb.copy(), 1))
