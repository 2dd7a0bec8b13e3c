"""CPU: the self-contained oracle port (oracle/pypde_port.py + fortran_kernels.c) against the
golden fixtures generated from the UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import kernels as K
from oracle import pypde_port as P


def test_primitives_bit_exact(golden_prim):
    g = golden_prim
    for N in (20, 33, 64):
        for kind in ("CH", "CD", "CN"):
            b = P.Basis(N, kind)
            key = "%s%d" % (kind, N)
            assert np.array_equal(b.forward(g[key + "_f"]), g[key + "_forward"])
            assert np.array_equal(b.backward(g[key + "_c"].copy()), g[key + "_backward"])
            for order in (1, 2):
                assert np.array_equal(b.deriv(g[key + "_c"], order), g[key + "_deriv%d" % order])
            if kind != "CH":
                assert np.array_equal(b.to_cheb(g[key + "_c"]), g[key + "_to_cheb"])
                assert np.array_equal(b.from_cheb(g[key + "_u"]), g[key + "_from_cheb"])


def test_banded_solvers_bit_exact(golden_prim):
    g = golden_prim
    l, d, u1, u2 = P.fdma_lu(g["fdma_A"])
    for axis in (0, 1):
        x = K.solve_fdma_2d(l, d, u1, u2, g["fdma_b%d" % axis].copy(), axis)
        assert np.array_equal(x, g["fdma_x%d" % axis])
        # against a dense solve (reference solver/test/test_methods.py:60-102)
        ref = np.linalg.solve(g["fdma_A"], g["fdma_b%d" % axis] if axis == 0 else g["fdma_b1"].T)
        assert np.allclose(x if axis == 0 else x.T, ref, rtol=1e-10, atol=1e-12)
        x = K.solve_twodma_2d(g["twodma_d"], g["twodma_u"], g["fdma_b%d" % axis].copy(), axis)
        assert np.array_equal(x, g["twodma_x%d" % axis])


def test_strided_views_equal_contiguous():
    """The C kernels take explicit strides: F-ordered views must give the same result."""
    rng = np.random.default_rng(3)
    n, m = 17, 9
    c = rng.standard_normal((n, m))
    assert np.array_equal(K.diff_2d(np.asfortranarray(c)), K.diff_2d(c))
    a, b, cc = rng.standard_normal(n - 2), rng.standard_normal(n) + 4, rng.standard_normal(n - 2)
    assert np.array_equal(K.solve_tdma_2d(a, b, cc, np.asfortranarray(c), 2), K.solve_tdma_2d(a, b, cc, c, 2))
    A = np.diag(rng.standard_normal(n) + 4) + np.diag(a, -2) + np.diag(cc, 2)
    x = K.solve_tdma_2d(a, b, cc, c, 2)
    assert np.allclose(x, np.linalg.solve(np.diag(b) + np.diag(a, -2) + np.diag(cc, 2), c))
    del A


def test_edge_cases():
    # shortest recurrences the Fortran indexing allows
    assert np.array_equal(K.diff_1d(np.array([1.0, 2.0, 3.0])), np.array([2.0, 12.0, 0.0]))
    c = np.zeros(6)
    c[5] = 1.0
    # d/dx T_5 = 10 T_4 + 10 T_2 + 5 T_0
    assert np.array_equal(K.diff_1d(c), np.array([5.0, 0, 10.0, 0, 10.0, 0]))
    # empty batch
    assert K.diff_2d(np.zeros((5, 0))).shape == (5, 0)


def test_fields_and_solvers(golden_fs):
    g = golden_fs
    N0, N1 = 40, 20
    for kx, ky in (("CD", "CN"), ("CH", "CH"), ("CN", "CD")):
        s = P.Space([P.Basis(N0, kx, 3 / 2), P.Basis(N1, ky, 3 / 2)])
        key = "f2d_%s%s" % (kx, ky)
        assert np.array_equal(s.forward(g[key + "_v"]), g[key + "_fwd"])
        assert np.array_equal(s.backward(g[key + "_vhat"]), g[key + "_bwd"])
        assert np.array_equal(s.dealias.backward(g[key + "_vhat"]), g[key + "_dbwd"])
        assert np.array_equal(s.dealias.forward(g[key + "_dbwd"] * g[key + "_dbwd"]), g[key + "_dfwd"])
        for deriv in ((1, 0), (0, 1), (2, 0), (1, 1)):
            assert np.array_equal(s.grad(g[key + "_vhat"], deriv, (0.75, 0.5)), g[key + "_grad%d%d" % deriv])
    N0, N1 = 50, 40
    for kx, ky in (("CD", "CN"), ("CN", "CN"), ("CD", "CD")):
        key = kx + ky
        bases = [P.Basis(N0, kx), P.Basis(N1, ky)]
        h = P.HelmholtzADI(bases, 0.013, (0.75, 0.5))
        r = h.solve_rhs(g["hh_" + key + "_rhs"])
        r += h.solve_old(g["hh_" + key + "_old"])
        assert np.array_equal(h.solve_lhs(r), g["hh_" + key + "_x"])
        # Poisson: depends on LAPACK inv/eig of this host -> tolerance, not bits
        p = P.PoissonEig(bases, singular=(key == "CNCN"), scale=(0.75, 0.5))
        x = p.solve_lhs(p.solve_rhs(g["hh_" + key + "_rhs"]))
        assert rel_l2(x, g["po_" + key + "_x"]) < 1e-9


def test_dealias_is_alias_free(golden_fs):
    """3/2-rule: the truncated coefficients of a product are independent of the quadrature
    grid once it has >= 3N/2 points (basis of the D' = 3N/2+1 grid used on the fused path)."""
    rng = np.random.default_rng(5)
    N = 32
    a = rng.standard_normal((N, 4))
    b = rng.standard_normal((N, 4))
    out = []
    for D in (48, 49, 64):
        bD = P.Basis(D, "CH")
        pa = bD.backward(np.pad(a, ((0, D - N), (0, 0))))
        pb = bD.backward(np.pad(b, ((0, D - N), (0, 0))))
        out.append(bD.forward(pa * pb)[:N])
    assert rel_l2(out[1], out[0]) < 1e-14 and rel_l2(out[2], out[0]) < 1e-14


@pytest.mark.parametrize("name,steps", [("rbc64_rk3_dealias", 10), ("rbc64_eu_nodealias", 10),
                                        ("zero32x40_beta05", 10), ("linear32x40", 10), ("rbc48x64_aspect2", 10)])
def test_rbc_port_matches_reference(name, steps):
    import sys
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "golden"))
    g = load_golden("rbc_" + name)
    cfg = _cases()[name]
    o = P.RBC2D(**cfg)
    o.set_velocity(m=1, n=1, amplitude=0.2)
    o.set_temperature(amplitude=0.2)
    k0, k1 = min(16, cfg["shape"][0] - 2), min(16, cfg["shape"][1] - 2)
    o.That_[:k0, :k1] += 1e-3 * np.random.default_rng(0).standard_normal((k0, k1))
    assert np.array_equal(o.That_, g["T0"]) and np.array_equal(o.Uhat, g["U0"])
    o.iterate(steps)
    for k, a in (("T", o.That_), ("U", o.Uhat), ("V", o.Vhat), ("pres", o.pres)):
        assert rel_l2(a, g["%s_%d" % (k, steps)]) < 1e-10, k
    nu = o.eval_Nu()
    assert abs(nu[0] - g["Nu_%d" % steps][0]) < 1e-10 * abs(g["Nu_%d" % steps][0])


def _cases():
    return {
        "rbc64_rk3_dealias": dict(case="rbc", shape=(64, 64), ra=1e5, pr=1.0, dt=0.01, tsave=None, dealias=True,
                                  integrator="rk3", beta=1.0, aspect=1.0),
        "rbc64_eu_dealias": dict(case="rbc", shape=(64, 64), ra=1e5, pr=1.0, dt=0.01, tsave=None, dealias=True,
                                 integrator="eu", beta=1.0, aspect=1.0),
        "rbc64_eu_nodealias": dict(case="rbc", shape=(64, 64), ra=1e4, pr=1.0, dt=0.02, tsave=None, dealias=False,
                                   integrator="eu", beta=1.0, aspect=1.0),
        "rbc48x64_aspect2": dict(case="rbc", shape=(48, 64), ra=1e4, pr=0.7, dt=0.02, tsave=None, dealias=True,
                                 integrator="rk3", beta=1.0, aspect=2.0),
        "zero32x40_beta05": dict(case="zero", shape=(32, 40), ra=1e4, pr=1.0, dt=0.01, tsave=None, dealias=True,
                                 integrator="rk3", beta=0.5, aspect=1.0),
        "linear32x40": dict(case="linear", shape=(32, 40), ra=1e4, pr=1.0, dt=0.01, tsave=None, dealias=True,
                            integrator="eu", beta=1.0, aspect=1.0),
        "rbc128_rk3_dealias": dict(case="rbc", shape=(128, 128), ra=1e6, pr=1.0, dt=0.005, tsave=None, dealias=True,
                                   integrator="rk3", beta=1.0, aspect=1.0),
    }


@pytest.mark.parametrize("name,cfg,snaps", [
    ("d48x40", dict(shape=(48, 40), dt=0.01, kappa=0.1, beta=0.5), (1, 10, 100)),
    ("d33x64_beta1", dict(shape=(33, 64), dt=0.02, kappa=0.05, beta=1.0), (1, 20))])
def test_diffusion_port_matches_reference(name, cfg, snaps):
    """oracle Diffusion2D against the states of the unmodified reference script
    diffusion/diff_2d-bc.py (tests/golden/make_golden_diffusion.py)."""
    g = load_golden("diffusion")
    o = P.Diffusion2D(**cfg)
    assert rel_l2(o.fhat, g[name + "_fhat"]) < 1e-13
    step = 0
    for s in snaps:
        while step < s:
            o.update()
            step += 1
        assert rel_l2(o.vhat, g["%s_vhat_%d" % (name, s)]) < 1e-12, s
    assert rel_l2(o.total(), g[name + "_total"]) < 1e-12


# name -> (config, forward steps before the adjoint iteration starts, snapshots): tests/golden/make_golden_adjoint.py
ADJOINT_CASES = {
    "adj32_eu": (dict(shape=(32, 32), dt=0.05, tsave=None, ra=5e3, pr=1.0, dealias=True, integrator="eu", beta=1.0),
                 5, (1, 4)),
    "adj24x32_rk3_aspect2": (dict(shape=(24, 32), dt=0.02, tsave=None, ra=1e4, pr=0.7, dealias=True,
                                  integrator="rk3", beta=1.0, aspect=2.0), 5, (1, 3)),
    "adj32_rk3_nodealias": (dict(shape=(32, 32), dt=0.05, tsave=None, ra=5e3, pr=1.0, dealias=False,
                                 integrator="rk3", beta=1.0), 3, (1, 3)),
}


@pytest.mark.parametrize("name", sorted(ADJOINT_CASES))
def test_adjoint_port_matches_reference(name):
    """oracle RBC2DAdjoint against the states of the unmodified reference navier/rbc2d_adj.py
    (tests/golden/make_golden_adjoint.py; bit for bit there, <= 1e-12 here: the Poisson setup is host dependent)."""
    g = load_golden("adjoint")
    cfg, pre, snaps = ADJOINT_CASES[name]
    o = P.RBC2DAdjoint(**cfg)
    o.NS.set_temperature(amplitude=0.2)          # like the reference's __main__: pre-iterate the forward model,
    for _ in range(pre):                         # hand its state (and its pressure history) to the adjoint iteration
        o.NS.update()
    o.That_[:], o.Uhat[:], o.Vhat[:] = o.NS.That_, o.NS.Uhat, o.NS.Vhat
    for key, arr in (("T", o.That_), ("U", o.Uhat), ("V", o.Vhat)):
        assert rel_l2(arr, g["%s_init_%s" % (name, key)]) < 1e-12, key
    step = 0
    for s in snaps:
        while step < s:
            o.update()
            step += 1
        st = o.state()
        for key in ("T", "U", "V", "pres", "TA", "UA", "VA"):
            assert rel_l2(st[key], g["%s_%s_%d" % (name, key, s)]) < 1e-12, (key, s)
