"""GPU parity of the full IMEX time step: pypde_b200.navier.rbc2d.NavierStokes against
  (a) the CPU oracle (oracle/pypde_port.py, bit-identical to the unmodified reference, see
      tests/golden/make_golden.py) run on THIS host from the same seeded state, and
  (b) the golden states the unmodified reference produced in the build container.

Tolerance of the north star: <= 1e-12 relative (L2 norm-wise) in the spectral coefficients.

Two facts measured on the B200 box shape the assertions (tools/diag_rbc.py):
  * the reference's Poisson setup (numpy inv + eig of a matrix with cond ~1e12,
    templates/poisson.py:93-98) is not reproducible across host CPUs beyond 64x64: the SAME
    oracle gives U, V differing by 4e-12 .. 2e-11 from the golden at 128x128 after one step on
    another CPU.  (b) therefore allows the host's own oracle-vs-golden distance on top of 1e-12;
  * the Nusselt diagnostic (rbc2d_base.py:344-363) differentiates at the wall after a
    physical-space round trip, which amplifies 1-ulp transform differences by ~N^2: it is
    checked (i) through the oracle's own diagnostic applied to the CUDA state (<= max(1e-12, 4 N^2 eps)) and
    (ii) through the CUDA diagnostic with the conditioning bound 16 N^2 eps.
"""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden, rel_l2
from test_oracle_cpu import _cases

pytestmark = pytest.mark.gpu
TOL = 1e-12
EPS = np.finfo(float).eps


def _pert(cfg):
    k0, k1 = min(16, cfg["shape"][0] - 2), min(16, cfg["shape"][1] - 2)
    return k0, k1, 1e-3 * np.random.default_rng(0).standard_normal((k0, k1))


def make(cfg, **kw):
    import torch
    from pypde_b200.navier import rbc2d
    ns = rbc2d.NavierStokes(**cfg, **kw)
    ns.set_velocity(m=1, n=1, amplitude=0.2)
    ns.set_temperature(amplitude=0.2)
    k0, k1, pert = _pert(cfg)
    ns.T.vhat[:k0, :k1] += torch.as_tensor(pert, device=ns.T.vhat.device)
    return ns


def make_oracle(cfg):
    from oracle import pypde_port as P
    o = P.RBC2D(**cfg)
    o.set_velocity(m=1, n=1, amplitude=0.2)
    o.set_temperature(amplitude=0.2)
    k0, k1, pert = _pert(cfg)
    o.That_[:k0, :k1] += pert
    return o


def H(t):
    return t.detach().cpu().numpy()


def oracle_nu_of(o, ns):
    """The oracle's Nusselt diagnostic evaluated on the CUDA path's state."""
    saved = (o.That_, o.Vhat)
    o.That_, o.Vhat = H(ns.T.vhat).copy(), H(ns.V.vhat).copy()
    nu = o.eval_Nu()
    o.That_, o.Vhat = saved
    return nu


@pytest.mark.parametrize("name,snaps", [
    ("rbc64_rk3_dealias", (1, 10, 100)),
    ("rbc64_eu_dealias", (1, 10, 100)),
    ("rbc64_eu_nodealias", (1, 10, 100)),
    ("rbc48x64_aspect2", (1, 10)),
    ("zero32x40_beta05", (1, 10)),
    ("linear32x40", (1, 10)),
    ("rbc128_rk3_dealias", (1, 10)),
])
def test_rbc_step_parity(name, snaps):
    cfg = _cases()[name]
    g = load_golden("rbc_" + name)
    ns, o = make(cfg), make_oracle(cfg)
    assert rel_l2(H(ns.T.vhat), g["T0"]) < 1e-14 and rel_l2(H(ns.U.vhat), g["U0"]) < 1e-14
    N = max(cfg["shape"])
    step = 0
    for s in snaps:
        while step < s:
            ns.update()
            ns.update_time()
            o.update()
            step += 1
        for k, t, r in (("T", ns.T.vhat, o.That_), ("U", ns.U.vhat, o.Uhat), ("V", ns.V.vhat, o.Vhat),
                        ("pres", ns.pres.vhat, o.pres)):
            err = rel_l2(H(t), r)
            assert err < TOL, "%s step %d %s: CUDA vs oracle rel L2 %.3e" % (name, s, k, err)
            gk = g["%s_%d" % (k, s)]
            host = rel_l2(r, gk)          # 0 where this host's LAPACK reproduces the build container's
            errg = rel_l2(H(t), gk)
            assert errg < TOL + 2 * host, "%s step %d %s: CUDA vs golden %.3e (host %.3e)" % (name, s, k, errg, host)
        nu_o = o.eval_Nu()
        nu_x = oracle_nu_of(o, ns)
        tol_nu = max(TOL, 4 * N * N * EPS)     # conditioning of the wall derivative, see module docstring
        assert abs(nu_x[0] - nu_o[0]) <= tol_nu * abs(nu_o[0]), ("Nu (oracle diagnostic on CUDA state)", nu_x, nu_o)
        assert abs(nu_x[1] - nu_o[1]) <= tol_nu * max(1.0, abs(nu_o[1]))
        with contextlib.redirect_stdout(io.StringIO()):
            nu, nuv = ns.eval_Nu()
        bound = 16 * N * N * EPS
        assert abs(nu - nu_o[0]) <= bound * abs(nu_o[0]), ("Nu (CUDA diagnostic)", nu, nu_o[0], bound)
        assert abs(nuv - nu_o[1]) <= bound * max(1.0, abs(nu_o[1]))
    assert abs(ns.time - snaps[-1] * ns.dt) < 1e-12


def test_reference_dealias_grid_also_matches():
    """dealias_grid="reference" (exactly int(3N/2) points, dense DCT) and the default FFT-friendly
    grid agree with the oracle and with each other."""
    cfg = _cases()["rbc64_rk3_dealias"]
    a, b, o = make(cfg, dealias_grid="reference"), make(cfg), make_oracle(cfg)
    assert tuple(a.U.dealias.shape_physical) == (96, 96) and tuple(b.U.dealias.shape_physical) == (97, 97)
    for _ in range(5):
        a.update()
        b.update()
        o.update()
    for x, y, r in ((a.T.vhat, b.T.vhat, o.That_), (a.U.vhat, b.U.vhat, o.Uhat), (a.pres.vhat, b.pres.vhat, o.pres)):
        assert rel_l2(H(x), r) < TOL and rel_l2(H(y), r) < TOL


@pytest.mark.parametrize("name", ["rbc64_rk3_dealias", "rbc64_eu_nodealias", "rbc48x64_aspect2", "zero32x40_beta05"])
def test_reference_ordered_stepper(name):
    """stepper="reference" (operator by operator, the reference's order) against the oracle."""
    cfg = _cases()[name]
    ns, o = make(cfg, stepper="reference"), make_oracle(cfg)
    for _ in range(5):
        ns.update()
        o.update()
    for t, r in ((ns.T.vhat, o.That_), (ns.U.vhat, o.Uhat), (ns.V.vhat, o.Vhat), (ns.pres.vhat, o.pres)):
        assert rel_l2(H(t), r) < TOL


def test_cuda_graph_step_matches():
    """graph=True replays the captured RK3 step: identical bits to the eager batched stepper."""
    import torch
    cfg = _cases()["rbc64_rk3_dealias"]
    a, b = make(cfg), make(cfg, graph=True)
    for _ in range(6):
        a.update()
        b.update()
    torch.cuda.synchronize()
    for x, y in ((a.T.vhat, b.T.vhat), (a.U.vhat, b.U.vhat), (a.V.vhat, b.V.vhat), (a.pres.vhat, b.pres.vhat)):
        assert torch.equal(x, y)
    # re-binding after the user replaced a field tensor
    b.T.vhat = b.T.vhat.clone()
    a.update()
    b.update()
    assert torch.equal(a.T.vhat, b.T.vhat)


def test_nusselt_diagnostic_conditioning():
    """Evidence for the Nu tolerance: the ORACLE's own Nu moves by > 1e-13 relative when its
    input state is perturbed by 1e-15 relative (so 1e-12 is at the diagnostic's noise floor)."""
    cfg = _cases()["rbc64_rk3_dealias"]
    o = make_oracle(cfg)
    o.iterate(3)
    nu0 = o.eval_Nu()[0]
    rng = np.random.default_rng(1)
    o.That_ = o.That_ * (1.0 + 1e-15 * rng.standard_normal(o.That_.shape))
    nu1 = o.eval_Nu()[0]
    assert abs(nu1 - nu0) / abs(nu0) < 64 * 64 * 16 * EPS


def test_divergence_stays_small():
    """Domain property (the reference has no test of the stepper): projection keeps |div u| small."""
    import torch
    ns = make(_cases()["rbc64_rk3_dealias"])
    for _ in range(20):
        ns.update()
    div = ns.divergence_velocity(ns.U, ns.V)
    assert float(torch.linalg.norm(div)) < 1e-2
    assert torch.isfinite(ns.T.vhat).all()


def test_slab_stepper_two_gpus():
    """Slab decomposition over 2 GPUs (NCCL all-to-all transposes) against the oracle; skipped on
    single-GPU boxes (the CPU side of the plumbing is covered by tests/test_slab_cpu.py)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    script = os.path.join(os.path.dirname(__file__), "dist_slab_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                         capture_output=True, text=True, timeout=600)
    assert "SLAB OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("batched", [True, False])
def test_ensemble_members_match_standalone_runs(batched):
    """Ensemble == standalone runs.  batched: one launch list / one CUDA graph for all members (every pass takes a
    job per member, dense DCT / projections / products through the batched entry points); not batched: one CUDA
    graph per member on several streams."""
    import torch
    from pypde_b200.navier.ensemble import Ensemble
    from pypde_b200.navier import rbc2d
    kw = dict(case="rbc", shape=(32, 32), pr=1.0, dt=0.01, tsave=None, dealias=True, integrator="rk3",
              beta=1.0, aspect=1.0)
    ras = [1e4, 3e4, 1e5, 3e5, 1e6]
    ens = Ensemble(ras, streams=3, batched=batched, **kw)
    assert ens.batched == batched
    solo = [rbc2d.NavierStokes(ra=r, **kw) for r in ras]
    for m in ens.members + solo:
        m.set_velocity(m=1, n=1, amplitude=0.2)
        m.set_temperature(amplitude=0.2)
    for _ in range(5):
        ens.update()
        for s in solo:
            s.update()
    torch.cuda.synchronize()
    for m, s in zip(ens.members, solo):
        for a, b in ((m.T.vhat, s.T.vhat), (m.U.vhat, s.U.vhat), (m.V.vhat, s.V.vhat), (m.pres.vhat, s.pres.vhat)):
            assert rel_l2(H(a), H(b)) < 1e-13, rel_l2(H(a), H(b))
        if not batched:
            assert torch.equal(m.T.vhat, s.T.vhat) and torch.equal(m.U.vhat, s.U.vhat)
    # sharding: rank r of 2 gets every second member
    assert Ensemble(ras, rank=1, world=2, **kw).indices == [1, 3]


def test_batched_ensemble_vs_oracle_128():
    """configs[4] grid: 6 members of the 128 x 128 Ra sweep advanced together, two of them against the oracle."""
    import torch
    from pypde_b200.navier.ensemble import Ensemble
    kw = dict(case="rbc", shape=(128, 128), pr=1.0, dt=0.005, tsave=None, dealias=True, integrator="rk3",
              beta=1.0, aspect=1.0)
    ras = list(np.logspace(4, 8, 6))
    ens = Ensemble(ras, **kw)
    assert ens.batched
    for m in ens.members:
        m.set_velocity(m=1, n=1, amplitude=0.2)
        m.set_temperature(amplitude=0.2)
    for _ in range(3):
        ens.update()
    torch.cuda.synchronize()
    from oracle import pypde_port as P
    for k in (0, 4):
        o = P.RBC2D(ra=float(ras[k]), **kw)
        o.set_velocity(m=1, n=1, amplitude=0.2)
        o.set_temperature(amplitude=0.2)
        o.iterate(3)
        m = ens.members[k]
        for t, r in ((m.T.vhat, o.That_), (m.U.vhat, o.Uhat), (m.V.vhat, o.Vhat), (m.pres.vhat, o.pres)):
            assert rel_l2(H(t), r) < TOL, (k, rel_l2(H(t), r))


def _preiterated_pair(cfg, pre):
    from oracle import pypde_port as P
    from pypde_b200.navier import rbc2d
    ns, o = rbc2d.NavierStokes(**cfg), P.RBC2D(**cfg)
    ns.set_temperature(amplitude=0.2)
    o.set_temperature(amplitude=0.2)
    for _ in range(pre):
        ns.update()
        o.update()
    return ns, o


def test_steady_state_residual_matches_oracle():
    """NavierStokesSteadyState.steady_fun (rbc2d_base.py:318-341): residual (NS(X) - X) / dt of the flat host vector
    [T, U, V], evaluated with the device step, against the oracle's forward step.  The residual is a difference of
    two states that agree to ~1e-14, divided by dt: its relative accuracy is 1e-14 |X| / |NS(X) - X|."""
    cfg = dict(case="rbc", shape=(32, 32), ra=5e3, pr=1.0, dt=0.05, tsave=None, dealias=True, integrator="eu",
               beta=1.0, aspect=1.0)
    ns, o = _preiterated_pair(cfg, 5)
    X = ns.vectorify()
    parts = ns.reshape(X)
    assert [p.shape for p in parts] == [tuple(f.vhat.shape) for f in (ns.T, ns.U, ns.V)]
    assert np.array_equal(np.concatenate([p.ravel() for p in parts]), X)
    Xo = np.concatenate([o.That_.ravel(), o.Uhat.ravel(), o.Vhat.ravel()])
    assert rel_l2(X, Xo) < 1e-12
    F = ns.steady_fun(X, ns, None)
    o.That_[:], o.Uhat[:], o.Vhat[:] = parts
    o.update()
    Fo = (np.concatenate([o.That_.ravel(), o.Uhat.ravel(), o.Vhat.ravel()]) - X) / cfg["dt"]
    amplification = np.linalg.norm(X) / (np.linalg.norm(Fo) * cfg["dt"])
    assert rel_l2(F, Fo) < max(1e-12, 1e-13 * amplification), (rel_l2(F, Fo), amplification)


def test_solve_steady_state_mechanics():
    """solve_steady_state (rbc2d_base.py:266-296): SciPy's Newton-Krylov on the host drives the device step.
    Checked for its mechanics (flat host vector in, result with a finite vector of the same size out, the model's
    fields hold the last iterate): like the reference's, the residual function carries the pressure history of
    the model from one evaluation to the next, so convergence is not a property of X alone; the residual itself
    is pinned against the oracle in test_steady_state_residual_matches_oracle."""
    from pypde_b200.navier import rbc2d
    cfg = dict(case="rbc", shape=(24, 24), ra=1e3, pr=1.0, dt=0.1, tsave=None, dealias=True, integrator="eu",
               beta=1.0, aspect=1.0)
    ns = rbc2d.NavierStokes(**cfg)
    X0 = ns.vectorify()
    assert X0.shape == (22 * 22 * 3,) and not X0.any()
    ns.set_temperature(amplitude=0.05)
    X1 = ns.vectorify()
    assert np.linalg.norm(ns.steady_fun(X1, ns, None)) > 1e-3
    sol = ns.solve_steady_state(X0=X1, maxiter=2, disp=False, tol=1e-10)
    assert np.asarray(sol.x).shape == X1.shape and np.isfinite(np.asarray(sol.x)).all()
    assert np.isfinite(ns.vectorify()).all()
