"""GPU parity of the full IMEX time step: pypde_b200.navier.rbc2d.NavierStokes against the
golden states of the UNMODIFIED reference (tests/golden/rbc_*.npz) and the CPU oracle run
on the same host.  Tolerance of the north star: <= 1e-12 relative (L2) in the spectral
coefficients and in the Nusselt number after 100 steps."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden, rel_l2
from test_oracle_cpu import _cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make(cfg):
    from pypde_b200.navier import rbc2d
    ns = rbc2d.NavierStokes(**cfg)
    ns.set_velocity(m=1, n=1, amplitude=0.2)
    ns.set_temperature(amplitude=0.2)
    k0, k1 = min(16, cfg["shape"][0] - 2), min(16, cfg["shape"][1] - 2)
    pert = 1e-3 * np.random.default_rng(0).standard_normal((k0, k1))
    import torch
    ns.T.vhat[:k0, :k1] += torch.as_tensor(pert, device=ns.T.vhat.device)
    return ns


def H(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name,snaps", [
    ("rbc64_rk3_dealias", (1, 10, 100)),
    ("rbc64_eu_dealias", (1, 10, 100)),
    ("rbc64_eu_nodealias", (1, 10, 100)),
    ("rbc48x64_aspect2", (1, 10)),
    ("zero32x40_beta05", (1, 10)),
    ("linear32x40", (1, 10)),
    ("rbc128_rk3_dealias", (1, 10)),
])
def test_rbc_against_reference_golden(name, snaps):
    g = load_golden("rbc_" + name)
    ns = make(_cases()[name])
    assert rel_l2(H(ns.T.vhat), g["T0"]) < 1e-14 and rel_l2(H(ns.U.vhat), g["U0"]) < 1e-14
    step = 0
    for s in snaps:
        while step < s:
            ns.update()
            ns.update_time()
            step += 1
        for k, t in (("T", ns.T.vhat), ("U", ns.U.vhat), ("V", ns.V.vhat), ("pres", ns.pres.vhat)):
            err = rel_l2(H(t), g["%s_%d" % (k, s)])
            assert err < TOL, "%s step %d field %s: rel L2 %.3e" % (name, s, k, err)
        with contextlib.redirect_stdout(io.StringIO()):
            nu, nuv = ns.eval_Nu()
        assert abs(nu - g["Nu_%d" % s][0]) <= TOL * abs(g["Nu_%d" % s][0]), (name, s, nu, g["Nu_%d" % s][0])
        assert abs(nuv - g["Nu_%d" % s][1]) <= 1e-11 * max(1.0, abs(g["Nu_%d" % s][1]))
    assert abs(ns.time - snaps[-1] * ns.dt) < 1e-12


def test_divergence_stays_small():
    """Domain property (no reference test exists): the projection keeps |div u| small."""
    import torch
    ns = make(_cases()["rbc64_rk3_dealias"])
    for _ in range(20):
        ns.update()
    div = ns.divergence_velocity(ns.U, ns.V)
    assert float(torch.linalg.norm(div)) < 1e-2
    assert torch.isfinite(ns.T.vhat).all()
