"""GPU parity of the adjoint-descent iteration (pypde_b200/navier/rbc2d_adj.py)
against the CPU oracle oracle/pypde_port.py::RBC2DAdjoint, which tests/golden/make_golden_adjoint.py pins bit for
bit to the unmodified reference navier/rbc2d_adj.py.

Protocol of the reference's own __main__: pre-iterate the forward model, hand its state to the adjoint
iteration, step.  Tolerances: the iteration divides the difference of two forward states by dt (:193-197), so
the 1e-14-level differences of the forward step (DCT summation order) are amplified by 1/(dt |residual|/|state|)
in the adjoint fields; the state itself moves by dt times those fields per step.
Measured on B200 (tools/diag_adjoint.py): T <= 3e-15, U, V <= 1.4e-13, pres <= 3.8e-13, TA <= 5.4e-13,
UA, VA <= 7e-15 over all cases and steps -- inside the 1e-12 of the north star."""
import numpy as np
import pytest

from conftest import rel_l2
from test_oracle_cpu import ADJOINT_CASES

pytestmark = pytest.mark.gpu

TOL_STATE, TOL_ADJOINT = 1e-12, 2e-12


def run_pair(name):
    import torch
    from oracle import pypde_port as P
    from pypde_b200.navier.rbc2d_adj import NavierStokesAdjoint
    cfg, pre, snaps = ADJOINT_CASES[name]
    A, o = NavierStokesAdjoint(**cfg), P.RBC2DAdjoint(**cfg)
    A.NS.set_temperature(amplitude=0.2)
    o.NS.set_temperature(amplitude=0.2)
    for _ in range(pre):
        A.NS.update()
        o.NS.update()
    A.U.vhat[:], A.V.vhat[:], A.T.vhat[:] = A.NS.U.vhat, A.NS.V.vhat, A.NS.T.vhat
    o.Uhat[:], o.Vhat[:], o.That_[:] = o.NS.Uhat, o.NS.Vhat, o.NS.That_
    errs = []
    for step in range(1, max(snaps) + 1):
        A.update()
        o.update()
        torch.cuda.synchronize()
        ref = o.state()
        dev = {"T": A.T.vhat, "U": A.U.vhat, "V": A.V.vhat, "pres": A.pres.vhat, "TA": A.TA.vhat, "UA": A.UA.vhat,
               "VA": A.VA.vhat}
        errs.append({k: rel_l2(v.cpu().numpy(), ref[k]) for k, v in dev.items()})
    return errs


@pytest.mark.parametrize("name", sorted(ADJOINT_CASES))
def test_adjoint_iteration_matches_oracle(name):
    for step, e in enumerate(run_pair(name), 1):
        for key in ("T", "U", "V", "pres"):
            assert e[key] < TOL_STATE, (step, key, e)
        for key in ("TA", "UA", "VA"):
            assert e[key] < TOL_ADJOINT, (step, key, e)


def test_adjoint_api_surface():
    """Same names as the reference class (rbc2d_adj.py:15-356)."""
    from pypde_b200.navier import NavierStokesAdjoint
    for name in ("update", "update_NS", "update_U", "update_V", "update_T", "update_P", "update_pres",
                 "update_velocity", "conv", "conv_term_adj_ux", "conv_term_adj_uz", "setup_solver", "reset_time",
                 "set_temperature", "iterate", "callback", "save"):
        assert callable(getattr(NavierStokesAdjoint, name)), name
    with pytest.raises(ValueError):
        NavierStokesAdjoint(case="nope")
