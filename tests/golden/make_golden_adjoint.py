"""
Golden fixture for SURVEY.md section 8(f) item 1 (navier/rbc2d_adj.py): tests/golden/adjoint.npz.

Runs ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden_adjoint.py

The UNMODIFIED reference class navier.rbc2d_adj.NavierStokesAdjoint is run under oracle/shim.py (forward model
pre-iterated, state handed to the adjoint iteration, as the reference's own __main__ does), compared bit for bit
with oracle/pypde_port.py::RBC2DAdjoint after every step, and the states are stored.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import pypde_port as P  # noqa: E402
from oracle import shim  # noqa: E402

CASES = {
    "adj32_eu": (dict(shape=(32, 32), dt=0.05, tsave=None, ra=5e3, pr=1.0, dealias=True, integrator="eu", beta=1.0),
                 5, (1, 4)),
    "adj24x32_rk3_aspect2": (dict(shape=(24, 32), dt=0.02, tsave=None, ra=1e4, pr=0.7, dealias=True,
                                  integrator="rk3", beta=1.0, aspect=2.0), 5, (1, 3)),
    "adj32_rk3_nodealias": (dict(shape=(32, 32), dt=0.05, tsave=None, ra=5e3, pr=1.0, dealias=False,
                                 integrator="rk3", beta=1.0), 3, (1, 3)),
}
FIELDS = (("T", "T"), ("U", "U"), ("V", "V"), ("P", "P"), ("pres", "pres"), ("TA", "TA"), ("UA", "UA"), ("VA", "VA"))


def main():
    shim.load_reference()
    adj = importlib.import_module("navier.rbc2d_adj")
    out = {}
    for name, (cfg, pre, snaps) in CASES.items():
        ref = adj.NavierStokesAdjoint(**cfg)
        por = P.RBC2DAdjoint(**cfg)
        ref.NS.set_temperature(amplitude=0.2)
        por.NS.set_temperature(amplitude=0.2)
        for _ in range(pre):
            ref.NS.update()
            por.NS.update()
        ref.U.vhat[:], ref.V.vhat[:], ref.T.vhat[:] = ref.NS.U.vhat, ref.NS.V.vhat, ref.NS.T.vhat
        por.Uhat[:], por.Vhat[:], por.That_[:] = por.NS.Uhat, por.NS.Vhat, por.NS.That_
        out[name + "_init_T"], out[name + "_init_U"], out[name + "_init_V"] = (por.That_.copy(), por.Uhat.copy(),
                                                                               por.Vhat.copy())
        step = 0
        for s in snaps:
            while step < s:
                ref.update()
                por.update()
                step += 1
                st = por.state()
                for key, attr in FIELDS:
                    assert np.array_equal(st[key], np.asarray(getattr(ref, attr).vhat)), (name, step, key)
            for key, _ in FIELDS:
                out["%s_%s_%d" % (name, key, s)] = st[key]
        print(name, "port == reference bit for bit over", step, "adjoint steps")
    np.savez_compressed(os.path.join(HERE, "adjoint.npz"), **out)


if __name__ == "__main__":
    main()
