"""Diagnostic (run under gpurun): separates host-LAPACK sensitivity from CUDA-path error.
For each golden case prints rel-L2 of: oracle-on-this-host vs golden(reference, build container),
product vs oracle-on-this-host, product vs golden."""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
from conftest import load_golden, rel_l2  # noqa: E402
from test_oracle_cpu import _cases  # noqa: E402
from test_gpu_rbc import make, H  # noqa: E402
from oracle import pypde_port as P  # noqa: E402

names = sys.argv[1:] or ["rbc64_rk3_dealias", "rbc64_eu_nodealias", "rbc128_rk3_dealias", "zero32x40_beta05"]
for name in names:
    cfg = _cases()[name]
    g = load_golden("rbc_" + name)
    snaps = sorted(int(k.split("_")[1]) for k in g if k.startswith("Nu_"))
    ns = make(cfg)
    o = P.RBC2D(**cfg)
    o.set_velocity(m=1, n=1, amplitude=0.2)
    o.set_temperature(amplitude=0.2)
    k0, k1 = min(16, cfg["shape"][0] - 2), min(16, cfg["shape"][1] - 2)
    o.That_[:k0, :k1] += 1e-3 * np.random.default_rng(0).standard_normal((k0, k1))
    step = 0
    for s in snaps:
        while step < s:
            ns.update(); ns.update_time(); o.update(); step += 1
        for k, t, r in (("T", ns.T.vhat, o.That_), ("U", ns.U.vhat, o.Uhat), ("V", ns.V.vhat, o.Vhat),
                        ("pres", ns.pres.vhat, o.pres), ("P", ns.P.vhat, o.Phat)):
            gk = g["%s_%d" % (k, s)]
            print("%-20s step %3d %-4s  oracle_here-vs-golden %.2e   cuda-vs-oracle_here %.2e   cuda-vs-golden %.2e"
                  % (name, s, k, rel_l2(r, gk), rel_l2(H(t), r), rel_l2(H(t), gk)))
        with contextlib.redirect_stdout(io.StringIO()):
            nu = ns.eval_Nu()
        nuo = o.eval_Nu()
        gn = g["Nu_%d" % s]
        print("%-20s step %3d Nu    oracle_here-vs-golden %.2e   cuda-vs-oracle_here %.2e   cuda-vs-golden %.2e"
              % (name, s, abs(nuo[0] - gn[0]) / abs(gn[0]), abs(nu[0] - nuo[0]) / abs(nuo[0]), abs(nu[0] - gn[0]) / abs(gn[0])))
