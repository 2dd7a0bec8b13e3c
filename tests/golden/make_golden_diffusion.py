"""
Golden fixture for BASELINE.json configs[2] (diffusion/diff_2d-bc.py): tests/golden/diffusion.npz.

Runs ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden_diffusion.py

The reference's example is a script; its `Diffusion2d` class definition (everything above the
first module-level statement that instantiates it) is executed UNMODIFIED under oracle/shim.py,
stepped, and compared bit-for-bit with oracle/pypde_port.py::Diffusion2D.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import pypde_port as P  # noqa: E402
from oracle import shim  # noqa: E402

shim.load_reference()

SCRIPT = "/root/reference/diffusion/diff_2d-bc.py"
CASES = {
    "d48x40": (dict(shape=(48, 40), dt=0.01, kappa=0.1, beta=0.5), (1, 10, 100)),
    "d33x64_beta1": (dict(shape=(33, 64), dt=0.02, kappa=0.05, beta=1.0), (1, 20)),
}


def reference_class():
    src = open(SCRIPT).read()
    cut = src.index("\nD = Diffusion2d(")
    ns = {"__name__": "diff_2d_bc_reference"}
    exec(compile(src[:cut], SCRIPT, "exec"), ns)
    return ns["Diffusion2d"]


def main():
    Ref = reference_class()
    out = {}
    for name, (cfg, snaps) in CASES.items():
        ref = Ref(tsave=None, **cfg)
        por = P.Diffusion2D(**cfg)
        assert np.array_equal(np.asarray(ref._fhat), por.fhat), name
        out[name + "_fhat"] = por.fhat.copy()
        out[name + "_bc_v"] = por.bc_v.copy()
        step = 0
        for s in snaps:
            while step < s:
                ref.update()
                ref.update_time()
                por.update()
                step += 1
            assert np.array_equal(ref.field.vhat, por.vhat), (name, s)
            out["%s_vhat_%d" % (name, s)] = por.vhat.copy()
            print("%-14s step %4d  |vhat|=%.15e  bit-equal reference/port" % (name, s, np.linalg.norm(por.vhat)))
        ref.field.backward()
        assert np.array_equal(ref.field.v + ref.fieldbc.v, por.total()), name
        out[name + "_total"] = por.total()
    np.savez_compressed(os.path.join(HERE, "diffusion.npz"), **out)
    print("written", os.path.join(HERE, "diffusion.npz"))


if __name__ == "__main__":
    main()
