"""Per-operator cost of the fused axis passes (pde_pass_run): programs of increasing length on 3 x (n x n) arrays,
both layouts, CUDA-event timed.  Usage: python tools/bench_pass.py [n] > gpurun_out/bench_pass.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import passes as PS  # noqa: E402
from pypde_b200.bases.spectralbase import Base  # noqa: E402
from pypde_b200.templates.hholtz import solverplan_hholtz2d_adi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    njobs = 3
    dev = "cuda"
    rng = np.random.default_rng(0)
    A = [torch.as_tensor(rng.standard_normal((n, n)), device=dev) for _ in range(njobs)]
    B = [torch.zeros((n, n), dtype=torch.float64, device=dev) for _ in range(njobs)]
    C2 = [torch.zeros((n, n), dtype=torch.float64, device=dev) for _ in range(njobs)]
    G = [torch.as_tensor(rng.standard_normal((n, n)), device=dev) for _ in range(4)]
    base = Base(n, "CN")
    sol = solverplan_hholtz2d_adi([Base(n, "CD"), Base(n, "CN")], lam=3e-4, scale=(0.5, 0.5))
    band, plan = sol.plan_for_rhs[0].band, sol.plan_for_lhs[0]
    M = n - 2
    progs = {
        "load_store": lambda p, a, b, c: p.load(a).store(b),
        "load_store_store": lambda p, a, b, c: p.load(a).store(b).store(c),
        "load_scale_store": lambda p, a, b, c: p.load(a).scale(2.0).store(b),
        "load_stencil_store": lambda p, a, b, c: p.load(a).stencil(base).store(b),
        "load_diff_store": lambda p, a, b, c: p.load(a).diff(0.5).store(b),
        "load_band_store": lambda p, a, b, c: p.load(a).band(band).store(b),
        "load_fdma_store": lambda p, a, b, c: p.load(a).fdma(plan).store(b),
        "load_fromcheb_store": lambda p, a, b, c: p.load(a).from_cheb(base).store(b),
        "lincomb3_store": lambda p, a, b, c: p.lincomb([(1.0, a), (0.5, G[0]), (0.25, G[1])]).store(b),
        "lincomb5_store": lambda p, a, b, c: p.lincomb([(1.0, a), (0.5, G[0]), (0.25, G[1]), (2.0, G[2]), (3.0, G[3])]).store(b),
        "load_axpy_store": lambda p, a, b, c: p.load(a).axpy(0.5, G[0]).store(b),
        "px1_like": lambda p, a, b, c: p.load(a).stencil(base).store(b).diff(0.5).store(c),
        "py4_like": lambda p, a, b, c: p.load(a).band(band).fdma(plan).store(b).stencil(base).store(c),
    }
    out = {"n": n, "njobs": njobs, "rows": []}
    only = sys.argv[2] if len(sys.argv) > 2 else None
    for layout in (PS.ROW, PS.COL):
        for name, build in progs.items():
            if only and name != only:
                continue
            L = PS.PassLaunch(layout, n, PS.TableCache())
            for j in range(njobs):
                build(L.job(n), A[j], B[j], C2[j])
            L.finalize()
            for _ in range(3):
                L.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                L.run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            nld = name.count("load") + (3 if "lincomb3" in name else 5 if "lincomb5" in name else 0) + name.count("axpy")
            nst = name.count("store") + (2 if name in ("px1_like", "py4_like") else 0) - (1 if name in ("px1_like", "py4_like") else 0)
            gb = (nld + nst) * njobs * n * n * 8 / 1e9
            out["rows"].append({"layout": "ROW" if layout else "COL", "program": name, "us": round(us, 1),
                                "us_per_seq_strip": round(us / (njobs * n / 8 / 148), 2), "GBps": round(gb / (us * 1e-6), 0)})
            print("%-4s %-22s %9.1f us   %7.0f GB/s" % ("ROW" if layout else "COL", name, us, gb / (us * 1e-6)), file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
