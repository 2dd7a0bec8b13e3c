"""debug: from_cheb (TdmaFwd + TdmaBwd through pde_sweep, axis 1) at n = 2046 against the oracle; prints mismatches"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import Base, _cabi as C
from oracle import pypde_port as P
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from test_gpu_primitives import _sweep_jobs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2046
dev = torch.device("cuda")
for kind in ("CD", "CN"):
    rng = np.random.default_rng(n)
    b, o = Base(n + 2, kind), P.Basis(n + 2, kind)
    s, a, den, w = b._tables()
    rden = 1.0 / den
    for wd in (40, 9):
        u = rng.standard_normal((wd, n + 2))
        du = torch.as_tensor(u, device=dev)
        dv = torch.full((wd, n), np.nan, dtype=torch.float64, device=dev)
        jobs = [dict(**{"in": [du, du]}, out=dv, nseq=wd, tab={0: s, 1: a, 2: den, 4: rden})]
        C.check(C.lib().pde_sweep(1, 1, n, 1, _sweep_jobs(C, jobs), C.stream()))
        g = dv.cpu().numpy().copy()
        jobs = [dict(**{"in": [dv]}, out=dv, nseq=wd, tab={3: w})]
        C.check(C.lib().pde_sweep(2, 1, n, 1, _sweep_jobs(C, jobs), C.stream()))
        v = dv.cpu().numpy()
        ref = o.from_cheb(np.ascontiguousarray(u.T)).T
        bad = np.argwhere(v != ref)
        print(kind, wd, "mismatches", len(bad), "first", bad[:6].tolist(), "rows", sorted(set(bad[:, 0].tolist()))[:10],
              "cols", (int(bad[:, 1].min()), int(bad[:, 1].max())) if len(bad) else None)
        if len(bad):
            r, c = bad[0]
            print("   v", v[r, c], "ref", ref[r, c], "rel", abs(v[r, c] - ref[r, c]) / abs(ref[r, c]))
