"""Where do the CUDA diffusion step and the oracle differ? (run on the GPU box)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion"))
from diff_2d_bc import Diffusion2d  # noqa: E402
from oracle import pypde_port as P  # noqa: E402
from pypde_b200 import grad  # noqa: E402


def rel(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else a
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


for shape in ((48, 40), (33, 64), (1024, 1024)):
    cfg = dict(shape=shape, dt=0.01, kappa=0.1, beta=0.5)
    D = Diffusion2d(tsave=None, **cfg)
    o = P.Diffusion2D(**cfg)
    print(shape, "bc_v", rel(D.fieldbc.v, o.bc_v), "bc_vhat", rel(D.fieldbc.vhat, o.bc_vhat), "fhat", rel(D._fhat, o.fhat))
    # same fhat on both sides: isolates the step
    D2 = Diffusion2d(tsave=None, **cfg)
    D2._fhat_cache = torch.as_tensor(o.fhat, device="cuda")
    o2 = P.Diffusion2D(**cfg)
    for step in range(1, 4):
        D.update(); o.update(); D2.update(); o2.update()
        g1 = grad(D2.field, deriv=(0, 2)); g2 = grad(D2.field, deriv=(2, 0))
        print("  step", step, "vhat", rel(D.field.vhat, o.vhat), " with oracle fhat:", rel(D2.field.vhat, o2.vhat),
              "grads", rel(g1, o2.space.grad(o2.vhat, (0, 2))), rel(g2, o2.space.grad(o2.vhat, (2, 0))))
    r = torch.as_tensor(o.fhat, device="cuda")
    a = D.solver.solve_rhs(r.clone())
    print("  solve_rhs(fhat) same input:", rel(a, o.solver.solve_rhs(o.fhat.copy())),
          " own fhat:", rel(D.solver.solve_rhs(D._fhat.clone()), o.solver.solve_rhs(o.fhat.copy())))
