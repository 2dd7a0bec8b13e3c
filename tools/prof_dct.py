"""Profiling driver (run under ncu): a few launches of the FFT DCT-I and the chain kernels at
rbc2048 sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import ops, Base  # noqa: E402

L, nb = 3073, 2048
plan = ops.DctPlan.get(L)
# row pitch padded by 64 bytes like the stepper's work arrays (fast_stepper._new): a 16 KB pitch maps a whole
# column strip onto a few HBM channels and makes the axis-0 timing erratic (0.2 .. 0.66 ms run to run)
x = torch.zeros((L, nb + 8), dtype=torch.float64, device="cuda")[:, :nb]
x.copy_(torch.randn((L, nb), dtype=torch.float64, device="cuda"))
xt = x.T.contiguous()
for _ in range(3):
    y0 = ops.dct1(plan, ops.BWD, x, axis=0)
    y1 = ops.dct1(plan, ops.BWD, xt, axis=1)
b = Base(2048, "CD")
c = torch.randn((2046, 2048), dtype=torch.float64, device="cuda")
u = torch.randn((2048, 2048), dtype=torch.float64, device="cuda")
for _ in range(2):
    b.derivative(c, 1, axis=0)
    b.derivative(c.T.contiguous(), 1, axis=1)
    b.from_chebyshev(u, axis=0)
    b.from_chebyshev(u, axis=1)
torch.cuda.synchronize()
from pypde_b200 import _cabi as C  # noqa: E402


def padded(r, c):
    return torch.zeros((r, c + 8), dtype=torch.float64, device="cuda")[:, :c]


for axis, arr in ((0, x), (1, xt)):
    out = padded(*arr.shape)                    # same shape: L -> L along the transform axis
    ldx, ldy = arr.stride(0), out.stride(0)

    def run():
        C.check(C.lib().pde_dct1(plan.handle, ops.BWD, C.p(arr), ldx, L, C.p(out), ldy, L, nb, axis, C.stream()))

    for rep in range(3):
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("dct axis %d: %.3f ms  %.1f GB/s algorithmic" % (axis, ms, 16.0 * L * nb / ms / 1e6))

from pypde_b200 import PlanLHS
import numpy as np
n = 2046
rng = np.random.default_rng(0)
A = np.zeros((n, n))
for off in (-2, 0, 2, 4):
    A += np.diag(rng.standard_normal(n - abs(off)) * 0.3 + (3.0 if off == 0 else 0.0), off)
xs = torch.randn((n, n), dtype=torch.float64, device="cuda")


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for axis in (0, 1):
    plan = PlanLHS(A, ndim=2, axis=axis, method="fdma")
    print("fdma axis %d: %.3f ms" % (axis, timeit(lambda: plan.solve(xs))))
    print("diff axis %d: %.3f ms" % (axis, timeit(lambda: b.derivative(c, 1, axis=axis))))
    uu = u if axis == 0 else u.T.contiguous()
    print("from_cheb axis %d: %.3f ms" % (axis, timeit(lambda: b.from_chebyshev(uu, axis=axis))))
