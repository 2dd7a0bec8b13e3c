"""ncu driver: one eager rbc2048 stage (all batched kernels at bench sizes)."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
from pypde_b200.navier import rbc2d
cfg = bench.WORKLOADS[os.environ.get("WL", "rbc2048")]
ns = rbc2d.NavierStokes(**cfg)
bench.init_state(ns, cfg["shape"])
for _ in range(2):
    ns.update()
torch.cuda.synchronize()
