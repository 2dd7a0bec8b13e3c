"""Multi-GPU check of the slab-decomposed stepper (run with torchrun, one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_slab_check.py
Every rank steps the same problem; rank 0 compares the gathered state with the CPU oracle and
with the single-GPU batched stepper."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rel_l2  # noqa: E402
from test_oracle_cpu import _cases  # noqa: E402
from test_gpu_rbc import make, make_oracle, H  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    for name, steps in (("rbc64_rk3_dealias", 10), ("rbc48x64_aspect2", 5), ("rbc64_eu_nodealias", 5),
                        ("rbc128_rk3_dealias", 3)):
        cfg = _cases()[name]
        ns = make(cfg, slab=True)
        ref = make(cfg)
        for _ in range(steps):
            ns.update()
            ref.update()
        ns.sync_fields()
        torch.cuda.synchronize()
        if rank == 0:
            o = make_oracle(cfg)
            o.iterate(steps)
            for k, t, s, r in (("T", ns.T.vhat, ref.T.vhat, o.That_), ("U", ns.U.vhat, ref.U.vhat, o.Uhat),
                               ("V", ns.V.vhat, ref.V.vhat, o.Vhat), ("pres", ns.pres.vhat, ref.pres.vhat, o.pres)):
                e_o, e_s = rel_l2(H(t), r), rel_l2(H(t), H(s))
                worst = max(worst, e_o)
                print("%-20s world %d  %-4s slab-vs-oracle %.2e  slab-vs-single %.2e" % (name, world, k, e_o, e_s))
        dist.barrier()
    if rank == 0:
        print("transposes per step:", ns._fast.comm.calls // max(1, steps), "WORST", worst)
        assert worst < 1e-12, worst
        print("SLAB OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
