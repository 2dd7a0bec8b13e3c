"""Multi-GPU check of the slab-decomposed stepper (run with torchrun, one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_slab_check.py [--big]
Every rank steps the same problem; rank 0 compares the gathered state with the CPU oracle and with the
single-GPU stepper.  --big adds rbc512 (3 steps, oracle) and rbc2048 (2 steps, slab against the single-GPU
stepper: SURVEY.md §8(d) "1-GPU and 8-GPU runs of the new code").  Graph replay is exercised as well."""
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


def states(ns):
    return (("T", ns.T.vhat), ("U", ns.U.vhat), ("V", ns.V.vhat), ("pres", ns.pres.vhat))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    big = "--big" in sys.argv
    worst = 0.0
    cases = [("rbc64_rk3_dealias", 10, False), ("rbc48x64_aspect2", 5, False), ("rbc64_eu_nodealias", 5, True),
             ("rbc128_rk3_dealias", 3, True)]
    for name, steps, graph in cases:
        cfg = _cases()[name]
        if cfg["shape"][1] // 4 < world or cfg["shape"][0] < 2 * world:
            continue
        ns = make(cfg, slab=True, graph=graph)
        ref = make(cfg)
        for _ in range(steps):
            ns.update()
            ref.update()
        ns.sync_fields()
        torch.cuda.synchronize()
        kind = type(ns._fast).__name__
        if hasattr(ns._fast, "check"):
            ns._fast.check()
        if rank == 0:
            o = make_oracle(cfg)
            o.iterate(steps)
            for (k, t), (_, s), r in zip(states(ns), states(ref), (o.That_, o.Uhat, o.Vhat, o.pres)):
                e_o, e_s = rel_l2(H(t), r), rel_l2(H(t), H(s))
                worst = max(worst, e_o)
                print("%-20s %s world %d graph %d  %-4s slab-vs-oracle %.2e  slab-vs-single %.2e" % (
                    name, kind, world, graph, k, e_o, e_s), flush=True)
        ns.close()
        dist.barrier()
    if big:
        from test_gpu_large import _cfg
        for tag, cfg, steps, oracle in (("rbc512", _cfg(512, 1e8, 1e-3), 3, True), ("rbc2048", _cfg(2048, 1e10, 1e-4), 2, False)):
            ns = make(cfg, slab=True, graph=True)
            for _ in range(steps):
                ns.update()
            ns.sync_fields()
            torch.cuda.synchronize()
            if hasattr(ns._fast, "check"):
                ns._fast.check()
            if rank == 0:
                ref = make(cfg)
                for _ in range(steps):
                    ref.update()
                torch.cuda.synchronize()
                o = None
                if oracle:
                    o = make_oracle(cfg)
                    o.iterate(steps)
                for n, ((k, t), (_, s)) in enumerate(zip(states(ns), states(ref))):
                    e_s = rel_l2(H(t), H(s))
                    msg = "%-8s world %d  %-4s slab-vs-single %.2e" % (tag, world, k, e_s)
                    if o is not None:
                        e_o = rel_l2(H(t), (o.That_, o.Uhat, o.Vhat, o.pres)[n])
                        msg += "  slab-vs-oracle %.2e  single-vs-oracle %.2e" % (e_o, rel_l2(H(s), (o.That_, o.Uhat, o.Vhat, o.pres)[n]))
                    print(msg, flush=True)
                    # two valid roundings of a step whose own half-ulp response is 5e-13 (512) ... 3e-11 (2048)
                    assert e_s < (5e-12 if tag == "rbc512" else 2e-10), (tag, k, e_s)
                del ref
            ns.close()
            dist.barrier()
    if rank == 0:
        print("WORST", worst)
        assert worst < 1e-12, worst
        print("SLAB OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
