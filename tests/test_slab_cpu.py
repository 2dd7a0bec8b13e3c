"""CPU (gloo, world_size 2 and 3): the slab-decomposition plumbing of pypde_b200.navier.slab -
balanced partitions (uneven last slabs), bundled X<->Y distributed transposes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pypde_b200.navier.slab import SlabComm, partition


def test_partition_uneven():
    p = partition(2046, 8)
    assert sum(s for _, s in p) == 2046 and p[0] == (0, 256) and p[-1][1] == 255
    assert partition(5, 8)[5:] == [(5, 0)] * 3
    off = 0
    for o, s in partition(3073, 4):
        assert o == off
        off += s


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, rows, cols, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = SlabComm()
        g = torch.Generator().manual_seed(7)
        full = torch.randn((K, rows, cols), dtype=torch.float64, generator=g)     # K global arrays
        rp, cp = partition(rows, world), partition(cols, world)
        c0, cw = cp[rank]
        r0, rw = rp[rank]
        xb = torch.cat([full[k][:, c0:c0 + cw] for k in range(K)], dim=1).contiguous()   # (rows, K*cw)
        yb = torch.empty((rw, K * cols), dtype=torch.float64)
        comm.x2y(xb, yb, K, rows, cols)
        ok = all(torch.equal(yb[:, k * cols:(k + 1) * cols], full[k][r0:r0 + rw]) for k in range(K))
        xb2 = torch.empty_like(xb)
        comm.y2x(yb, xb2, K, rows, cols)
        ok = ok and torch.equal(xb2, xb) and comm.calls == 2
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rows,cols,K", [(2, 10, 7, 3), (3, 11, 8, 1), (2, 64, 62, 5)])
def test_bundled_transposes_gloo(world, rows, cols, K):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, rows, cols, K, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
