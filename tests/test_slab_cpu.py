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


# ---------------------------------------------------------------------------------------------------
# peer-memory slab decomposition (navier/peer.py): layout arithmetic and the segment description of the
# row passes, emulated with numpy buffers as "peer memory"; rendezvous of the IPC handles over gloo.
# ---------------------------------------------------------------------------------------------------
def test_slab_layout_and_row_segments():
    from pypde_b200.navier.peer import SlabLayout, partition as part4, row_segments
    assert part4(2048, 8, 4) == [(256 * r, 256) for r in range(8)]
    assert [w for _, w in part4(66, 4, 4)] == [20, 16, 16, 14] and sum(w for _, w in part4(66, 4, 4)) == 66
    for N0, N1, D0, P in ((64, 64, 97, 2), (48, 64, 73, 4), (2048, 2048, 3073, 8), (130, 132, 197, 3)):
        lays = [SlabLayout(N0, N1, D0, D0, P, r) for r in range(P)]
        assert sum(l.W for l in lays) == N1 and sum(l.N0r for l in lays) == N0 and sum(l.D0r for l in lays) == D0
        assert sum(l.M1c for l in lays) == N1 - 2 and sum(l.M0r for l in lays) == N0 - 2
        assert all(l.c0 % 4 == 0 and l.W % 2 == 0 for l in lays)
        # emulate: every rank owns a (rows x Wmax) slab of a global (N0 x ncols) array inside a flat buffer
        W = lays[0].Wmax
        rng = np.random.default_rng(P)
        for ncols in (N1, N1 - 2):
            G = rng.standard_normal((N0, ncols))
            offset = 256
            bufs = []
            for l in lays:
                b = np.zeros(offset // 8 + N0 * W)
                w = l.ncols_local(ncols)
                b[offset // 8:].reshape(N0, W)[:, :w] = G[:, l.c0:l.c0 + w]
                bufs.append(b)
            bases = [1000000 * (s + 1) for s in range(P)]          # fake addresses, 8-byte "memory" = bufs
            for l in lays:
                ptrs, lds, starts = row_segments(bases, offset, W, l.r0, l.col_starts(ncols), ncols)
                assert starts[-1] == ncols and all(s % 2 == 0 for s in starts[:-1])
                for q in (0, l.N0r - 1):
                    row = np.empty(ncols)
                    for i in range(ncols):
                        s = max(k for k in range(P) if starts[k] <= i)
                        addr = ptrs[s] + 8 * (q * lds[s] + i - starts[s])       # the kernel's formula
                        row[i] = bufs[s][(addr - bases[s]) // 8]
                    assert np.array_equal(row, G[l.r0 + q])
    with pytest.raises(ValueError):
        SlabLayout(16, 16, 25, 25, 8, 0)


def _handles_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pypde_b200.navier.peer import exchange_handles
        mine = bytes([rank]) * 64
        got = exchange_handles(mine)
        out[rank] = got == [bytes([r]) * 64 for r in range(world)]
    finally:
        dist.destroy_process_group()


def test_ipc_handle_rendezvous_gloo():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_handles_worker, args=(2, port, out), nprocs=2, join=True)
    assert all(out[r] for r in range(2))
