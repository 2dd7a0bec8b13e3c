"""
Peer memory of the slab decomposition: one process per GPU on one node, every rank's exchange buffers
mapped into every other rank's address space (CUDA IPC over NVLink / NVSwitch), plus the device-side
barrier that separates the passes.  torch.distributed is used for the rendezvous only (exchange of the
64-byte IPC handles); the data path has no NCCL call: the row passes of the stepper load from and store
to the peers' slabs directly (csrc/pass.cu, segmented operands).

Layout arithmetic (`SlabLayout`) is plain Python and is unit-tested on CPU (tests/test_slab_cpu.py).
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _cabi as C


def partition(n, parts, unit=1):
    """Balanced contiguous split of n items in multiples of `unit` (the last part takes the remainder):
    list of (offset, size)."""
    blocks = -(-n // unit)
    base, rem = divmod(blocks, parts)
    out, off = [], 0
    for r in range(parts):
        sz = (base + (1 if r < rem else 0)) * unit
        sz = max(0, min(sz, n - off))
        out.append((off, sz))
        off += sz
    return out


class SlabLayout:
    """Who owns what.  X layout: all rows, my block of columns (axis-0 operators); Y layout: my block of
    rows, all columns (axis-1 operators).  Column blocks are multiples of 4 (16-byte units of the row
    passes, strips of the column passes)."""

    def __init__(self, N0, N1, D0, D1, P, r):
        self.N0, self.N1, self.D0, self.D1, self.P, self.r = N0, N1, D0, D1, P, r
        self.M0, self.M1 = N0 - 2, N1 - 2
        self.cp = partition(N1, P, 4)            # columns of the X layout
        self.rp = partition(N0, P)               # rows of the Y layout (N0-row arrays)
        self.dp = partition(D0, P)               # rows of the Y layout (dealiased arrays)
        self.c0, self.W = self.cp[r]
        self.r0, self.N0r = self.rp[r]
        self.d0, self.D0r = self.dp[r]
        self.Wmax = max(w for _, w in self.cp)
        self.M1c = max(0, min(self.c0 + self.W, self.M1) - self.c0)     # my valid Galerkin columns
        self.M0r = max(0, min(self.r0 + self.N0r, self.M0) - self.r0)   # my valid Galerkin rows
        for s in range(P):
            c0, w = self.cp[s]
            if min(w, self.rp[s][1], self.dp[s][1]) <= 0 or min(c0 + w, self.M1) - c0 <= 0 \
                    or min(self.rp[s][0] + self.rp[s][1], self.M0) - self.rp[s][0] <= 0:
                raise ValueError("grid %dx%d too small for %d slabs" % (N0, N1, P))

    def col_starts(self, ncols):
        """segment starts of a row with `ncols` valid columns"""
        return [c0 for c0, _ in self.cp] + [ncols]

    def ncols_local(self, ncols, s=None):
        c0, w = self.cp[self.r if s is None else s]
        return max(0, min(c0 + w, ncols) - c0)


def row_segments(bases, offset, ld, row0, col_starts, ncols):
    """Segment description of the rows of an X-layout array as a row pass on another rank sees them:
    rank s keeps its columns [col_starts[s], col_starts[s+1]) of every row in a (rows x ld) matrix at
    `offset` bytes into its buffer `bases[s]`; the pass numbers its rows from `row0`.
    Returns (pointers, leading dimensions, element starts) for pde_pass_ins / passes.Operand: element i of
    local row q lives at ptr[s] + 8 (q ld + i - start[s])."""
    ptrs = [b + offset + 8 * row0 * ld for b in bases]
    starts = list(col_starts[:len(bases)]) + [ncols]
    return ptrs, [ld] * len(bases), starts


def exchange_handles(handle, group=None):
    """All ranks' 64-byte IPC handles, in rank order (torch.distributed object all-gather: works on gloo and nccl)."""
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(handle), group=group)
    assert all(isinstance(h, bytes) and len(h) == len(out[0]) for h in out)
    return out


class PeerMem:
    """One cudaMalloc'ed buffer per rank, opened by every other rank."""

    def __init__(self, nbytes, group=None):
        self.group = group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes)
        L = C.lib()
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        C.check(L.pde_ipc_alloc(ctypes.byref(ptr), self.nbytes, handle))
        self.local = ptr.value
        handles = exchange_handles(handle.raw, group)
        self.base = []
        for s, h in enumerate(handles):
            if s == self.rank:
                self.base.append(self.local)
            else:
                q = ctypes.c_void_p()
                C.check(L.pde_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q)))
                self.base.append(q.value)
        # barrier state: slots [0, size) of the buffer are the flags
        dev = C.device()
        self.flag_ptrs = torch.tensor(self.base, dtype=torch.int64, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.closed = False
        torch.cuda.synchronize()
        dist.barrier(group=group)

    HEADER = 256        # bytes reserved for the flags

    def view(self, offset, shape):
        """torch view of a part of the LOCAL buffer"""
        n = 1
        for s in shape:
            n *= s
        assert offset % 16 == 0 and offset + 8 * n <= self.nbytes

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (self.local + offset, False),
                                        "version": 2, "strides": None}
        t = torch.as_tensor(raw, device=C.device())
        assert t.data_ptr() == self.local + offset
        return t

    def barrier_args(self):
        """(C function, args) of the device-side barrier for a launch list"""
        return (C.lib().pde_peer_barrier, (ctypes.c_void_p(self.flag_ptrs.data_ptr()), ctypes.c_void_p(self.epoch.data_ptr()),
                                           self.rank, self.size, ctypes.c_void_p(self.err.data_ptr())))

    def check(self):
        if int(self.err.item()) != 0:
            raise C.PdeError("pde_peer_barrier: a rank did not arrive (timeout)")

    def close(self):
        if self.closed:
            return
        self.closed = True
        torch.cuda.synchronize()
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        L = C.lib()
        for s, b in enumerate(self.base):
            if s != self.rank:
                L.pde_ipc_close(ctypes.c_void_p(b))
        torch.cuda.synchronize()
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        L.pde_ipc_free(ctypes.c_void_p(self.local))
