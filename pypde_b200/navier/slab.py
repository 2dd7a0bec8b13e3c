"""
Slab decomposition of the rbc2d time step over the GPUs of one node.

Every operator of the step acts along ONE axis and is independent along the other, so
  * axis-0 operators run in the X layout: every rank holds all rows and a block of columns,
  * axis-1 operators (including the two dense projections with Hy, Qy) run in the Y layout:
    every rank holds a block of rows and all columns,
and a change of axis is a distributed transpose (NCCL all-to-all over NVLink).  Arrays that
travel together are stored as ONE "bundle" tensor with their columns side by side
((rows, K * cols), array k = bundle[:, k*cols:(k+1)*cols], leading dimension K*cols), so that

  X -> Y: the block for rank s is the contiguous row range of s (zero-copy send); the receiver
          unpacks the P blocks into its (rows_s, K * all columns) bundle;
  Y -> X: the sender packs the column ranges per destination; the receiver gets row blocks that
          land contiguously in its (all rows, K * cols_s) bundle (zero-copy receive).

All spectral arrays are padded to the common kind (N0, N1) (Galerkin arrays have 2 unused
rows / columns), dealiased ones to (D0, N1), so one row partition (of N0, D0) and one column
partition (of N1) serve every exchange.  10 exchanges per IMEX stage.
"""
import ctypes

import torch
import torch.distributed as dist


def partition(n, parts):
    """Balanced contiguous split: list of (offset, size)."""
    base, rem = divmod(n, parts)
    out, off = [], 0
    for r in range(parts):
        sz = base + (1 if r < rem else 0)
        out.append((off, sz))
        off += sz
    return out


class SlabComm:
    """Distributed transposes between the X and Y layouts for bundles of K arrays."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.bytes_sent = 0
        self.calls = 0

    def _repack(self, direction, bundle, blocked, rows, K, cols, cp):
        """One kernel (pde_slab_repack) on CUDA tensors; per-rank torch copies otherwise (gloo tests)."""
        if bundle.is_cuda:
            from .. import _cabi as C
            off = (ctypes.c_int * (self.size + 1))(*([o for o, _ in cp] + [cols]))
            C.check(C.lib().pde_slab_repack(direction, C.p(bundle), C.p(blocked), rows, K, cols, self.size, off,
                                            C.stream()))
            return
        b3 = bundle.view(rows, K, cols)
        pos = 0
        for s in range(self.size):
            o, w = cp[s]
            n = rows * K * w
            if n:
                blk = blocked[pos:pos + n].view(rows, K, w)
                if direction:
                    blk.copy_(b3[:, :, o:o + w])
                else:
                    b3[:, :, o:o + w].copy_(blk)
            pos += n

    def x2y(self, xb, yb, K, rows, cols):
        """xb: (rows, K*cols_r) X bundle of this rank; yb: (rows_r, K*cols) Y bundle (output)."""
        P, r = self.size, self.rank
        rp, cp = partition(rows, P), partition(cols, P)
        assert xb.is_contiguous() and yb.is_contiguous()
        assert tuple(xb.shape) == (rows, K * cp[r][1]) and tuple(yb.shape) == (rp[r][1], K * cols)
        in_split = [rp[s][1] * K * cp[r][1] for s in range(P)]
        out_split = [rp[r][1] * K * cp[s][1] for s in range(P)]
        recv = torch.empty(sum(out_split), dtype=xb.dtype, device=xb.device)
        dist.all_to_all_single(recv, xb.reshape(-1), out_split, in_split, group=self.group)
        self.bytes_sent += (sum(in_split) - in_split[r]) * 8
        self.calls += 1
        self._repack(0, yb, recv, rp[r][1], K, cols, cp)
        return yb

    def y2x(self, yb, xb, K, rows, cols):
        """yb: (rows_r, K*cols) Y bundle; xb: (rows, K*cols_r) X bundle (output)."""
        P, r = self.size, self.rank
        rp, cp = partition(rows, P), partition(cols, P)
        assert xb.is_contiguous() and yb.is_contiguous()
        assert tuple(xb.shape) == (rows, K * cp[r][1]) and tuple(yb.shape) == (rp[r][1], K * cols)
        in_split = [rp[r][1] * K * cp[s][1] for s in range(P)]
        out_split = [rp[s][1] * K * cp[r][1] for s in range(P)]
        send = torch.empty(sum(in_split), dtype=yb.dtype, device=yb.device)
        self._repack(1, yb, send, rp[r][1], K, cols, cp)
        dist.all_to_all_single(xb.reshape(-1), send, out_split, in_split, group=self.group)
        self.bytes_sent += (sum(in_split) - in_split[r]) * 8
        self.calls += 1
        return xb
