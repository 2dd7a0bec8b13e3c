"""Offline shared-memory wavefront model of k_dct_fft_t (pypde_b200/csrc/dct_fft_t.cuh).

A 16-byte (double2) shared access of a warp is served per quarter-warp (8 lanes = 128 bytes);
a quarter needs as many wavefronts as the largest number of DISTINCT 16-byte words that fall
into the same bank group (word index mod 8).  The script replays the thread -> element maps
of the load phase, every DIF pass and the split phase and prints wavefronts vs the ideal
(one per quarter-warp), for the old and the new maps.  No GPU needed.
"""
import sys
from collections import defaultdict


def wavefronts(addrs):
    """addrs: list over lanes (None = inactive) of 16-byte word indices"""
    tot = 0
    for q in range(0, len(addrs), 8):
        groups = defaultdict(set)
        for a in addrs[q:q + 8]:
            if a is not None:
                groups[a % 8].add(a)
        if groups:
            tot += max(len(v) for v in groups.values())
    return tot


def ideal(addrs):
    return sum(1 for q in range(0, len(addrs), 8) if any(a is not None for a in addrs[q:q + 8]))


class Layout:
    def __init__(self, P, radices, S, T, seq_pad, remap):
        self.P, self.rad, self.S, self.T = P, radices, S, T
        self.M1 = P // radices[0]
        self.PS = P + P // self.M1 + seq_pad
        self.remap = remap

    def phys(self, i):
        return i + i // self.M1


def sim(P, radices, S, T, axis, seq_pad=0, remap=False, verbose=True):
    L = Layout(P, radices, S, T, seq_pad, remap)
    res = []
    # ---- load phase (full CTA): one 16-byte store per element
    w = i = 0
    tot = S * P
    for it in range(tot // T):
        for w0 in range(0, T, 32):
            ad = []
            for lane in range(32):
                t = w0 + lane
                if axis == 1:
                    if P % T == 0:
                        s, m = it // (P // T), (it % (P // T)) * T + t
                    else:
                        s, m = it * (T // P) + t // P, t % P
                else:
                    C = T // S
                    s, m = t % S, it * C + t // S
                ad.append(s * L.PS + L.phys(m))
            w += wavefronts(ad)
            i += ideal(ad)
    res.append(("load", w, i))
    # ---- passes
    ncur = P
    for pi, R in enumerate(radices):
        M = ncur // R
        per_seq = P // R
        total = S * per_seq
        rs = M + 1 if ncur == P else M
        w = i = 0
        nblk = P // ncur
        for b0 in range(0, total, T):
            for w0 in range(0, T, 32):
                ads = [[] for _ in range(R)]
                for lane in range(32):
                    b = b0 + w0 + lane
                    if b >= total:
                        for r in range(R):
                            ads[r].append(None)
                        continue
                    s, bb = divmod(b, per_seq)
                    if remap and ncur != P:
                        blk, j = remap_map(bb, M, nblk, L)
                    else:
                        blk, j = divmod(bb, M)
                    i0 = blk * ncur + j
                    p = s * L.PS + (i0 if ncur == P else L.phys(i0))
                    for r in range(R):
                        ads[r].append(p + r * rs)
                for r in range(R):
                    w += 2 * wavefronts(ads[r])      # read + write
                    i += 2 * ideal(ads[r])
        res.append(("pass%d r%d" % (pi + 1, R), w, i))
        ncur = M
    # ---- split
    H = P // 2
    R1 = radices[0]

    def pos(k):
        n, out = P, 0
        for R in radices:
            n //= R
            out += (k % R) * n
            k //= R
        return out
    w = i = 0
    for it in range(S * H // T):
        for w0 in range(0, T, 32):
            a, b = [], []
            for lane in range(32):
                t = w0 + lane
                if axis == 1:
                    if H % T == 0:
                        s, k = it // (H // T), (it % (H // T)) * T + t
                    else:
                        s, k = it * (T // H) + t // H, t % H
                else:
                    C = T // S
                    s, k = t % S, it * C + t // S
                kk = 0 if k == 0 else P - k
                a.append(s * L.PS + pos(k) + k % R1)
                b.append(s * L.PS + pos(kk) + kk % R1)
            w += wavefronts(a) + wavefronts(b)
            i += ideal(a) + ideal(b)
    res.append(("split", w, i))
    tw = sum(r[1] for r in res)
    ti = sum(r[2] for r in res)
    if verbose:
        print("P=%d rad=%s S=%d T=%d axis=%d seq_pad=%d remap=%s  PS=%d" % (P, radices, S, T, axis, seq_pad, remap, L.PS))
        for name, ww, ii in res:
            print("   %-10s wavefronts %8d  ideal %8d  x%.2f" % (name, ww, ii, ww / ii))
        print("   %-10s wavefronts %8d  ideal %8d  x%.2f" % ("total", tw, ti, tw / ti))
    return tw, ti


def remap_map(bb, M, nblk, L):
    """new thread -> (block, j) map of the passes after the first: consecutive butterflies walk the
    blocks with a stride of one first-level block (M1 elements + 1 pad = odd), j is the slow index."""
    ncur_blocks = nblk                      # blocks of the current level, per sequence
    # blocks per first-level block
    R1 = L.rad[0]
    per_first = ncur_blocks // R1           # current-level blocks inside one first-level block
    j, bi = divmod(bb, ncur_blocks)
    # bi enumerates blocks: fastest index = which first-level block (stride M1+1 in memory)
    inner, first = divmod(bi, R1)
    blk = first * per_first + inner
    return blk, j


if __name__ == "__main__":
    cases = [(3072, (16, 16, 4, 3), 1, 192, 1), (3072, (16, 16, 4, 3), 4, 384, 0),
             (2048, (16, 16, 8), 1, 128, 1), (2048, (16, 16, 8), 4, 256, 0),
             (1536, (16, 16, 2, 3), 2, 192, 1), (768, (16, 16, 3), 4, 192, 1),
             (6144, (16, 16, 8, 3), 1, 384, 1), (6144, (16, 16, 8, 3), 2, 384, 0),
             (4096, (16, 16, 16), 1, 256, 1), (4096, (16, 16, 16), 2, 256, 0), (96, (16, 2, 3), 32, 192, 1)]
    for P, rad, S, T, axis in cases:
        sim(P, rad, S, T, axis)
        pad = 0
        if axis == 0 and S > 1:
            pad = (8 // S - (P + P // (P // rad[0])) % 8) % 8
        sim(P, rad, S, T, axis, seq_pad=pad, remap=True)
