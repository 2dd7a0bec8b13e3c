"""Host-side check of the shared-memory layout of the FFT DCT-I kernels (pypde_b200/csrc/dct_fft_t.cuh,
dct_bluestein.cuh): tools/sim_fft_banks.py replays the thread -> address map of every phase (load, DIF passes,
split) and counts 128-byte wavefronts per quarter-warp.  The thread -> butterfly map and the sequence pitch of
the kernels must stay conflict-free for every specialised transform length; the old map is kept in the
simulator as the reference point (ncu measured 10.2 M conflict wavefronts of 18.0 M with it)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import sim_fft_banks as sim  # noqa: E402

# (P, radices, S, T, axis) of dispatch_fft_t / dispatch_bluestein
CASES = [(3072, (16, 16, 12), 1, 192, 1), (3072, (16, 16, 12), 4, 384, 0), (1536, (16, 8, 12), 2, 192, 1),
         (1536, (16, 8, 12), 4, 384, 0), (3072, (16, 16, 4, 3), 1, 192, 1), (3072, (16, 16, 4, 3), 4, 384, 0),
         (6144, (16, 16, 8, 3), 1, 384, 1), (6144, (16, 16, 8, 3), 2, 384, 0), (2048, (16, 16, 8), 1, 128, 1),
         (2048, (16, 16, 8), 4, 256, 0), (4096, (16, 16, 16), 1, 256, 1), (4096, (16, 16, 16), 2, 256, 0),
         (768, (16, 16, 3), 4, 192, 1), (384, (16, 8, 3), 8, 192, 1), (1024, (16, 16, 4), 2, 128, 1),
         (512, (16, 16, 2), 4, 128, 1), (256, (16, 16), 8, 128, 1)]


def seq_pad(P, rad, S, axis):
    """SeqPad of dct_fft_t.cuh"""
    if axis != 0 or S <= 1 or 8 % S:
        return 0
    return (8 // S - (P + rad[0]) % 8) % 8


@pytest.mark.parametrize("P,rad,S,T,axis", CASES)
def test_fft_layout_is_conflict_free(P, rad, S, T, axis):
    new, ideal = sim.sim(P, rad, S, T, axis, seq_pad=seq_pad(P, rad, S, axis), remap=True, verbose=False)
    # the split phase reads pairs (k, P - k): a few quarter-warps straddle a first-level block (<= 5 % overall)
    assert new <= 1.05 * ideal, (new, ideal)


def test_old_map_had_the_conflicts_ncu_measured():
    old, ideal = sim.sim(3072, (16, 16, 4, 3), 4, 384, 0, remap=False, verbose=False)
    assert old > 1.8 * ideal
    old, ideal = sim.sim(2048, (16, 16, 8), 1, 128, 1, remap=False, verbose=False)
    assert old > 2.5 * ideal
