"""
Host-side construction of the NSGT filter-bank tables the CUDA kernels are driven by.

This is the part of `cqt_pytorch.CQT.__init__` the reference relies on through
timbre_trap/framework/cqtwrapper.py:31-35 (window lengths, positions, centred Hann windows,
dual windows, `max_window_length`).  The package itself is an un-vendored, un-pinned
dependency of the reference (requirements.txt:15), so the construction follows the published
constant-Q NSGT design (Velasco et al. 2011; Holighaus et al. 2013) - see DESIGN.md
"parity unpinned".  Everything is computed in float64 and rounded once to fp32.

Output is in the packed form of include/timbre_trap_b200.h::tt_cqt_plan_create: for bin k,
`length[k]` non-zero window taps starting at crop index `first[k]` / spectrum index `start[k]`.
"""

import math

import numpy as np

__all__ = ['FilterBank']


def _periodic_hann(n):
    if n == 1:
        return np.ones(1)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


class FilterBank:
    def __init__(self, n_octaves, bins_per_octave, sample_rate, block_length, power_of_2_length=True):
        n_bins = int(n_octaves) * int(bins_per_octave)
        L = int(block_length)
        nyquist = float(sample_rate) / 2.0
        centres = (nyquist / 2.0 ** n_octaves) * np.exp2(np.arange(n_bins, dtype=np.float64) / bins_per_octave)
        spread = 2.0 ** (1.0 / bins_per_octave) - 2.0 ** (-1.0 / bins_per_octave)
        to_taps = L / float(sample_rate)

        # window lengths (the Nyquist band only takes part in sizing the crop) and centre positions
        lengths = np.maximum(np.rint(centres * spread * to_taps), 1).astype(np.int64)
        nyquist_len = max(int(np.rint(nyquist * spread * to_taps)), 1)
        positions = np.rint(centres * to_taps).astype(np.int64)

        crop = max(int(lengths.max()), nyquist_len)
        if power_of_2_length:
            crop = 1 << int(math.ceil(math.log2(crop)))

        first = np.floor(crop / 2.0 - lengths / 2.0).astype(np.int64)
        start = (positions - crop // 2 + first) % L
        offset = np.zeros(n_bins + 1, dtype=np.int64)
        offset[1:] = np.cumsum(lengths)

        win = np.concatenate([_periodic_hann(int(n)) for n in lengths])
        # dual = window / (sum over all bins of window^2 at the same spectrum position)
        where = np.concatenate([(start[k] + np.arange(lengths[k])) % L for k in range(n_bins)])
        diagonal = np.zeros(L, dtype=np.float64)
        np.add.at(diagonal, where, win ** 2)
        covered = diagonal[where] > 0
        dual = np.where(covered, win / np.where(covered, diagonal[where], 1.0), 0.0)

        self.n_bins = n_bins
        self.block_length = L
        self.max_window_length = crop
        self.start = start.astype(np.int32)
        self.length = lengths.astype(np.int32)
        self.first = first.astype(np.int32)
        self.offset = offset.astype(np.int32)
        self.win = np.ascontiguousarray(win, dtype=np.float32)
        self.dual = np.ascontiguousarray(dual, dtype=np.float32)
        self.n_taps = int(offset[-1])
