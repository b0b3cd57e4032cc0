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
    @classmethod
    def from_buffers(cls, block_length, windows, windows_range_indices, windows_inverse):
        """
        Tables from the buffers a checkpoint of the reference carries under `sliCQ.*` (registered by `cqt_pytorch.CQT`; the
        reference pickles whole modules, experiments/train.py:511, so every published / trained checkpoint holds them):

          windows                (F, M)          zero-padded analysis windows
          windows_range_indices  (F, M) integer  spectrum index of every crop tap (any representative mod block_length)
          windows_inverse        (F, M) per-tap dual windows, or (block_length,) a per-spectrum-position factor applied after
                                 the overlap-add (then the per-tap dual is windows * windows_inverse[index])

        With these the kernels reproduce the transform the checkpoint was trained with, whatever the construction choices of
        the package version that produced it (DESIGN.md "parity unpinned").  Raises ValueError when the buffers do not describe
        a filter bank the kernels can run (taps of a bin must be consecutive spectrum positions inside [0, L/2]).
        """
        win = np.asarray(windows, dtype=np.float64)
        idx = np.asarray(windows_range_indices).astype(np.int64)
        inv = np.asarray(windows_inverse, dtype=np.float64)
        L = int(block_length)
        if win.ndim != 2 or idx.shape != win.shape:
            raise ValueError(f'windows {win.shape} / windows_range_indices {idx.shape}: expected two (F, M) tables')
        F, M = win.shape
        if M & (M - 1):
            raise ValueError(f'max_window_length {M} is not a power of two (the reference constructs with power_of_2_length=True)')
        idx = idx % L
        if inv.shape == win.shape:
            dual_dense = inv
        elif inv.shape == (L,):
            dual_dense = win * inv[idx]
        else:
            raise ValueError(f'windows_inverse {inv.shape}: expected {(F, M)} or {(L,)}')
        start, length, first, packed_w, packed_d = [], [], [], [], []
        for k in range(F):
            nz = np.flatnonzero((win[k] != 0) | (dual_dense[k] != 0))
            lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (M // 2, M // 2 + 1)
            taps = idx[k, lo:hi]
            if np.any((taps - taps[0]) != np.arange(hi - lo)) or taps[-1] > L // 2:
                raise ValueError(f'bin {k}: the window taps are not consecutive one-sided spectrum positions')
            start.append(int(taps[0])); length.append(hi - lo); first.append(lo)
            packed_w.append(win[k, lo:hi]); packed_d.append(dual_dense[k, lo:hi])
        self = cls.__new__(cls)
        self.n_bins, self.block_length, self.max_window_length = F, L, M
        self.start = np.asarray(start, dtype=np.int32)
        self.length = np.asarray(length, dtype=np.int32)
        self.first = np.asarray(first, dtype=np.int32)
        self.offset = np.concatenate([[0], np.cumsum(length)]).astype(np.int32)
        self.win = np.ascontiguousarray(np.concatenate(packed_w), dtype=np.float32)
        self.dual = np.ascontiguousarray(np.concatenate(packed_d), dtype=np.float32)
        self.n_taps = int(self.offset[-1])
        self.source = 'checkpoint buffers'
        return self

    source = 'constructor arguments (restated NSGT design, parity unpinned)'

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
