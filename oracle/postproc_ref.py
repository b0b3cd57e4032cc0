"""
TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the evaluation post-processing that follows `transcribe` in the
reference.  Imported by tests/ only; the product path is csrc/postproc_kernels.cu.

  filter_non_peaks / threshold     timbre_trap/utils/processing.py:66-124.  PINNED: tests/golden/postproc.npz holds the outputs
                                   of the reference's own functions (scripts/make_golden.py imports them; scipy is in the image).
  binary_map                       PitchDataset.activations_to_multi_pitch (datasets/PitchDataset.py:309-349) up to the point
                                   where it turns bins into Hz lists, plus the bin mask of experiments/evaluate.py:48.
  multipitch_counts / prf          mir_eval.multipitch as driven by utils/experiments.py:354-396.  mir_eval is not in the image
                                   (PARITY UNPINNED for this function): restated from its published algorithm - per frame the
                                   size of a maximum bipartite matching between reference and estimated pitches whose distance is
                                   within the window (0.5 semitone), precision = sum tp / sum est, recall = sum tp / sum ref,
                                   f1 = 2pr / (p + r + eps).  On a common frame grid and with pitches quantised to CQT bins the
                                   distance is a bin difference, and in one dimension the greedy two-cursor pairing is a maximum
                                   matching.
"""

import sys

import numpy as np


def filter_non_peaks(arr):
    """processing.py:66-98: values that are strict local maxima along axis -2 (zero rows beyond the edges), else 0."""
    a = np.asarray(arr, dtype=np.float64)
    pad = [(0, 0)] * a.ndim
    pad[-2] = (1, 1)
    p = np.pad(a, pad)
    mid, lo, hi = p[..., 1:-1, :], p[..., :-2, :], p[..., 2:, :]
    return np.where((mid > lo) & (mid > hi), mid, 0.0)


def threshold(arr, t=0.5):
    """processing.py:101-124."""
    return (np.asarray(arr) >= t).astype(np.float64)


def binary_map(activations, t=0.5, peaks_only=False, bin_lo=0, bin_hi=None):
    """(..., F, T) activations -> uint8 map of active bins (PitchDataset.py:335-340, restricted to bins [bin_lo, bin_hi))."""
    a = np.asarray(activations, dtype=np.float32)
    keep = a >= np.float32(t)
    if peaks_only:
        keep &= _strict_peaks(a)
    F = a.shape[-2]
    mask = np.zeros(F, dtype=bool)
    mask[bin_lo: F if bin_hi is None else bin_hi] = True
    return (keep & mask[:, None]).astype(np.uint8)


def _strict_peaks(a):
    pad = [(0, 0)] * a.ndim
    pad[-2] = (1, 1)
    p = np.pad(a, pad)
    return (p[..., 1:-1, :] > p[..., :-2, :]) & (p[..., 1:-1, :] > p[..., 2:, :])


def multipitch_counts(est, ref, tol_bins):
    """est, ref: (F, T) {0,1} maps on a common frame grid -> (true positives, n_est, n_ref)."""
    est, ref = np.asarray(est) != 0, np.asarray(ref) != 0
    tp = 0
    for e, r in zip(est.T, ref.T):
        ei, ri = np.flatnonzero(e), np.flatnonzero(r)
        i = j = 0
        while i < len(ei) and j < len(ri):
            if abs(int(ei[i]) - int(ri[j])) <= tol_bins:
                tp += 1; i += 1; j += 1
            elif ei[i] < ri[j]:
                i += 1
            else:
                j += 1
    return tp, int(est.sum()), int(ref.sum())


def prf(tp, n_est, n_ref):
    """precision / recall as mir_eval.multipitch computes them, f1 as utils/experiments.py:389."""
    p = tp / n_est if n_est else 0.0
    r = tp / n_ref if n_ref else 0.0
    return p, r, 2 * p * r / (p + r + sys.float_info.epsilon)
