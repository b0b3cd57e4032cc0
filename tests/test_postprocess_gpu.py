"""
GPU parity of the evaluation post-processing (csrc/postproc_kernels.cu) through the C ABI: bit-exact
against oracle/postproc_ref.py and against the golden outputs of the reference's own numpy functions.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_filter_non_peaks_and_threshold_golden(golden_dir):
    from timbre_trap_b200.framework import postprocess as PP
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    a = torch.from_numpy(g['activations']).cuda()
    assert np.array_equal(PP.filter_non_peaks(a).cpu().numpy(), g['filter_non_peaks'])
    assert np.array_equal(PP.threshold(a, 0.5).cpu().numpy(), g['threshold'].astype(np.float32))
    assert np.array_equal(PP.activations_to_binary(a, 0.5).cpu().numpy(), g['threshold'])
    assert np.array_equal(PP.activations_to_binary(a, 0.5, peaks_only=True).cpu().numpy(), g['peaks_threshold'])
    assert np.array_equal(PP.filter_non_peaks(a[0]).cpu().numpy(), g['filter_non_peaks'][0])          # (F, T) input


@pytest.mark.parametrize('B,F,T', [(3, 540, 1000), (1, 72, 1), (2, 1, 300), (1, 540, 3073)])
def test_binary_map_and_counts_vs_oracle(B, F, T):
    from oracle import postproc_ref as R
    from timbre_trap_b200.framework import postprocess as PP
    rng = np.random.default_rng(B * 1000 + F + T)
    act = (np.round(rng.random((B, F, T)) * 32) / 32).astype(np.float32)
    ref = (rng.random((B, F, T)) < 0.02).astype(np.uint8)
    a = torch.from_numpy(act).cuda()
    for t, peaks, lo, hi in ((0.5, False, 0, None), (0.75, True, 0, None), (0.9, True, F // 10, F - F // 8)):
        got = PP.activations_to_binary(a, t, peaks, lo, hi).cpu().numpy()
        want = R.binary_map(act, t, peaks, lo, hi)
        assert got.dtype == np.uint8 and np.array_equal(got, want)
    est = R.binary_map(act, 0.9, True)
    for tol in (0, 2, 5):
        counts = PP.multipitch_counts(torch.from_numpy(est).cuda(), torch.from_numpy(ref).cuda(), tol).cpu().numpy()
        assert counts.shape == (B, 3) and counts.dtype == np.int64
        for b in range(B):
            assert tuple(counts[b]) == R.multipitch_counts(est[b], ref[b], tol)
    sc = PP.multipitch_scores(torch.from_numpy(est).cuda(), torch.from_numpy(ref).cuda(), 2)
    tp, ne, nr = (sum(R.multipitch_counts(est[b], ref[b], 2)[k] for b in range(B)) for k in range(3))
    p, r, f = R.prf(tp, ne, nr)
    assert sc['precision'] == p and sc['recall'] == r and sc['f1-score'] == f


def test_postprocess_edge_cases():
    from timbre_trap_b200._lib import TimbreTrapB200Error
    from timbre_trap_b200.framework import postprocess as PP
    z = torch.zeros(2, 12, 40, device='cuda')
    assert int(PP.activations_to_binary(z, 0.5).sum()) == 0 and int(PP.activations_to_binary(z, 0.0).sum()) == z.numel()
    assert int(PP.activations_to_binary(z, 0.0, peaks_only=True).sum()) == 0            # a plateau has no strict peak
    ones = torch.ones(1, 12, 40, dtype=torch.uint8, device='cuda')
    assert PP.multipitch_counts(ones, ones, 0).tolist() == [[480, 480, 480]]
    assert PP.multipitch_counts(ones, torch.zeros_like(ones), 3).tolist() == [[0, 480, 0]]
    assert PP.multipitch_scores(torch.zeros_like(ones), torch.zeros_like(ones), 2) == {'precision': 0.0, 'recall': 0.0, 'f1-score': 0.0}
    assert PP.activations_to_binary(torch.zeros(0, 12, 40, device='cuda')).shape == (0, 12, 40)
    with pytest.raises(TimbreTrapB200Error):
        PP.filter_non_peaks(torch.zeros(2, 3, 4))
    with pytest.raises(ValueError):
        PP.multipitch_counts(ones, ones[:, :6], 1)


def test_transcribe_then_score_end_to_end():
    """transcribe -> peak picking + threshold -> precision / recall / f1 without leaving the device (evaluate.py:98-116)."""
    from oracle import postproc_ref as R
    from tests.helpers import tonal_clip_with_pitches
    from timbre_trap_b200.framework import TimbreTrap
    from timbre_trap_b200.framework import postprocess as PP
    torch.manual_seed(0)
    model = TimbreTrap(8000, 6, 12, 0.5, 32, 1).cuda().eval()
    audio, _ = tonal_clip_with_pitches(3 * model.sliCQ.block_length, 8000, seed=3)
    act = model.transcribe(audio.cuda())
    t = float(act.quantile(0.97))
    est = PP.activations_to_binary(act, t, peaks_only=True)
    ref = PP.activations_to_binary(act, t, peaks_only=False)                 # a denser map of the same activations as "reference"
    sc = PP.multipitch_scores(est, ref, 1)
    e, r = est.cpu().numpy(), ref.cpu().numpy()
    tp, ne, nr = (sum(R.multipitch_counts(e[b], r[b], 1)[k] for b in range(e.shape[0])) for k in range(3))
    assert (sc['precision'], sc['recall'], sc['f1-score']) == R.prf(tp, ne, nr)
    assert sc['precision'] == 1.0 and 0.0 < sc['recall'] <= 1.0                # every picked peak is an active bin of the dense map
