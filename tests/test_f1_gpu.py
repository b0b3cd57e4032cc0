"""
"Unchanged multi-pitch F1 on synthetic clips" (BASELINE.json north_star).  The reference's evaluation post-processing
(evaluate.py:105-116: peak picking along frequency + threshold + frame-wise pitch matching) is restated in tests/helpers.py
and applied to the activations of the CUDA path and of the CPU oracle.

Caveat ("F1 vacuity"): no trained checkpoint ships with the reference and with random-init weights no
activation reaches the reference threshold 0.5 - the maps are flat, noise-like fields, so local maxima and near-threshold cells
flip under ANY perturbation at the level of the stated bf16 tolerance (1e-2).  The test therefore checks what determines F1:
  (1) threshold decisions are sandwiched: every CUDA detection at t is an oracle detection at t - delta, and every oracle
      detection at t + delta is a CUDA detection (delta = the stated activation tolerance), so F1 can only differ through
      cells inside the tolerance band;
  (2) with the band excluded the two detection sets are identical, hence identical F1 against any ground truth;
  (3) the peak-picked F1 between the two sides is reported (informational on random weights).
"""
import numpy as np
import pytest
import torch

from tests.helpers import multipitch_prf, pick_peaks, tonal_clip_with_pitches

pytestmark = pytest.mark.gpu

DELTA = 1e-2


def test_f1_decisions_unchanged_within_tolerance():
    from oracle import model_ref as R
    from timbre_trap_b200.framework import TimbreTrap
    model = TimbreTrap(22050, 9, 60, 3, latent_size=128, model_complexity=2)
    sd = R.init_state_dict(540, 128, 2, seed=0)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    ref_cqt = R.CQTRef(9, 60, 22050, 3)
    for seed in (0, 5):
        audio, midis = tonal_clip_with_pitches(66150, 22050, seed)
        want = R.transcribe_ref(audio, sd, ref_cqt)[0].numpy()
        got = model.transcribe(audio.cuda())[0].cpu().numpy()
        assert np.abs(got - want).max() <= DELTA
        for q in (0.5, 0.9, 0.98):
            t = float(np.quantile(want, q))
            det_gpu = got >= t
            assert not (det_gpu & ~(want >= t - DELTA)).any()          # no detection the oracle would not make at t - delta
            assert not ((want >= t + DELTA) & ~det_gpu).any()          # no miss of a detection the oracle makes at t + delta
            band = np.abs(want - t) < DELTA
            assert np.array_equal(det_gpu[~band], (want >= t)[~band])
            gt = np.zeros_like(det_gpu)
            for m in midis:
                k = int(round((m - model.sliCQ.midi_freqs[0]) * 5))
                if 0 <= k < 540:
                    gt[k] = True
            f_gpu = multipitch_prf(det_gpu & ~band, gt, 2)[2]
            f_ref = multipitch_prf((want >= t) & ~band, gt, 2)[2]
            assert f_gpu == f_ref
        t = float(np.quantile(want, 0.9))
        f1 = multipitch_prf(pick_peaks(got, t), pick_peaks(want, t), tol_bins=2)[2]
        print(f'seed {seed}: peak-picked F1 (CUDA picks vs oracle picks, random-init weights, t = {t:.3f}): {f1:.3f}')
        assert f1 > 0.5
