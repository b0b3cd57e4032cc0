"""GPU parity of the three objectives against the reference-produced golden scalars and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_objectives_match_golden(golden_dir):
    from timbre_trap_b200.framework import compute_consistency_loss, compute_reconstruction_loss, compute_transcription_loss
    g = np.load(os.path.join(golden_dir, 'objectives.npz'))
    a, b, d = (torch.from_numpy(g[k]).cuda() for k in 'abd')
    est, tgt = torch.from_numpy(g['est']).cuda(), torch.from_numpy(g['tgt']).cuda()
    np.testing.assert_allclose(float(compute_reconstruction_loss(a, b)), float(g['reconstruction']), rtol=1e-6)
    np.testing.assert_allclose(float(compute_transcription_loss(est, tgt, False)), float(g['transcription_plain']), rtol=1e-6)
    np.testing.assert_allclose(float(compute_transcription_loss(est, tgt, True)), float(g['transcription_weighted']), rtol=2e-6)
    cs, cc = compute_consistency_loss(a, b, d)
    np.testing.assert_allclose([float(cs), float(cc)], [float(g['consistency_spectral']), float(g['consistency_score'])], rtol=1e-6)


def test_objectives_large_and_layouts():
    from oracle import model_ref as R
    from timbre_trap_b200.framework import compute_reconstruction_loss, compute_transcription_loss
    rng = np.random.default_rng(0)
    # channels-last views (what CQT.forward and the decoder return) and plain contiguous tensors, odd sizes
    a = torch.from_numpy(rng.standard_normal((3, 37, 129, 2)).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal((3, 37, 129, 2)).astype(np.float32))
    want = float(R.reconstruction_loss_ref(a.permute(0, 3, 1, 2).double(), b.permute(0, 3, 1, 2).double()))
    got_view = float(compute_reconstruction_loss(a.cuda().permute(0, 3, 1, 2), b.cuda().permute(0, 3, 1, 2)))
    got_mixed = float(compute_reconstruction_loss(a.cuda().permute(0, 3, 1, 2), b.permute(0, 3, 1, 2).contiguous().cuda()))
    np.testing.assert_allclose([got_view, got_mixed], [want, want], rtol=2e-6)
    est = torch.from_numpy(rng.uniform(0, 1, (4, 540, 300)).astype(np.float32))
    tgt = torch.from_numpy((rng.uniform(0, 1, (4, 540, 300)) ** 8).astype(np.float32))
    tgt[torch.from_numpy(rng.uniform(size=tgt.shape) < 0.01)] = 1.0
    tgt[:, :, 7] = 1.0     # every bin positive: scaling would be 0 -> 1 (objectives.py:67)
    tgt[:, :, 9] = 0.0
    for w in (False, True):
        want = float(R.transcription_loss_ref(est.double(), tgt.double(), w))
        np.testing.assert_allclose(float(compute_transcription_loss(est.cuda(), tgt.cuda(), w)), want, rtol=5e-6)
    # bit-reproducible
    x, y = a.cuda().permute(0, 3, 1, 2), b.cuda().permute(0, 3, 1, 2)
    assert float(compute_reconstruction_loss(x, y)) == float(compute_reconstruction_loss(x, y))
