"""GPU parity of the forward half of the loss step (experiments/train.py:404-458) against the CPU oracle."""
import numpy as np
import pytest
import torch

from tests.helpers import tonal_clip

pytestmark = pytest.mark.gpu

SMALL = dict(sample_rate=8000, n_octaves=6, bins_per_octave=12, secs_per_block=0.5)


def test_step_losses_match_oracle():
    from oracle import model_ref as R
    from timbre_trap_b200.framework import TimbreTrap
    from timbre_trap_b200.framework.train import compute_step_losses
    model = TimbreTrap(latent_size=None, model_complexity=1, **SMALL)
    sd = R.init_state_dict(model.sliCQ.n_bins, None, 1, seed=3)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    c = R.CQTRef(SMALL['n_octaves'], SMALL['bins_per_octave'], SMALL['sample_rate'], SMALL['secs_per_block'])
    audio = tonal_clip(3 * c.block_length, SMALL['sample_rate'], seed=4, n_batch=3)
    rng = np.random.default_rng(2)
    T = 3 * c.max_window_length
    gt = torch.zeros((2, c.n_bins, T))                               # 2 of the 3 items carry ground truth (train.py:429)
    for b in range(2):
        for k in rng.integers(5, c.n_bins - 5, size=3):
            gt[b, k, :] = 1.0
            gt[b, k - 1, :] = gt[b, k + 1, :] = 0.6
    got = compute_step_losses(model, audio.cuda(), gt.cuda())

    coeffs = c(audio)
    rec, lat, trn, trn_rec, trn_scr = R.forward_ref(audio, sd, c, consistency=True)
    act = torch.tanh(c.to_magnitude(trn))
    want = dict(reconstruction=R.reconstruction_loss_ref(rec, coeffs), transcription=R.transcription_loss_ref(act[:2], gt, True))
    want['consistency_spectral'], want['consistency_score'] = R.consistency_loss_ref(trn_rec[:2], trn_scr[:2], trn[:2])
    want['total'] = sum(want.values())
    # bf16 conv stack: every loss is a sum of squared differences of quantities carrying ~1e-2 relative noise, so besides a
    # relative tolerance each loss gets an absolute floor of (noise level)^2 x the energy of the compared tensors - the
    # consistency terms of a random-init model sit at that floor (the two decodes of one latent are almost identical)
    energy = float(R.reconstruction_loss_ref(trn[:2], torch.zeros_like(trn[:2])))
    for k, v in want.items():
        np.testing.assert_allclose(float(got[k]), float(v), rtol=3e-2, atol=(1.5e-2) ** 2 * energy, err_msg=k)
