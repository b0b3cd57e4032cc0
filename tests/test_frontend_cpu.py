"""Pins the CPU oracle of the audio front end (oracle/frontend_ref.py) against tests/golden/frontend.npz - produced by the library
calls the reference makes (torchaudio) and by the reference's own PitchDataset.multi_pitch_to_activations (scripts/make_golden.py
frontend) - and, when torchaudio is importable, against torchaudio itself.  CPU only."""
import os

import numpy as np
import pytest

from oracle import frontend_ref as FR


def test_prepare_audio_oracle_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'frontend.npz'))
    n = 0
    for k in g.files:
        if k.startswith('audio_'):
            _, fs, ch = k.split('_')
            got = FR.prepare_audio_ref(g[k], int(fs), 22050)
            want = g[f'prepared_{fs}_{ch}']
            assert got.shape == want.shape
            assert np.abs(got - want).max() < 2e-6, (k, np.abs(got - want).max())
            assert abs(np.abs(got).max() - 1.0) < 1e-12
            n += 1
    assert n == 5
    assert FR.prepare_audio_ref(np.zeros((2, 100)), 44100, 22050).max() == 0.0        # silent clip: no divide (AudioDataset.py:75)


def test_resample_oracle_vs_torchaudio():
    torchaudio = pytest.importorskip('torchaudio')
    import torch
    rng = np.random.default_rng(1)
    for fs, sr in ((44100, 22050), (48000, 22050), (11025, 22050), (22050, 16000)):
        x = rng.standard_normal((1, 9001)).astype(np.float32)
        want = torchaudio.functional.resample(torch.from_numpy(x), fs, sr).numpy()
        got = FR.resample_ref(x, fs, sr)
        assert got.shape == want.shape and np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_multi_pitch_to_activations_oracle_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'frontend.npz'))
    mp = list(g['pitches_dense'])
    assert np.array_equal(FR.multi_pitch_to_activations_ref(mp, g['midi_freqs']), g['activations'])
    assert np.array_equal(FR.multi_pitch_to_activations_ref(mp, g['midi_freqs'], 0), g['activations_noblur'])
    assert np.array_equal(FR.multi_pitch_to_activations_ref([np.empty(0)] * 7, g['midi_freqs']), g['activations_empty'])
    a = g['activations']
    assert a.max() == 1.0 and a.min() == 0.0 and (a == 1).sum() >= 300


def test_sdr_oracle_definition():
    """sdr_ref restates torchmetrics' definition (absent from the image: unpinned); sanity anchors of the definition itself."""
    rng = np.random.default_rng(2)
    t = rng.standard_normal(6000)
    n = rng.standard_normal(6000)
    for snr_db in (0.0, 10.0, 25.0):
        p = t + n * np.linalg.norm(t) / np.linalg.norm(n) * 10 ** (-snr_db / 20)
        sdr = FR.sdr_ref(p, t, filter_length=64)
        assert abs(sdr - snr_db) < 1.0, (snr_db, sdr)               # white distortion: the 64-tap filter can explain ~1 % of it
    # invariant to gain and to any short FIR applied to the target inside the estimate
    h = np.array([0.5, -0.2, 0.1])
    p = np.convolve(t, h)[:6000] + 0.1 * n
    assert abs(FR.sdr_ref(p, t, 64) - FR.sdr_ref(3.0 * p, t, 64)) < 1e-9
    assert FR.sdr_ref(p, t, 64) > FR.sdr_ref(t + 0.3 * n, t, 64)
