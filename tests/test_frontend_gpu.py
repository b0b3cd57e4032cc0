"""GPU parity of the device-side front end (framework/frontend.py, csrc/frontend_kernels.cu) against the CPU oracle
(oracle/frontend_ref.py, pinned by tests/test_frontend_cpu.py) and the reference-generated golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prepare_audio_vs_golden_and_oracle(golden_dir):
    from oracle import frontend_ref as FR
    from timbre_trap_b200.framework import frontend as FE
    g = np.load(os.path.join(golden_dir, 'frontend.npz'))
    for k in g.files:
        if not k.startswith('audio_'):
            continue
        _, fs, ch = k.split('_')
        got = FE.prepare_audio(torch.from_numpy(g[k]).cuda(), int(fs), 22050).cpu().numpy()
        want = g[f'prepared_{fs}_{ch}']
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 3e-6, (k, np.abs(got - want).max())
        assert abs(np.abs(got).max() - 1.0) < 1e-6
    # ragged / edge cases: a length that is not a multiple of the phase count, one sample, silence, no normalisation
    rng = np.random.default_rng(3)
    for fs, n, ch in ((48000, 12345, 2), (44100, 1, 1), (16000, 37, 4), (96000, 100001, 2)):
        x = rng.standard_normal((ch, n)).astype(np.float32)
        got = FE.prepare_audio(torch.from_numpy(x).cuda(), fs, 22050).cpu().numpy()
        want = FR.prepare_audio_ref(x, fs, 22050)
        assert got.shape == want.shape and np.abs(got - want).max() < 3e-6, (fs, n)
        raw = FE.prepare_audio(torch.from_numpy(x).cuda(), fs, 22050, normalise=False).cpu().numpy()
        assert np.abs(raw - FR.resample_ref(x.astype(np.float64).mean(0, keepdims=True), fs, 22050)).max() < 3e-6 * max(1.0, np.abs(raw).max())
    z = FE.prepare_audio(torch.zeros(2, 5000).cuda(), 44100, 22050)
    assert z.shape == (1, 2500) and float(z.abs().max()) == 0.0


def test_prepare_audio_long_clip_properties():
    """A 10-minute stereo 48 kHz clip (size the oracle is not run at): linearity, peak exactly 1, a pure tone stays a pure tone."""
    from timbre_trap_b200.framework import frontend as FE
    fs, secs = 48000, 600
    t = torch.arange(fs * secs, device='cuda', dtype=torch.float64) / fs
    tone = torch.sin(2 * torch.pi * 1000.0 * t).float()
    x = torch.stack([tone, 0.5 * tone])
    y = FE.prepare_audio(x, fs, 22050)
    assert y.shape == (1, 22050 * secs) and abs(float(y.abs().max()) - 1.0) < 1e-6
    want = torch.sin(2 * torch.pi * 1000.0 * torch.arange(22050 * secs, device='cuda', dtype=torch.float64) / 22050).float()
    mid = slice(1000, -1000)
    assert float((y[0, mid] / y[0, mid].abs().max() - want[mid]).abs().max()) < 2e-3          # far from the edges: the same tone
    a = FE.prepare_audio(x, fs, 22050, normalise=False)
    b = FE.prepare_audio(2.0 * x, fs, 22050, normalise=False)
    assert torch.equal(b, 2.0 * a)


def test_multi_pitch_to_activations_vs_golden(golden_dir):
    from timbre_trap_b200.framework import frontend as FE
    g = np.load(os.path.join(golden_dir, 'frontend.npz'))
    dense, freqs = g['pitches_dense'], g['midi_freqs']
    for blur, key in ((2.5, 'activations'), (0, 'activations_noblur')):
        got = FE.multi_pitch_to_activations(torch.from_numpy(dense).cuda(), freqs, blur).cpu().numpy()
        err = np.abs(got - g[key])
        assert got.shape == g[key].shape and err.max() < 1e-6, (key, err.max(), np.unravel_index(err.argmax(), err.shape),
                                                                 got.flat[err.argmax()], g[key].flat[err.argmax()])
        assert np.array_equal(got == 1.0, g[key] == 1.0)                       # the exact ones the transcription loss keys on (objectives.py:65)
    listed = FE.multi_pitch_to_activations([row[row != 0] for row in dense], freqs, device='cuda').cpu().numpy()
    assert np.abs(listed - g['activations']).max() < 1e-6
    empty = FE.multi_pitch_to_activations([np.empty(0)] * 7, freqs, device='cuda')
    assert empty.shape == (540, 7) and float(empty.abs().max()) == 0.0


def test_sdr_vs_oracle():
    from oracle import frontend_ref as FR
    from timbre_trap_b200.framework import frontend as FE
    rng = np.random.default_rng(5)
    t = rng.standard_normal((3, 9000)).astype(np.float32)
    p = (t + 0.3 * rng.standard_normal((3, 9000))).astype(np.float32)
    p[1] = np.convolve(t[1], [0.6, 0.3, -0.1])[:9000] + 0.05 * rng.standard_normal(9000).astype(np.float32)
    got = FE.signal_distortion_ratio(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda(), filter_length=128).cpu().numpy()
    want = np.array([FR.sdr_ref(p[i], t[i], 128) for i in range(3)])
    assert np.abs(got - want).max() < 1e-3, (got, want)
    # default filter length on a longer signal, against the FFT form of the same correlations
    n = 200000
    tt_ = torch.randn(n, device='cuda')
    pp = tt_ + 0.1 * torch.randn(n, device='cuda')
    sdr = float(FE.signal_distortion_ratio(pp[None], tt_[None]))
    assert abs(sdr - 20.0) < 0.5, sdr
