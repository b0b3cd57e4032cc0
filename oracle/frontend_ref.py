"""
CPU oracle of the audio front end and the SDR metric.  TEST INFRASTRUCTURE ONLY (same rule as oracle/nsgt_ref.py).

  prepare_audio_ref            AudioDataset.get_audio after torchaudio.load (timbre_trap/datasets/AudioDataset.py:69-77).  The
                               resampler is the library call the reference makes (torchaudio.functional.resample); this file restates
                               its published algorithm in numpy float64 and tests/test_frontend_cpu.py PINS it against the installed
                               torchaudio and against tests/golden/frontend.npz (generated with torchaudio by scripts/make_golden.py).
  multi_pitch_to_activations_ref   PitchDataset.multi_pitch_to_activations (datasets/PitchDataset.py:233-307), pinned by the same
                               golden file, which holds the output of the reference's own static method.
  sdr_ref                      torchmetrics' signal_distortion_ratio (experiments/evaluate.py:51,122-127).  torchmetrics is absent
                               from the image: PARITY UNPINNED for this one function; it restates the published definition
                               (fast_bss_eval / Scheibler 2021: SDR from the coherence of the optimal 512-tap distortion filter).
"""

import math

import numpy as np


def resample_ref(x, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio.functional.resample(x (..., N)) with its defaults (sinc_interp_hann).  The filter table is built in FLOAT32 as
    torchaudio builds it for float32 audio (its large sinc arguments make a float64 table differ at the 1e-5 level); the
    convolution runs in float64."""
    x = np.asarray(x, dtype=np.float64)
    if int(orig_freq) == int(new_freq):
        return x.copy()
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    f32 = np.float32
    idx = np.arange(-width, width + orig, dtype=f32)[None, :] / f32(orig)
    t = np.clip((np.arange(0, -new, -1, dtype=f32)[:, None] / f32(new) + idx) * f32(base), -lowpass_filter_width, lowpass_filter_width).astype(f32)
    window = np.cos(t * f32(math.pi) / f32(lowpass_filter_width) / f32(2)) ** 2
    t = (t * f32(math.pi)).astype(f32)
    with np.errstate(invalid='ignore', divide='ignore'):
        kern = (np.where(t == 0, f32(1.0), np.sin(t) / t) * (window * f32(base / orig))).astype(np.float64)           # (new, K)
    n = x.shape[-1]
    padded = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(width, width + orig)])
    K = kern.shape[1]
    n_frames = (padded.shape[-1] - K) // orig + 1
    frames = np.lib.stride_tricks.sliding_window_view(padded, K, axis=-1)[..., ::orig, :][..., :n_frames, :]
    out = np.einsum('...fk,jk->...fj', frames, kern).reshape(x.shape[:-1] + (-1,))
    return out[..., : int(math.ceil(new * n / orig))]


def prepare_audio_ref(audio, fs, sample_rate):
    """(C, N) -> (1, N'): mean over channels, resample, infinity-norm normalise when the peak is non-zero (AudioDataset.py:71-77)."""
    mono = np.asarray(audio, dtype=np.float64).mean(axis=0, keepdims=True)
    out = resample_ref(mono, fs, sample_rate)
    peak = np.abs(out).max() if out.size else 0.0
    return out / peak if peak else out


def multi_pitch_to_activations_ref(multi_pitch, midi_freqs, n_bins_blur_decay=2.5):
    """PitchDataset.py:233-307: nearest bin (ties to the lower bin, scipy interp1d kind='nearest'), Gaussian blur along frequency
    (sigma = 2 * decay / 5 bins, truncated at 4 sigma, zero boundary), renormalise by the smallest annotated cell, clip."""
    midi_freqs = np.asarray(midi_freqs, dtype=np.float64)
    F, T = len(midi_freqs), len(multi_pitch)
    act = np.zeros((F, T))
    mids = 0.5 * (midi_freqs[1:] + midi_freqs[:-1])
    cells = []
    for t, p in enumerate(multi_pitch):
        p = np.asarray(p, dtype=np.float64)
        p = p[p != 0]
        midi = 12.0 * (np.log2(p) - np.log2(440.0)) + 69.0
        midi = midi[(midi >= midi_freqs[0]) & (midi <= midi_freqs[-1])]
        for k in np.searchsorted(mids, midi, side='left'):
            act[k, t] = 1.0
            cells.append((k, t))
    if cells and n_bins_blur_decay:
        sigma = (2 * n_bins_blur_decay) / 5
        radius = int(4.0 * sigma + 0.5)
        x = np.arange(-radius, radius + 1)
        w = np.exp(-0.5 / sigma ** 2 * x ** 2)
        w /= w.sum()
        padded = np.pad(act, ((radius, radius), (0, 0)))
        act = sum(w[i] * padded[i: i + F] for i in range(2 * radius + 1))
        act = np.clip(act / min(act[k, t] for k, t in cells), 0.0, 1.0)
    return act


def sdr_ref(preds, target, filter_length=512, zero_mean=False):
    """torchmetrics.functional.signal_distortion_ratio with default arguments, dense numpy float64 (small inputs only)."""
    p = np.asarray(preds, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    if zero_mean:
        p, t = p - p.mean(-1, keepdims=True), t - t.mean(-1, keepdims=True)
    t = t / max(np.linalg.norm(t), 1e-6)
    p = p / max(np.linalg.norm(p), 1e-6)
    n = len(t)
    r0 = np.array([np.dot(t[: n - l], t[l:]) for l in range(filter_length)])
    b = np.array([np.dot(t[: n - l], p[l:]) for l in range(filter_length)])
    idx = np.abs(np.arange(filter_length)[:, None] - np.arange(filter_length)[None, :])
    sol = np.linalg.solve(r0[idx], b)
    coh = float(b @ sol)
    return 10.0 * np.log10(coh / (1 - coh))
