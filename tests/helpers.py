import numpy as np
import torch


def tonal_clip(n_samples, sample_rate, seed, n_batch=1):
    """Same generator as scripts/make_golden.py (harmonic tones + -30 dB noise, peak-normalised)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / sample_rate
    out = []
    for _ in range(n_batch):
        x = np.zeros(n_samples)
        for midi in rng.integers(40, 90, size=4):
            f0 = 440.0 * 2 ** ((midi - 69) / 12)
            for h in range(1, 5):
                if f0 * h < 0.45 * sample_rate:
                    x += np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi)) / h
        x += 10 ** (-30 / 20) * rng.standard_normal(n_samples)
        out.append(x / np.abs(x).max())
    return torch.from_numpy(np.stack(out)[:, None, :].astype(np.float32))


def rel_err(a, b):
    """(max-abs error / max|b|, l2 error / ||b||) - the norm-relative bounds of SURVEY.md section 8c."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)), float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
