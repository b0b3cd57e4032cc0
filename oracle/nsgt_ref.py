"""
CPU oracle for the NSGT / sliCQ transform that `timbre_trap.framework.CQT` inherits.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may.  The product path is the CUDA extension and fails loudly without it.

PARITY UNPINNED.  The arithmetic lives in the third-party package `cqt_pytorch`
(PyPI `cqt-pytorch`, GitHub archinetai/cqt-pytorch), which the reference lists
*unpinned* at /root/reference/requirements.txt:15 and subclasses at
/root/reference/timbre_trap/framework/cqtwrapper.py:2,10,31-35.  That package is not
vendored in the reference, not installed in this image, and there is no network, and the
reference ships no tests or golden vectors for it.  This file therefore restates the
*published algorithm* (non-stationary Gabor frames with a constant-Q frequency-domain
Hann filter bank: Velasco, Holighaus, Doerfler, Grill 2011; Holighaus et al. 2013) in the
shape the package is documented to have, anchored on the reference's own call sites:

  * ctor kwargs `num_octaves, num_bins_per_octave, sample_rate, block_length,
    power_of_2_length`                                   (cqtwrapper.py:31-35)
  * attributes `block_length`, `max_window_length`       (cqtwrapper.py:40,231,271;
                                                          modules.py:230,237)
  * `encode(audio (B,1,T)) -> complex (B,1,F,T')`        (cqtwrapper.py:67,91)
  * `decode(complex (B,1,F,T')) -> real (B,1,T)`         (cqtwrapper.py:204-207)
  * exactly n_octaves*bins_per_octave bins, `max_window_length` frames per block,
    blocks concatenated along time                       (cqtwrapper.py:43,271)

Every choice that cannot be derived from the reference is a named constant below (U1..U8
in SURVEY.md Appendix C).  The CUDA kernels are table-driven from `NSGTTables`, so
switching to the real upstream package - once it can be imported - is a table change,
not a kernel change.

Arithmetic per block of L = block_length samples (F bins, M = max_window_length):

  analysis   X = fft_L(x)
             c_k[n] = (1/M) * sum_{m<M} X[(pos_k - M/2 + m) mod L] * W_k[m] * e^{+2 pi i m n / M}
  synthesis  Y[j]  = sum_{k,m : idx_k[m]=j} fft_M(c_k)[m] * Wd_k[m]
             y     = real(ifft_L(Y))
  dual       Wd_k[m] = W_k[m] / sum_{k',m' : idx_k'[m']=idx_k[m]} W_k'[m']^2
"""

import math

import numpy as np

__all__ = ['NSGTTables', 'NSGTOracle', 'make_tables']

# --- open choices -------------------------------------
U1_INCLUDE_NYQUIST_BAND_IN_MAX = True    # M = max over the F bins AND the Nyquist band
U3_ROUND_HALF_EVEN = True                # torch.round semantics for lengths / positions
U4_HANN_PERIODIC = True                  # torch.hann_window default (periodic=True)
U5_PAD_LEFT_FLOOR = True                 # left padding = floor(M/2 - len/2)
U6_CROP_ORIGIN_HALF_M = True             # crop starts at pos - M//2 (no fftshift)
U7_IFFT_SCALED = True                    # analysis ifft carries the 1/M factor
U8_DUAL_IS_FRAME_DIAGONAL = True         # Wd = W / sum W^2 gathered at the crop indices


def _round(x):
    """Round an ndarray the way torch.round does (half to even) or half away from zero."""
    if U3_ROUND_HALF_EVEN:
        return np.rint(x)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def _hann(n):
    """Hann window of n taps (periodic unless U4 says otherwise); hann(1) = [1]."""
    n = int(n)
    if n <= 0:
        return np.zeros(0)
    if n == 1:
        return np.ones(1)
    denom = n if U4_HANN_PERIODIC else n - 1
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / denom)


class NSGTTables:
    """
    Everything the transform needs, computed once on the host in float64.

    Dense (upstream-shaped) tables
      idx      (F, M) int64    spectrum index of every crop tap       ~ `windows_range_indices`
      win      (F, M) float64  centred zero-padded Hann               ~ `windows`
      win_inv  (F, M) float64  dual window                            ~ `windows_inverse`

    Packed (kernel-shaped) tables - only the non-zero taps of each bin
      length[k]   number of non-zero taps of bin k
      first[k]    m-index inside the M-crop of the first non-zero tap (the left padding)
      start[k]    spectrum index of that tap  (= idx[k, first[k]])
      offset[k]   prefix sum of length (start of bin k in the packed arrays)
      win_packed / dual_packed   concatenated non-zero taps (sum(length) values each)
    """

    def __init__(self, num_octaves, num_bins_per_octave, sample_rate, block_length, power_of_2_length=False):
        F = int(num_octaves) * int(num_bins_per_octave)
        L = int(block_length)
        sr = float(sample_rate)

        f_nyq = sr / 2.0
        f_min = f_nyq / (2.0 ** num_octaves)
        freqs = f_min * 2.0 ** (np.arange(F, dtype=np.float64) / num_bins_per_octave)
        # centre frequencies: F bins, then the Nyquist band (mirror image is never analysed)
        freqs_all = np.concatenate([freqs, [f_nyq]])
        # constant-Q bandwidths, Omega_k = f_k * (2^(1/b) - 2^(-1/b))
        q_spread = 2.0 ** (1.0 / num_bins_per_octave) - 2.0 ** (-1.0 / num_bins_per_octave)
        lengths_all = _round(freqs_all * q_spread * L / sr).astype(np.int64)
        lengths_all = np.maximum(lengths_all, 1)
        positions = _round(freqs * L / sr).astype(np.int64)

        m_max = int(lengths_all.max() if U1_INCLUDE_NYQUIST_BAND_IN_MAX else lengths_all[:F].max())
        if power_of_2_length:
            m_max = 1 << int(math.ceil(math.log2(m_max)))
        M = m_max

        length = lengths_all[:F].copy()
        if U5_PAD_LEFT_FLOOR:
            first = np.floor(M / 2.0 - length / 2.0).astype(np.int64)
        else:
            first = np.ceil(M / 2.0 - length / 2.0).astype(np.int64)
        origin = positions - (M // 2 if U6_CROP_ORIGIN_HALF_M else 0)

        idx = (origin[:, None] + np.arange(M, dtype=np.int64)[None, :]) % L
        win = np.zeros((F, M), dtype=np.float64)
        for k in range(F):
            win[k, first[k]:first[k] + length[k]] = _hann(length[k])

        # frame-operator diagonal on the spectrum grid, then gathered back at the taps
        diag = np.zeros(L, dtype=np.float64)
        np.add.at(diag, idx.reshape(-1), (win ** 2).reshape(-1))
        gathered = diag[idx]
        if U8_DUAL_IS_FRAME_DIAGONAL:
            win_inv = np.where(gathered > 0, win / np.where(gathered > 0, gathered, 1.0), 0.0)
        else:
            win_inv = win.copy()

        self.n_bins = F
        self.block_length = L
        self.max_window_length = M
        self.sample_rate = sr
        self.frequencies = freqs
        self.positions = positions
        self.idx = idx
        self.win = win
        self.win_inv = win_inv
        self.frame_diagonal = diag

        # packed tables
        self.length = length.astype(np.int32)
        self.first = first.astype(np.int32)
        self.start = ((origin + first) % L).astype(np.int32)
        self.offset = np.concatenate([[0], np.cumsum(length)]).astype(np.int32)
        self.win_packed = np.concatenate([win[k, first[k]:first[k] + length[k]] for k in range(F)])
        self.dual_packed = np.concatenate([win_inv[k, first[k]:first[k] + length[k]] for k in range(F)])

    @property
    def n_taps(self):
        return int(self.offset[-1])


def make_tables(n_octaves, bins_per_octave, sample_rate, secs_per_block):
    """Tables for the reference wrapper's ctor call (cqtwrapper.py:31-35)."""
    return NSGTTables(n_octaves, bins_per_octave, sample_rate, int(secs_per_block * sample_rate), True)


class NSGTOracle:
    """
    numpy restatement of `cqt_pytorch.CQT` (see the module docstring for what is and is
    not pinned).  `dtype` selects the arithmetic: complex128 is the truth the CUDA kernels
    are judged against; complex64 mimics the fp32 path the reference would take.
    """

    def __init__(self, num_octaves, num_bins_per_octave, sample_rate, block_length,
                 power_of_2_length=False, dtype=np.complex128):
        self.tables = NSGTTables(num_octaves, num_bins_per_octave, sample_rate, block_length, power_of_2_length)
        self.block_length = self.tables.block_length
        self.max_window_length = self.tables.max_window_length
        self.n_bins = self.tables.n_bins
        self.cdtype = np.dtype(dtype)
        self.rdtype = np.float64 if self.cdtype == np.complex128 else np.float32

    # -- analysis --------------------------------------------------------------------
    def encode(self, waveform):
        """(B, C, n*L) real -> (B, C, F, n*M) complex; blocks are concatenated on time."""
        t = self.tables
        L, M, F = t.block_length, t.max_window_length, t.n_bins
        x = np.asarray(waveform, dtype=self.rdtype)
        B, C, N = x.shape
        assert N % L == 0, 'audio must be a whole number of blocks (cqtwrapper.py:215-233 pads first)'
        n = N // L
        blocks = x.reshape(B, C, n, L)
        spectrum = np.fft.fft(blocks, axis=-1).astype(self.cdtype)           # (B,C,n,L)
        crops = spectrum[..., t.idx]                                          # (B,C,n,F,M)
        crops = crops * t.win.astype(self.rdtype)
        coeffs = np.fft.ifft(crops, axis=-1).astype(self.cdtype)
        if not U7_IFFT_SCALED:
            coeffs = coeffs * M
        # (B,C,n,F,M) -> (B,C,F,n*M)
        return np.ascontiguousarray(np.moveaxis(coeffs, 2, 3)).reshape(B, C, F, n * M)

    # -- synthesis -------------------------------------------------------------------
    def block_spectrum(self, transform):
        """The overlap-added one-sided spectrum Y of every block: (B, C, n, L) complex."""
        t = self.tables
        L, M, F = t.block_length, t.max_window_length, t.n_bins
        c = np.asarray(transform).astype(self.cdtype)
        B, C, F_, T = c.shape
        assert F_ == F and T % M == 0
        n = T // M
        c = np.moveaxis(c.reshape(B, C, F, n, M), 3, 2)                       # (B,C,n,F,M)
        taps = np.fft.fft(c, axis=-1).astype(self.cdtype)
        if not U7_IFFT_SCALED:
            taps = taps / M
        taps = taps * t.win_inv.astype(self.rdtype)
        Y = np.zeros((B * C * n, L), dtype=self.cdtype)
        flat = taps.reshape(B * C * n, F * M)
        np.add.at(Y, (np.arange(B * C * n)[:, None], t.idx.reshape(1, -1)), flat)
        return Y.reshape(B, C, n, L)

    def decode(self, transform):
        """(B, C, F, n*M) complex -> (B, C, n*L) real (NOT peak-normalised; the wrapper does that)."""
        Y = self.block_spectrum(transform)
        B, C, n, L = Y.shape
        y = np.fft.ifft(Y, axis=-1).real.astype(self.rdtype)
        return y.reshape(B, C, n * L)
