"""
Pins the CPU oracle (oracle/model_ref.py) against vectors produced by the reference's own
classes (scripts/make_golden.py).  CPU only.  The NSGT arithmetic itself is NOT pinned by
these vectors (the reference's `cqt_pytorch` dependency is un-vendored): both sides used
oracle/nsgt_ref.py for it - see the header there.
"""
import json
import os

import numpy as np
import torch

from oracle import model_ref as R

SMALL = dict(sample_rate=8000, n_octaves=6, bins_per_octave=12, secs_per_block=0.5)


def _close(a, b, rtol=1e-5, atol=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    atol = rtol * np.abs(b).max() if atol is None else atol
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= atol, (np.abs(a - b).max(), atol)


def _sub(x, fs=2, ts=3):
    return x[..., ::fs, ::ts]


def test_geometry(golden_dir):
    geo = json.load(open(os.path.join(golden_dir, 'geometry.json')))
    for name, g in geo.items():
        cfg = g['cfg']
        c = R.CQTRef(cfg['n_octaves'], cfg['bins_per_octave'], cfg['sample_rate'], cfg['secs_per_block'])
        assert c.block_length == g['block_length'] and c.max_window_length == g['max_window_length']
        assert c.n_bins == g['n_bins'] and c.hop_length == g['hop_length']
        assert abs(c.midi_freqs[0] - g['midi_first']) < 1e-9 and abs(c.midi_freqs[-1] - g['midi_last']) < 1e-9
        for p, v in g['expected_frames'].items():
            assert c.get_expected_frames(int(p)) == v
        for t, v in g['expected_samples'].items():
            assert c.get_expected_samples(float(t)) == v
        for p, v in g['padded_len'].items():
            assert c.pad_to_block_length(torch.zeros(1, 1, int(p))).size(-1) == v
        np.testing.assert_allclose(c.get_times(6), g['times_head'], rtol=1e-12)
    b = geo['base']
    assert (b['block_length'], b['max_window_length'], b['n_bins']) == (66150, 1024, 540)
    assert abs(b['midi_first'] - 16.7656) < 1e-3


def test_wrapper_small(golden_dir):
    g = np.load(os.path.join(golden_dir, 'wrapper_small.npz'))
    c = R.CQTRef(**{k: SMALL[k] for k in ('n_octaves', 'bins_per_octave', 'sample_rate', 'secs_per_block')})
    audio = torch.from_numpy(g['audio'])
    coeffs = c(audio)
    assert tuple(coeffs.stride()) == tuple(g['coeffs_strides'])      # NHWC memory of the (B,2,F,T) view
    _close(_sub(coeffs), g['coeffs_sub'])
    _close(torch.view_as_real(c.to_complex(coeffs))[:, ::2, ::3], g['complex_ri_sub'])
    mag = c.to_magnitude(coeffs)
    _close(_sub(mag), g['magnitude_sub'])
    _close(_sub(R.to_decibels_ref(mag)), g['decibels_sub'], atol=2e-6)
    _close(_sub(R.to_decibels_ref(mag, rescale=False)), g['decibels_raw_sub'], atol=2e-4)
    _close(c.decode(coeffs), g['decoded'], atol=2e-6)
    _close(c.decode(c.to_complex(coeffs).unsqueeze(-3)), g['decoded_from_complex'], atol=2e-6)
    assert np.all(g['decoded_zero'] == 0) and float(c.decode(torch.zeros_like(coeffs)).abs().max()) == 0.0


def _model_case(golden_dir, tag, complexity, latent, skip):
    g = np.load(os.path.join(golden_dir, f'model_small_{tag}.npz'))
    c = R.CQTRef(**{k: SMALL[k] for k in ('n_octaves', 'bins_per_octave', 'sample_rate', 'secs_per_block')})
    sd = R.init_state_dict(c.n_bins, latent, complexity, seed=3)
    if skip:
        sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
    audio = torch.from_numpy(g['audio'])
    whole = c.pad_to_block_length(audio)
    lat, emb = R.encoder_ref(c(whole), sd)
    _close(lat, g['latents'], rtol=2e-5)
    assert [list(e.shape) for e in emb] == g['emb_shapes'].tolist()
    np.testing.assert_allclose([float(e.norm()) for e in emb], g['emb_norms'], rtol=1e-5)
    rec, lat2, trn, trn_rec, trn_scr = R.forward_ref(whole, sd, c, consistency=True)
    _close(_sub(rec), g['reconstruction_sub'], rtol=2e-5)
    _close(_sub(trn), g['transcription_sub'], rtol=2e-5)
    _close(_sub(trn_rec), g['transcription_rec_sub'], rtol=2e-5)
    _close(_sub(trn_scr), g['transcription_scr_sub'], rtol=2e-5)
    _close(_sub(torch.tanh(c.to_magnitude(trn))), g['activations_sub'], rtol=2e-5)
    _close(_sub(R.inference_ref(whole, sd, c, True)), g['inference_trn_sub'], rtol=2e-5)
    _close(_sub(R.chunked_inference_ref(audio, sd, c, False)), g['chunked_rec_sub'], rtol=2e-5)
    _close(_sub(R.transcribe_ref(audio, sd, c)), g['transcribe_sub'], rtol=2e-5)
    # the synthesis dual window reaches ~5e3 (small config) where a single Hann tail covers the spectrum, so
    # thread-order noise of the fp32 convs (6e-8 on the coefficients) is amplified on this random-weight output
    _close(R.reconstruct_ref(audio, sd, c), g['reconstruct'], atol=2e-3)


def test_model_small_c1(golden_dir):
    _model_case(golden_dir, 'c1', 1, None, False)


def test_model_small_c2_skip(golden_dir):
    _model_case(golden_dir, 'c2skip', 2, 24, True)


def test_objectives(golden_dir):
    g = np.load(os.path.join(golden_dir, 'objectives.npz'))
    a, b, d = (torch.from_numpy(g[k]) for k in 'abd')
    est, tgt = torch.from_numpy(g['est']), torch.from_numpy(g['tgt'])
    np.testing.assert_allclose(float(R.reconstruction_loss_ref(a, b)), float(g['reconstruction']), rtol=1e-6)
    np.testing.assert_allclose(float(R.transcription_loss_ref(est, tgt, False)), float(g['transcription_plain']), rtol=1e-6)
    np.testing.assert_allclose(float(R.transcription_loss_ref(est, tgt, True)), float(g['transcription_weighted']), rtol=1e-6)
    cs, cc = R.consistency_loss_ref(a, b, d)
    np.testing.assert_allclose([float(cs), float(cc)], [float(g['consistency_spectral']), float(g['consistency_score'])], rtol=1e-6)


def test_postprocessing_oracle_vs_reference_functions(golden_dir):
    """oracle/postproc_ref.py against tests/golden/postproc.npz = outputs of the reference's own filter_non_peaks / threshold
    (timbre_trap/utils/processing.py:66-124); the matching counts against a brute-force maximum bipartite matching."""
    import itertools
    from oracle import postproc_ref as R
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    a = g['activations']
    assert np.array_equal(R.filter_non_peaks(a).astype(np.float32), g['filter_non_peaks'])
    assert np.array_equal(R.threshold(a, 0.5).astype(np.uint8), g['threshold'])
    assert np.array_equal(R.binary_map(a, 0.5, peaks_only=True), g['peaks_threshold'])
    assert np.array_equal(R.binary_map(a, 0.5), g['threshold'])
    masked = R.binary_map(a, 0.5, bin_lo=3, bin_hi=40)
    assert masked[:, :3].sum() == 0 and masked[:, 40:].sum() == 0 and np.array_equal(masked[:, 3:40], g['threshold'][:, 3:40])

    def brute(e, r, tol):                       # maximum matching by exhaustive search (tiny frames only)
        ei, ri = np.flatnonzero(e), np.flatnonzero(r)
        best = 0
        for k in range(min(len(ei), len(ri)), 0, -1):
            for es in itertools.combinations(ei, k):
                for rs in itertools.permutations(ri, k):
                    if all(abs(int(x) - int(y)) <= tol for x, y in zip(es, rs)):
                        return k
        return best
    rng = np.random.default_rng(11)
    for tol in (0, 1, 2):
        est = (rng.random((9, 40)) < 0.3).astype(np.uint8)
        ref = (rng.random((9, 40)) < 0.3).astype(np.uint8)
        tp, ne, nr = R.multipitch_counts(est, ref, tol)
        assert ne == est.sum() and nr == ref.sum()
        assert tp == sum(brute(est[:, t], ref[:, t], tol) for t in range(est.shape[1]))
    same = (rng.random((30, 50)) < 0.1).astype(np.uint8)
    assert R.prf(*R.multipitch_counts(same, same, 0))[:2] == (1.0, 1.0)
    assert R.prf(*R.multipitch_counts(same, np.zeros_like(same), 2)) == (0.0, 0.0, 0.0)


def test_model_variants_small(golden_dir):
    """TimbreTrapFiLM / TimbreTrapMag / TimbreTrapMagDB (modules.py:780-1075): the oracle's variant functions against the reference's
    own classes (scripts/make_golden.py variants), including the inherited chunk loop's two-channel broadcast quirk."""
    import pytest
    c = R.CQTRef(**{k: SMALL[k] for k in ('n_octaves', 'bins_per_octave', 'sample_rate', 'secs_per_block')})
    for tag in ('film', 'mag', 'magdb'):
        g = np.load(os.path.join(golden_dir, f'model_small_{tag}.npz'))
        sd = R.init_state_dict(c.n_bins, 24, 2, seed=7, variant=tag)
        if tag == 'film':
            sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
        audio = torch.from_numpy(g['audio'])
        whole = c.pad_to_block_length(audio)
        rec, lat, trn, trn_rec, trn_scr = R.forward_variant_ref(tag, whole, sd, c, consistency=True)
        assert tuple(rec.shape) == tuple(g['out_shape']) and rec.shape[1] == (2 if tag == 'film' else 1)
        _close(lat, g['latents'], rtol=2e-5)
        for name, t in (('reconstruction', rec), ('transcription', trn), ('transcription_rec', trn_rec), ('transcription_scr', trn_scr)):
            _close(_sub(t), g[name + '_sub'], rtol=2e-5)
        _close(_sub(R.activations_variant_ref(tag, trn)), g['activations_sub'], rtol=2e-5)
        ch = R.chunked_inference_variant_ref(tag, audio, sd, c, True)
        _close(_sub(ch), g['chunked_trn_sub'], rtol=2e-5)
        # transcribe() of the magnitude variants keeps the broadcast channel axis (reference quirk: squeeze(-3) of a 2-channel tensor)
        assert len(g['transcribe_shape']) == (3 if tag == 'film' else 4)
