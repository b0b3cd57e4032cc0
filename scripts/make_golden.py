"""
Generate tests/golden/*.npz|json by running the REFERENCE's own classes (imported unmodified
from /root/reference) on seeded inputs.  Runs only in the build container - /root/reference
does not exist on the GPU box - so the vectors are committed together with this script.

Two dependencies of the reference are missing from the image and are stubbed in sys.modules
before the import:
  * `cqt_pytorch.CQT`  -> an nn.Module shell around oracle/nsgt_ref.NSGTOracle (this is the one
                          boundary whose parity is UNPINNED: the fixtures pin the wrapper, the
                          conv autoencoder, the chunk loop and the objectives, not the NSGT
                          arithmetic itself)
  * `librosa`          -> hz_to_midi / midi_to_hz (one scalar call, cqtwrapper.py:45)

usage:  python scripts/make_golden.py
"""

import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.nsgt_ref import NSGTOracle          # noqa: E402
from oracle.model_ref import init_state_dict    # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

SMALL = dict(sample_rate=8000, n_octaves=6, bins_per_octave=12, secs_per_block=0.5)
BASE = dict(sample_rate=22050, n_octaves=9, bins_per_octave=60, secs_per_block=3)


def install_stubs():
    class _CQT(torch.nn.Module):
        def __init__(self, num_octaves, num_bins_per_octave, sample_rate, block_length, power_of_2_length=False):
            super().__init__()
            self._o = NSGTOracle(num_octaves, num_bins_per_octave, sample_rate, block_length, power_of_2_length)
            self.block_length = self._o.block_length
            self.max_window_length = self._o.max_window_length

        def encode(self, waveform):
            return torch.from_numpy(self._o.encode(waveform.detach().cpu().numpy()).astype(np.complex64))

        def decode(self, transform):
            return torch.from_numpy(self._o.decode(transform.detach().cpu().numpy()).astype(np.float32))

    cqt_pytorch = types.ModuleType('cqt_pytorch')
    cqt_pytorch.CQT = _CQT
    librosa = types.ModuleType('librosa')
    librosa.hz_to_midi = lambda f: 12.0 * (np.log2(np.asanyarray(f)) - np.log2(440.0)) + 69.0
    librosa.midi_to_hz = lambda m: 440.0 * (2.0 ** ((np.asanyarray(m) - 69.0) / 12.0))
    sys.modules['cqt_pytorch'] = cqt_pytorch
    sys.modules['librosa'] = librosa
    sys.path.insert(0, '/root/reference')


def tonal_clip(n_samples, sample_rate, seed, n_batch=1):
    """Harmonic tones + -30 dB noise, peak-normalised."""
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


def sub(x, fs, ts):
    """Strided sub-sample over the last two axes (keeps fixtures small)."""
    return x[..., ::fs, ::ts].contiguous().numpy()


def main():
    install_stubs()
    from timbre_trap.framework import CQT, TimbreTrap                                       # the reference
    from timbre_trap.framework import compute_reconstruction_loss, compute_transcription_loss, compute_consistency_loss
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---- 1. geometry -----------------------------------------------------------------
    geo = {}
    for name, cfg in (('small', SMALL), ('base', BASE)):
        c = CQT(cfg['n_octaves'], cfg['bins_per_octave'], cfg['sample_rate'], cfg['secs_per_block'])
        probes = [0, 1, c.block_length - 1, c.block_length, c.block_length + 1, 3 * c.block_length, 1234567]
        geo[name] = dict(cfg=cfg, block_length=c.block_length, max_window_length=c.max_window_length,
                         hop_length=c.hop_length, n_bins=c.n_bins,
                         midi_first=float(c.midi_freqs[0]), midi_last=float(c.midi_freqs[-1]),
                         midi_freqs_head=[float(v) for v in c.get_midi_freqs()[:5]],
                         expected_frames={str(p): c.get_expected_frames(p) for p in probes},
                         expected_samples={str(t): c.get_expected_samples(t) for t in (-1.0, 0.0, 0.37, 2.5)},
                         times_head=[float(v) for v in c.get_times(6)],
                         padded_len={str(p): int(c.pad_to_block_length(torch.zeros(1, 1, p)).size(-1)) for p in probes[1:]})
    with open(os.path.join(GOLDEN, 'geometry.json'), 'w') as f:
        json.dump(geo, f, indent=1)

    # ---- 2. CQT wrapper on the small config ---------------------------------------------
    c = CQT(SMALL['n_octaves'], SMALL['bins_per_octave'], SMALL['sample_rate'], SMALL['secs_per_block'])
    audio = tonal_clip(2 * c.block_length, SMALL['sample_rate'], seed=11, n_batch=2)
    coeffs = c(audio)
    mag = c.to_magnitude(coeffs)
    np.savez_compressed(os.path.join(GOLDEN, 'wrapper_small.npz'),
                        audio=audio.numpy(), coeffs_sub=sub(coeffs, 2, 3),
                        coeffs_strides=np.array(coeffs.stride()),
                        complex_ri_sub=torch.view_as_real(c.to_complex(coeffs))[:, ::2, ::3].contiguous().numpy(),
                        magnitude_sub=sub(mag, 2, 3), decibels_sub=sub(c.to_decibels(mag), 2, 3),
                        decibels_raw_sub=sub(c.to_decibels(mag, rescale=False), 2, 3),
                        decoded=c.decode(coeffs).numpy(),
                        decoded_from_complex=c.decode(c.to_complex(coeffs).unsqueeze(-3)).numpy(),
                        decoded_zero=c.decode(torch.zeros_like(coeffs)).numpy())

    # ---- 3. model on the small config (complexity 1 and 2, with/without skip) ---------
    for tag, complexity, latent, skip in (('c1', 1, None, False), ('c2skip', 2, 24, True)):
        model = TimbreTrap(SMALL['sample_rate'], SMALL['n_octaves'], SMALL['bins_per_octave'], SMALL['secs_per_block'],
                           latent_size=latent, model_complexity=complexity, skip_connections=skip).eval()
        sd = init_state_dict(model.sliCQ.n_bins, latent, complexity, seed=3)
        if skip:
            sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
        missing = model.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys and all(k.startswith('sliCQ.') for k in missing.missing_keys), missing
        assert len([k for k in model.state_dict() if not k.startswith('sliCQ.')]) == len(sd)
        audio = tonal_clip(int(1.3 * model.sliCQ.block_length), SMALL['sample_rate'], seed=5, n_batch=2)
        whole = model.sliCQ.pad_to_block_length(audio)
        lat, emb, _ = model.encode(whole)
        rec, lat2, trn, trn_rec, trn_scr, _ = model(whole, consistency=True)
        np.savez_compressed(os.path.join(GOLDEN, f'model_small_{tag}.npz'),
                            audio=audio.numpy(), latents=lat.numpy(),
                            emb_norms=np.array([float(e.norm()) for e in emb]),
                            emb_shapes=np.array([list(e.shape) for e in emb]),
                            reconstruction_sub=sub(rec, 2, 3), transcription_sub=sub(trn, 2, 3),
                            transcription_rec_sub=sub(trn_rec, 2, 3), transcription_scr_sub=sub(trn_scr, 2, 3),
                            inference_trn_sub=sub(model.inference(audio, True), 2, 3),
                            chunked_rec_sub=sub(model.chunked_inference(audio, False), 2, 3),
                            transcribe_sub=sub(model.transcribe(audio), 2, 3),
                            reconstruct=model.reconstruct(audio).numpy(),
                            activations_sub=sub(model.to_activations(trn), 2, 3))

    # ---- 4. objectives -----------------------------------------------------------------
    rng = np.random.default_rng(21)
    a = torch.from_numpy(rng.standard_normal((3, 2, 20, 17)).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal((3, 2, 20, 17)).astype(np.float32))
    d = torch.from_numpy(rng.standard_normal((3, 2, 20, 17)).astype(np.float32))
    est = torch.from_numpy(rng.uniform(0, 1, (3, 20, 17)).astype(np.float32))
    tgt = torch.from_numpy(rng.uniform(0, 1, (3, 20, 17)).astype(np.float32))
    tgt[rng.uniform(size=tgt.shape) < 0.1] = 1.0
    tgt[:, :, 3] = 0.0          # a frame with no positive mass
    tgt[0, :, 5] = 0.25         # a frame with mass but no exact ones
    cs, cc = compute_consistency_loss(a, b, d)
    np.savez_compressed(os.path.join(GOLDEN, 'objectives.npz'), a=a.numpy(), b=b.numpy(), d=d.numpy(),
                        est=est.numpy(), tgt=tgt.numpy(),
                        reconstruction=float(compute_reconstruction_loss(a, b)),
                        transcription_plain=float(compute_transcription_loss(est, tgt, False)),
                        transcription_weighted=float(compute_transcription_loss(est, tgt, True)),
                        consistency_spectral=float(cs), consistency_score=float(cc))

    # ---- 5. base config (BASELINE.json configs[0]) - sub-sampled ------------------------
    model = TimbreTrap(BASE['sample_rate'], BASE['n_octaves'], BASE['bins_per_octave'], BASE['secs_per_block'],
                       latent_size=128, model_complexity=2).eval()
    sd = init_state_dict(540, 128, 2, seed=0)
    model.load_state_dict(sd, strict=False)
    assert sum(v.numel() for v in sd.values()) == 614490
    audio = tonal_clip(66150, 22050, seed=0)
    coeffs = model.sliCQ(audio)
    lat, emb, _ = model.encode(audio)
    rec, _, trn, _, _, _ = model(audio)
    act = model.transcribe(audio)
    wav = model.reconstruct(audio)
    np.savez_compressed(os.path.join(GOLDEN, 'model_base_sub.npz'),
                        coeffs_sub=sub(coeffs, 9, 31), coeffs_norm=float(coeffs.norm()), coeffs_absmax=float(coeffs.abs().max()),
                        latents_sub=sub(lat, 5, 31), latents_norm=float(lat.norm()),
                        emb_norms=np.array([float(e.norm()) for e in emb]),
                        reconstruction_sub=sub(rec, 9, 31), reconstruction_norm=float(rec.norm()),
                        transcription_sub=sub(trn, 9, 31), transcription_norm=float(trn.norm()),
                        transcribe_sub=sub(act, 9, 31), transcribe_norm=float(act.norm()), transcribe_max=float(act.max()),
                        reconstruct_sub=wav[..., ::41].numpy(), reconstruct_norm=float(wav.norm()))
    make_variants()
    make_frontend()
    make_postproc()
    print('golden vectors written to', GOLDEN)
    for fn in sorted(os.listdir(GOLDEN)):
        print(f'  {fn:28s} {os.path.getsize(os.path.join(GOLDEN, fn)) / 1024:8.1f} KiB')


def make_variants():
    """tests/golden/model_small_{film,mag,magdb}.npz: the reference's TimbreTrapFiLM / TimbreTrapMag / TimbreTrapMagDB classes
    (modules.py:780-1075, imported unmodified) on the small configuration, complexity 2, skip connections on for the FiLM one."""
    install_stubs()
    from timbre_trap.framework import modules as M
    torch.set_grad_enabled(False)
    for tag, cls, skip in (('film', M.TimbreTrapFiLM, True), ('mag', M.TimbreTrapMag, False), ('magdb', M.TimbreTrapMagDB, False)):
        model = cls(SMALL['sample_rate'], SMALL['n_octaves'], SMALL['bins_per_octave'], SMALL['secs_per_block'],
                    latent_size=24, model_complexity=2, skip_connections=skip).eval()
        sd = init_state_dict(model.sliCQ.n_bins, 24, 2, seed=7, variant=tag)
        if skip:
            sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
        missing = model.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys and all(k.startswith('sliCQ.') for k in missing.missing_keys), missing
        audio = tonal_clip(int(1.3 * model.sliCQ.block_length), SMALL['sample_rate'], seed=5, n_batch=2)
        whole = model.sliCQ.pad_to_block_length(audio)
        lat, emb, _ = model.encode(whole)
        rec, _, trn, trn_rec, trn_scr, _ = model(whole, consistency=True)
        np.savez_compressed(os.path.join(GOLDEN, f'model_small_{tag}.npz'),
                            audio=audio.numpy(), latents=lat.numpy(), out_shape=np.array(rec.shape),
                            reconstruction_sub=sub(rec, 2, 3), transcription_sub=sub(trn, 2, 3),
                            transcription_rec_sub=sub(trn_rec, 2, 3), transcription_scr_sub=sub(trn_scr, 2, 3),
                            activations_sub=sub(model.to_activations(trn), 2, 3),
                            chunked_trn_sub=sub(model.chunked_inference(audio, True), 2, 3),
                            transcribe_shape=np.array(model.transcribe(audio).shape))


def reference_function(path, class_name, func_name, extra_globals):
    """A function of the reference, executed from its own source file WITHOUT importing the surrounding package (whose __init__
    pulls in dependencies missing from the image): the def is cut out of the file's AST at run time and exec'd unmodified."""
    import ast
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == func_name:
                    item.decorator_list = []
                    mod = ast.Module(body=[item], type_ignores=[])
                    ns = dict(extra_globals)
                    exec(compile(mod, path, 'exec'), ns)
                    return ns[func_name]
    raise KeyError(func_name)


def make_frontend():
    """tests/golden/frontend.npz: (a) AudioDataset.get_audio's arithmetic (datasets/AudioDataset.py:69-77) with the library calls the
    reference makes (torch.mean, torchaudio.functional.resample, infinity-norm divide) on seeded multi-channel clips at several
    source rates; (b) the reference's own PitchDataset.multi_pitch_to_activations (datasets/PitchDataset.py:233-307)."""
    import scipy
    import scipy.interpolate
    import scipy.ndimage
    import torchaudio
    import warnings
    install_stubs()
    import librosa
    rng = np.random.default_rng(17)
    out = {}
    for fs, ch in ((44100, 2), (48000, 1), (16000, 2), (22050, 3), (32000, 1)):
        t = np.arange(int(0.25 * fs)) / fs
        x = np.stack([np.sin(2 * np.pi * f0 * t + ph) for f0, ph in zip(rng.uniform(100, 3000, ch), rng.uniform(0, 6, ch))])
        x = (x + 0.05 * rng.standard_normal(x.shape)).astype(np.float32) * 0.3
        audio = torch.from_numpy(x)
        a = torch.mean(audio, dim=0, keepdim=True)                       # AudioDataset.py:71
        a = torchaudio.functional.resample(a, fs, 22050)                 # :73
        if a.abs().max():                                                # :75-77
            a /= a.abs().max()
        out[f'audio_{fs}_{ch}'] = x
        out[f'prepared_{fs}_{ch}'] = a.numpy()
    fn = reference_function('/root/reference/timbre_trap/datasets/PitchDataset.py', 'PitchDataset', 'multi_pitch_to_activations',
                            dict(np=np, scipy=scipy, filters=scipy.ndimage, librosa=librosa, warnings=warnings))
    midi_freqs = 16.765 + np.arange(540) / 5.0                            # the base CQT's bin grid (cqtwrapper.py:48)
    T = 200
    multi_pitch = []
    for i in range(T):
        n = int(rng.integers(0, 5)) if i % 17 else 0
        p = 440.0 * 2 ** ((rng.uniform(20, 120, n) - 69) / 12)
        if i % 29 == 3:
            p = np.concatenate([p, [0.0, 5.0, 30000.0]])                 # silence marker + out-of-range pitches
        if i == 50:
            # the outermost bins (a hair inside the range: exactly ON the boundary the reference's own `>= lb` test hangs on the last
            # bit of log2) and three neighbouring bins
            edge = midi_freqs[[0, 300, 301, 302, 539]] + np.array([1e-6, 0, 0, 0, -1e-6])
            p = np.concatenate([p, 440.0 * 2 ** ((edge - 69) / 12)])
        multi_pitch.append(p)
    P = max(len(p) for p in multi_pitch)
    dense = np.zeros((T, P))
    for i, p in enumerate(multi_pitch):
        dense[i, :len(p)] = p
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out['activations'] = fn(multi_pitch, midi_freqs)
        out['activations_noblur'] = fn(multi_pitch, midi_freqs, 0)
        out['activations_empty'] = fn([np.empty(0)] * 7, midi_freqs)
    out['pitches_dense'] = dense
    out['midi_freqs'] = midi_freqs
    np.savez_compressed(os.path.join(GOLDEN, 'frontend.npz'), **out)


def make_postproc():
    """tests/golden/postproc.npz: the reference's own filter_non_peaks / threshold (timbre_trap/utils/processing.py:66-124, numpy +
    scipy, imported unmodified) on a seeded activation map with plateaus, ties, edge peaks and exact-threshold values."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_processing', '/root/reference/timbre_trap/utils/processing.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(5)
    act = rng.random((2, 48, 37)).astype(np.float32)
    act = np.round(act * 16) / 16                       # many ties and plateaus
    act[0, 0, :5] = 1.0                                 # peaks at the lower edge
    act[0, -1, 5:9] = 1.0                               # ... and the upper edge
    act[1, 10:13, 3] = 0.75                             # a plateau: no strict peak
    act[1, 20, :] = 0.5                                 # exactly at the threshold
    peaks = mod.filter_non_peaks(act)
    binary = mod.threshold(act, 0.5)
    peaks_binary = mod.threshold(mod.filter_non_peaks(act), 0.5)
    np.savez_compressed(os.path.join(GOLDEN, 'postproc.npz'), activations=act, filter_non_peaks=peaks.astype(np.float32),
                        threshold=binary.astype(np.uint8), peaks_threshold=peaks_binary.astype(np.uint8))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'postproc':
        make_postproc()
    elif len(sys.argv) > 1 and sys.argv[1] == 'variants':
        make_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == 'frontend':
        make_frontend()
    else:
        main()
