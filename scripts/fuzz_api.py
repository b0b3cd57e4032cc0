"""One-off robustness sweep on a GPU: transcribe / reconstruct / forward over small geometries, clip lengths around the block length and
batch sizes, against the oracle.  Prints one line per case; exits non-zero if any case fails."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import model_ref as R
from tests.helpers import rel_err, tonal_clip
from timbre_trap_b200.framework import TimbreTrap

CASES = [
    (dict(sample_rate=8000, n_octaves=5, bins_per_octave=12, secs_per_block=0.064), None, 1, False),
    (dict(sample_rate=8000, n_octaves=5, bins_per_octave=12, secs_per_block=0.128), 8, 1, True),
    (dict(sample_rate=16000, n_octaves=6, bins_per_octave=12, secs_per_block=0.1), 24, 2, False),
    (dict(sample_rate=8000, n_octaves=6, bins_per_octave=12, secs_per_block=0.5), 3, 1, True),
    (dict(sample_rate=22050, n_octaves=9, bins_per_octave=60, secs_per_block=3), 128, 2, False),
]
bad = 0
for cfg, latent, cx, skip in CASES:
    try:
        model = TimbreTrap(cfg['sample_rate'], cfg['n_octaves'], cfg['bins_per_octave'], cfg['secs_per_block'], latent_size=latent, model_complexity=cx,
                           skip_connections=skip)
        sd = R.init_state_dict(model.sliCQ.n_bins, latent, cx, seed=2)
        if skip:
            sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
        model.load_state_dict(sd)
        model = model.cuda().eval()
        c = R.CQTRef(cfg['n_octaves'], cfg['bins_per_octave'], cfg['sample_rate'], cfg['secs_per_block'])
    except Exception:
        bad += 1
        print('CONSTRUCT FAIL', cfg, latent, cx, skip)
        traceback.print_exc()
        continue
    L = c.block_length
    big = cfg['secs_per_block'] >= 3
    lengths = [L] if big else [1, L // 3, L - 1, L, L + 1, 2 * L, int(2.5 * L), 5 * L + 7]
    for n in lengths:
        for batch in ((1,) if big else (1, 3)):
            tag = f"sr={cfg['sample_rate']} L={L} M={c.max_window_length} F={c.n_bins} latent={latent} cx={cx} skip={skip} n={n} B={batch}"
            try:
                audio = tonal_clip(max(n, 8), cfg['sample_rate'], seed=n % 97, n_batch=batch)[..., :n]
                act = model.transcribe(audio.cuda())
                want = R.transcribe_ref(audio, sd, c)
                ok = act.shape == want.shape and float((act.cpu() - want).abs().max()) <= 1e-2
                wav = model.reconstruct(audio.cuda())
                ok = ok and wav.shape[-1] == c.pad_to_block_length(audio).size(-1) and bool(torch.isfinite(wav).all())
                whole = c.pad_to_block_length(audio)
                rec = model(whole.cuda())[0]
                wr = R.forward_ref(whole, sd, c)[0]
                emax, el2 = rel_err(rec.cpu().numpy(), wr.numpy())
                ok = ok and el2 <= 1.5e-2 and emax <= 3e-2
                print('ok  ' if ok else 'FAIL', tag, 'act err', float((act.cpu() - want).abs().max()), 'rec', emax, el2)
                bad += 0 if ok else 1
            except Exception as e:
                bad += 1
                print('EXC ', tag, repr(e)[:300])
# ---- the variants over other geometries / latent sizes / clip lengths --------------------------------------------------------------
from timbre_trap_b200 import framework as FW

VARIANTS = dict(film=FW.TimbreTrapFiLM, mag=FW.TimbreTrapMag, magdb=FW.TimbreTrapMagDB)
for tag, cls in VARIANTS.items():
    for cfg, latent, cx, skip in ((dict(sample_rate=8000, n_octaves=5, bins_per_octave=12, secs_per_block=0.128), 8, 1, False),
                                  (dict(sample_rate=16000, n_octaves=4, bins_per_octave=24, secs_per_block=0.3), 40, 2, True),
                                  (dict(sample_rate=8000, n_octaves=6, bins_per_octave=12, secs_per_block=0.5), 17, 1, True)):
        try:
            model = cls(cfg['sample_rate'], cfg['n_octaves'], cfg['bins_per_octave'], cfg['secs_per_block'], latent_size=latent, model_complexity=cx,
                        skip_connections=skip)
            c = R.CQTRef(cfg['n_octaves'], cfg['bins_per_octave'], cfg['sample_rate'], cfg['secs_per_block'])
            sd = R.init_state_dict(c.n_bins, latent, cx, seed=5, variant=tag)
            if skip:
                sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
            model.load_state_dict(sd)
            model = model.cuda().eval()
        except Exception:
            bad += 1
            print('CONSTRUCT FAIL', tag, cfg, latent, cx, skip)
            traceback.print_exc()
            continue
        L = c.block_length
        for n in (L // 2, L, int(2.5 * L)):
            tagline = f"{tag} L={L} M={c.max_window_length} F={c.n_bins} latent={latent} cx={cx} skip={skip} n={n}"
            try:
                audio = tonal_clip(n, cfg['sample_rate'], seed=n % 89, n_batch=2)
                whole = c.pad_to_block_length(audio)
                want = R.forward_variant_ref(tag, whole, sd, c, consistency=True)
                got = model(whole.cuda(), consistency=True)
                errs = []
                ok = True
                for a, b in zip((got[0], got[2], got[3], got[4]), (want[0], want[2], want[3], want[4])):
                    emax, el2 = rel_err(a.cpu().numpy(), b.detach().numpy())
                    errs.append(round(el2, 4))
                    ok = ok and a.shape == b.shape and el2 <= 1.5e-2 and emax <= 3e-2
                ch = model.chunked_inference(audio.cuda(), True)
                wc = R.chunked_inference_variant_ref(tag, audio, sd, c, True)
                emax, el2 = rel_err(ch.cpu().numpy(), wc.detach().numpy())
                ok = ok and ch.shape == wc.shape and el2 <= 1.5e-2 and emax <= 3e-2
                act = model.transcribe(audio.cuda())
                # the magnitude variants inherit the two-channel chunk buffer (modules.py:244): squeeze(-3) in their to_activations is a
                # no-op on it, so transcribe() returns (B, 2, F, T) there - as the reference does
                want_shape = (2, c.n_bins, wc.shape[-1]) if tag == 'film' else (2, 2, c.n_bins, wc.shape[-1])
                ok = ok and tuple(act.shape) == want_shape and bool(torch.isfinite(act).all())
                if tuple(act.shape) != want_shape:
                    print('     transcribe shape', tuple(act.shape), 'expected', want_shape)
                print('ok  ' if ok else 'FAIL', tagline, 'forward l2', errs, 'chunked', round(emax, 4), round(el2, 4))
                bad += 0 if ok else 1
            except Exception as e:
                bad += 1
                print('EXC ', tagline, repr(e)[:300])

torch.cuda.synchronize()
print('failures:', bad)
sys.exit(1 if bad else 0)
