"""Second robustness sweep (GPU): the loss step's gradients on other geometries against the oracle's autograd, forward_tiled against the
un-tiled forward on other geometries / tile sizes, block sharding with 1-7 emulated ranks on ragged clips.  Exits non-zero on failure."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import model_ref as R
from tests.helpers import tonal_clip
from timbre_trap_b200.framework import TimbreTrap
from timbre_trap_b200.framework.train import TrainStep

bad = 0


def build(cfg, latent, cx, skip, seed=1):
    model = TimbreTrap(cfg['sample_rate'], cfg['n_octaves'], cfg['bins_per_octave'], cfg['secs_per_block'], latent_size=latent, model_complexity=cx,
                       skip_connections=skip)
    sd = R.init_state_dict(model.sliCQ.n_bins, latent, cx, seed=seed)
    if skip:
        sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
    model.load_state_dict(sd)
    c = R.CQTRef(cfg['n_octaves'], cfg['bins_per_octave'], cfg['sample_rate'], cfg['secs_per_block'])
    return model.cuda(), sd, c


GEOS = [(dict(sample_rate=8000, n_octaves=5, bins_per_octave=12, secs_per_block=0.128), 8, 1, False),
        (dict(sample_rate=16000, n_octaves=4, bins_per_octave=24, secs_per_block=0.3), 40, 2, True),
        (dict(sample_rate=16000, n_octaves=6, bins_per_octave=12, secs_per_block=0.1), 24, 2, False)]

# ---- 1. loss-step gradients -----------------------------------------------------------------------------------------------------
for cfg, latent, cx, skip in GEOS:
    tag = f"train F-geometry sr={cfg['sample_rate']} oct={cfg['n_octaves']} bpo={cfg['bins_per_octave']} latent={latent} cx={cx} skip={skip}"
    try:
        model, sd, c = build(cfg, latent, cx, skip)
        audio = tonal_clip(3 * c.block_length, cfg['sample_rate'], seed=4, n_batch=2)
        rng = np.random.default_rng(2)
        gt = torch.zeros((2, c.n_bins, 3 * c.max_window_length))
        for b in range(2):
            for k in rng.integers(5, c.n_bins - 5, size=3):
                gt[b, k, :] = 1.0
                gt[b, k - 1, :] = gt[b, k + 1, :] = 0.6
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        coeffs = c(audio)
        rec, lat, trn, trn_rec, trn_scr = R.forward_ref(audio, sdg, c, consistency=True)
        act = torch.tanh(c.to_magnitude(trn))
        total = (R.reconstruction_loss_ref(rec, coeffs) + R.transcription_loss_ref(act, gt, True) + sum(R.consistency_loss_ref(trn_rec, trn_scr, trn)))
        total.backward()
        ts = TrainStep(model)
        out = ts.losses(audio.cuda(), gt.cuda())
        ts.backward(out['total'])
        num = den = 0.0
        for k, p in model.named_parameters():
            num += float((p.grad.cpu() - sdg[k].grad).norm()) ** 2
            den += float(sdg[k].grad.norm()) ** 2
        err = (num / den) ** 0.5
        lerr = abs(float(out['total'].detach()) - float(total.detach())) / float(total.detach())
        ok = err <= 3e-2 and lerr <= 3e-2
        print('ok  ' if ok else 'FAIL', tag, 'grad rel-L2', round(err, 4), 'loss rel', round(lerr, 4))
        bad += 0 if ok else 1
    except Exception as e:
        bad += 1
        print('EXC ', tag, repr(e)[:400])
        traceback.print_exc()

# ---- 2. forward_tiled ---------------------------------------------------------------------------------------------------------------
for cfg, latent, cx, skip in GEOS:
    try:
        model, sd, c = build(cfg, latent, cx, skip)
        model.eval()
        audio = tonal_clip(11 * c.block_length + 13, cfg['sample_rate'], seed=9, n_batch=2).cuda()
        audio = model.sliCQ.pad_to_block_length(audio)
        whole = model(audio, consistency=True)
        for tile in (128, 384, 1024):
            tiled = model.forward_tiled(audio, consistency=True, tile_frames=tile)
            ok = all(torch.equal(a, b) for a, b in zip(whole[:5], tiled[:5]))
            print('ok  ' if ok else 'FAIL', f"tiled M={c.max_window_length} F={c.n_bins} cx={cx} skip={skip} tile={tile}")
            bad += 0 if ok else 1
    except Exception as e:
        bad += 1
        print('EXC  tiled', cfg, repr(e)[:400])
        traceback.print_exc()

# ---- 3. block sharding, emulated ranks ----------------------------------------------------------------------------------------------
for cfg, latent, cx, skip in GEOS[:2]:
    try:
        model, sd, c = build(cfg, latent, cx, skip)
        model.eval()
        L = c.block_length
        for n in (L - 3, 4 * L, 7 * L + 5):
            audio = tonal_clip(n, cfg['sample_rate'], seed=n % 83).cuda()
            act = model.transcribe(audio)
            whole = model.reconstruct(audio)
            for world in (1, 2, 3, 5, 7):
                parts = [model.transcribe_sharded(audio, rank=r, world=world, gather=False) for r in range(world)]
                ok = torch.equal(torch.cat(parts, dim=-1), act)
                both = [model.transcribe_and_reconstruct_sharded(audio, rank=r, world=world, gather=False) for r in range(world)]
                ok = ok and torch.equal(torch.cat([a for a, _ in both], dim=-1), act)
                ok = ok and sum(w.size(-1) for _, w in both) == whole.size(-1)
                print('ok  ' if ok else 'FAIL', f"sharded M={c.max_window_length} n={n} world={world}")
                bad += 0 if ok else 1
    except Exception as e:
        bad += 1
        print('EXC  sharded', cfg, repr(e)[:400])
        traceback.print_exc()

torch.cuda.synchronize()
print('failures:', bad)
sys.exit(1 if bad else 0)
