"""
Forward half of the reference's loss step (reference: experiments/train.py:393-467): everything from the batch of audio to
the scalar losses, on the CUDA kernels.  Differences to the reference that do not change values:
  * the CQT target is computed ONCE and shared with the model forward (the reference computes it twice, train.py:404 and
    modules.py:366 -> :88);
  * no per-loss `.item()` host syncs: the losses come back as 0-dim device tensors.

The backward half (dgrad / wgrad kernels, clip, AdamW, the NCCL gradient all-reduce) is SURVEY.md section 8 row a16's remaining
work and is not implemented in this round: this function is for evaluation / validation losses and as the parity anchor
(tests/test_train_losses_gpu.py) for the numbers a training step must reproduce.
"""

import torch

from .objectives import compute_consistency_loss, compute_reconstruction_loss, compute_transcription_loss

__all__ = ['compute_step_losses']


def compute_step_losses(model, audio, ground_truth, multipliers=None, late_start=False):
    """
    audio (B, 1, n*L) on the GPU, ground_truth (B_mpe, F, T) with B_mpe <= B (train.py:393-394, 429).
    Returns a dict of 0-dim tensors: reconstruction, transcription, consistency_spectral, consistency_score, total
    (train.py:424-458; `late_start` = True reproduces the epochs before n_epochs_late_start, where only the reconstruction
    term enters the total).
    """
    mult = dict(reconstruction=1, transcription=1, consistency=1)
    mult.update(multipliers or {})
    with torch.no_grad():
        coefficients = model.sliCQ.encode_interleaved(audio)                      # (B, F, T, 2), the target, computed once
        lat, emb = model.encoder.forward_c8(coefficients)
        skips = model._skips_c8(emb)
        reconstruction = model.decoder.forward_c8(lat, True, skips).permute(0, 3, 1, 2)
        transcription_coeffs = model.decoder.forward_c8(lat, False, skips)
        transcription = model.to_activations(transcription_coeffs.permute(0, 3, 1, 2))
        n_mpe = ground_truth.size(0)
        losses = dict(reconstruction=compute_reconstruction_loss(reconstruction, coefficients.permute(0, 3, 1, 2)),
                      transcription=compute_transcription_loss(transcription[:n_mpe], ground_truth.float(), True))
        total = mult['reconstruction'] * losses['reconstruction']
        if mult['consistency']:
            lat_t, emb_t = model.encoder.forward_c8(transcription_coeffs)
            skips_t = model._skips_c8(emb_t)
            trn_rec = model.decoder.forward_c8(lat_t, True, skips_t).permute(0, 3, 1, 2)
            trn_scr = model.decoder.forward_c8(lat_t, False, skips_t).permute(0, 3, 1, 2)
            sp, sc = compute_consistency_loss(trn_rec[:n_mpe], trn_scr[:n_mpe], transcription_coeffs.permute(0, 3, 1, 2)[:n_mpe])
            losses['consistency_spectral'], losses['consistency_score'] = sp, sc
        if not late_start:
            total = total + mult['transcription'] * losses['transcription']
            if mult['consistency']:
                total = total + mult['consistency'] * (losses['consistency_spectral'] + losses['consistency_score'])
        losses['total'] = total
    return losses
