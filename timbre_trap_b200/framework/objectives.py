"""
Drop-in for `timbre_trap.framework.objectives` (reference: timbre_trap/framework/objectives.py:1-104): the three
objectives with the reference's signatures, computed by the deterministic reduction kernels of csrc/loss_kernels.cu.
Forward values only (the gradient kernels belong to the training step).
"""

import ctypes

import torch

from .. import _lib

__all__ = ['compute_reconstruction_loss', 'compute_transcription_loss', 'compute_consistency_loss']


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _same_order(a, b):
    """Two (B, C, F, T) tensors as flat fp32 buffers in one common memory order (free for the kernels' own outputs)."""
    a = a.detach().float()
    b = b.detach().float()
    if a.stride() == b.stride() and a.permute(0, 2, 3, 1).is_contiguous():
        return a.permute(0, 2, 3, 1), b.permute(0, 2, 3, 1)
    return a.contiguous(), b.contiguous()


def _scratch(device):
    return torch.empty(_lib.lib().tt_loss_scratch_floats(), dtype=torch.float32, device=device)


def compute_reconstruction_loss(reconstructed, target):
    """objectives.py:11-33: squared error summed over (C, F), averaged over (B, T)."""
    _lib.require_cuda(reconstructed, 'reconstructed')
    _lib.require_cuda(target, 'target')
    if reconstructed.shape != target.shape:
        raise ValueError(f'shape mismatch {tuple(reconstructed.shape)} vs {tuple(target.shape)}')
    a, b = _same_order(reconstructed, target)
    B, T = reconstructed.size(0), reconstructed.size(-1)
    out = torch.empty((), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().tt_sum_sq_diff(_p(a), _p(b), a.numel(), 1.0 / (B * T), _p(out), _p(_scratch(a.device)),
                                             ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)))
    return out


def compute_transcription_loss(estimate, target, weight_positive_class=False):
    """objectives.py:36-74."""
    _lib.require_cuda(estimate, 'estimate')
    _lib.require_cuda(target, 'target')
    if estimate.shape != target.shape or estimate.dim() != 3:
        raise ValueError('estimate and target must both be (B, F, T)')
    e = estimate.detach().float().contiguous()
    g = target.detach().float().contiguous()
    B, F, T = e.shape
    out = torch.empty((), dtype=torch.float32, device=e.device)
    with torch.cuda.device(e.device):
        _lib.check(_lib.lib().tt_transcription_loss(_p(e), _p(g), B, F, T, int(bool(weight_positive_class)), _p(out),
                                                    _p(_scratch(e.device)),
                                                    ctypes.c_void_p(torch.cuda.current_stream(e.device).cuda_stream)))
    return out


def compute_consistency_loss(spectral_coefficients, transcription_coefficients, target):
    """objectives.py:77-104."""
    return (compute_reconstruction_loss(spectral_coefficients, target),
            compute_reconstruction_loss(transcription_coefficients, target))
