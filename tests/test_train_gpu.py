"""
GPU parity of the backward half of the loss step (experiments/train.py:470-496): the generic gradient kernels against torch
autograd, the whole step's parameter gradients against the CPU oracle's autograd (fp32), and clip + AdamW against torch.optim.
Tolerance on gradients: the forward runs in bf16 and activation gradients travel in bf16 between layers (the reference trains
under fp16 autocast), so per-parameter gradients are compared at rel-L2 <= 1e-1 (the small bias vectors are the noisiest) and the global gradient at 3e-2.
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import tonal_clip

pytestmark = pytest.mark.gpu

SMALL = dict(sample_rate=8000, n_octaves=6, bins_per_octave=12, secs_per_block=0.5)


@pytest.mark.parametrize('Cin,Cout,KH,KW,sh,d,ph,pw,H', [(3, 5, 3, 3, 1, 2, 2, 2, 11), (4, 8, 4, 1, 2, 1, 0, 0, 14), (6, 2, 1, 1, 1, 1, 0, 0, 5),
                                                          (8, 16, 7, 1, 1, 1, 0, 0, 7), (2, 4, 3, 3, 1, 1, 1, 1, 9), (5, 3, 3, 3, 1, 3, 3, 3, 13)])
def test_generic_conv_kernels_match_autograd(Cin, Cout, KH, KW, sh, d, ph, pw, H):
    from timbre_trap_b200.framework import train as TR
    g = torch.Generator().manual_seed(KH * 100 + Cin)
    B, T = 2, 37
    x = torch.randn((B, Cin, H, T), generator=g, requires_grad=True)
    w = torch.randn((Cout, Cin, KH, KW), generator=g, requires_grad=True)
    b = torch.randn((Cout,), generator=g, requires_grad=True)
    y = F.elu(F.conv2d(x, w, b, stride=(sh, 1), padding=(ph, pw), dilation=(d if KH > 1 else 1, d if KW > 1 else 1)))
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    geom = TR._geom(KH, KW, sh=sh, dh=d if KH > 1 else 1, dw=d if KW > 1 else 1, ph=ph, pw=pw)
    xc, wc, bc = x.detach().cuda(), w.detach().cuda(), b.detach().cuda()
    yk = TR._conv_fwd(xc, wc, bc, geom, True)
    assert torch.allclose(yk.cpu(), y.detach(), atol=1e-4, rtol=1e-4)
    dz = TR._elu_bwd(gy.cuda(), yk)
    dx = TR._conv_bwd_data(dz, wc, xc.shape, geom)
    dw, db = TR._conv_bwd_weight(xc, dz, wc.shape, geom, True)
    assert torch.allclose(dx.cpu(), x.grad, atol=2e-4, rtol=1e-3)
    assert torch.allclose(dw.cpu(), w.grad, atol=2e-3, rtol=1e-3)
    assert torch.allclose(db.cpu(), b.grad, atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize('Cin,Cout,KH,sh,op,H', [(6, 4, 4, 2, 1, 9), (8, 3, 4, 2, 0, 6), (5, 7, 6, 1, 0, 1)])
def test_transposed_roles(Cin, Cout, KH, sh, op, H):
    """convT forward / backward through the regular-conv kernels with the roles swapped (DecoderBlock.tconv, Decoder.convin)."""
    from timbre_trap_b200.framework import train as TR
    g = torch.Generator().manual_seed(KH + Cin)
    B, T = 2, 21
    x = torch.randn((B, Cin, H, T), generator=g, requires_grad=True)
    w = torch.randn((Cin, Cout, KH, 1), generator=g, requires_grad=True)
    y = F.conv_transpose2d(x, w, None, stride=(sh, 1), output_padding=(op, 0))
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    geom = TR._geom(KH, 1, sh=sh)
    wc, gyc, xc = w.detach().cuda(), gy.cuda(), x.detach().cuda()
    yk = TR._conv_bwd_data(xc, wc, (B, Cout, y.shape[2], T), geom)          # convT forward = bwd_data
    assert torch.allclose(yk.cpu(), y.detach(), atol=1e-4, rtol=1e-4)
    gx = TR._conv_fwd(gyc, wc, None, geom, False)                            # convT backward-data = regular forward
    assert gx.shape == x.shape and torch.allclose(gx.cpu(), x.grad, atol=2e-4, rtol=1e-3)
    dw, _ = TR._conv_bwd_weight(gyc, xc, wc.shape, geom, False)
    assert torch.allclose(dw.cpu(), w.grad, atol=2e-3, rtol=1e-3)
    assert torch.allclose(TR._channel_sum(gyc).cpu(), gy.sum(dim=(0, 2, 3)), atol=1e-3, rtol=1e-4)


def _setup(skip=False, latent=None):
    from oracle import model_ref as R
    from timbre_trap_b200.framework import TimbreTrap
    model = TimbreTrap(latent_size=latent, model_complexity=1, skip_connections=skip, **SMALL)
    sd = R.init_state_dict(model.sliCQ.n_bins, latent, 1, seed=3)
    if skip:
        sd['skip_weights'] = torch.tensor([0.9, 1.1, 0.8, 1.2, 0.7])
    model.load_state_dict(sd)
    model = model.cuda()
    c = R.CQTRef(SMALL['n_octaves'], SMALL['bins_per_octave'], SMALL['sample_rate'], SMALL['secs_per_block'])
    audio = tonal_clip(2 * c.block_length, SMALL['sample_rate'], seed=4, n_batch=3)
    rng = np.random.default_rng(2)
    gt = torch.zeros((2, c.n_bins, 2 * c.max_window_length))
    for b in range(2):
        for k in rng.integers(5, c.n_bins - 5, size=3):
            gt[b, k, :] = 1.0
            gt[b, k - 1, :] = gt[b, k + 1, :] = 0.6
    return R, model, sd, c, audio, gt


def _oracle_grads(R, sd, c, audio, gt):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    coeffs = c(audio)
    rec, lat, trn, trn_rec, trn_scr = R.forward_ref(audio, sd, c, consistency=True)
    act = torch.tanh(c.to_magnitude(trn))
    losses = dict(reconstruction=R.reconstruction_loss_ref(rec, coeffs), transcription=R.transcription_loss_ref(act[:2], gt, True))
    losses['consistency_spectral'], losses['consistency_score'] = R.consistency_loss_ref(trn_rec[:2], trn_scr[:2], trn[:2])
    total = sum(losses.values())
    total.backward()
    return {k: v.grad for k, v in sd.items()}, losses, float(total.detach())


@pytest.mark.parametrize('skip,latent', [(False, None), (True, None), (False, 40)])     # 40 latent channels: padded to 64 in the kernels
def test_step_gradients_match_oracle_autograd(skip, latent):
    from timbre_trap_b200.framework.train import TrainStep
    R, model, sd, c, audio, gt = _setup(skip, latent)
    want, want_losses, want_total = _oracle_grads(R, sd, c, audio, gt)
    ts = TrainStep(model)
    out = ts.losses(audio.cuda(), gt.cuda())
    np.testing.assert_allclose(float(out['total'].detach()), want_total, rtol=3e-2)
    ts.backward(out['total'])
    got = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    assert set(got) == set(want)
    total_norm = sum(float(v.norm()) ** 2 for v in want.values()) ** 0.5
    num = den = 0.0
    worst = (0.0, None)
    for k in want:
        err, ref = float((got[k] - want[k]).norm()), float(want[k].norm())
        num += err ** 2
        den += ref ** 2
        if err / max(ref, 1e-12) > worst[0]:
            worst = (err / max(ref, 1e-12), k)
        # relative to the parameter's own gradient, with an absolute floor for the tiny bias vectors (bf16 noise does not shrink with them)
        assert err <= 1e-1 * ref + 1e-2 * total_norm, (k, err, ref, total_norm)
    print('global gradient rel-L2 error', (num / den) ** 0.5, 'worst parameter', worst)
    assert (num / den) ** 0.5 <= 3e-2, ((num / den) ** 0.5, worst)


def test_clip_and_adamw_match_torch():
    from timbre_trap_b200.framework.train import TrainStep
    R, model, sd, c, audio, gt = _setup()
    ts = TrainStep(model, lr=1e-3, max_norm=10.0)
    ref_params = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    opt = torch.optim.AdamW(ref_params, lr=1e-3)
    g = torch.Generator(device='cuda').manual_seed(0)
    for step in range(3):
        for p, q in zip(model.parameters(), ref_params):
            p.grad = torch.randn(p.shape, device='cuda', generator=g) * (30.0 if step == 1 else 0.01)   # step 1 is clipped
            q.grad = p.grad.clone()
        norm = ts.optimizer_step()
        ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, 10.0)
        opt.step()
        assert abs(float(norm) - float(ref_norm)) <= 1e-4 * float(ref_norm)
        for p, q in zip(model.parameters(), ref_params):
            assert torch.allclose(p.detach(), q.detach(), atol=1e-6, rtol=1e-5)


def test_training_reduces_the_loss():
    from timbre_trap_b200.framework.train import TrainStep
    R, model, sd, c, audio, gt = _setup()
    ts = TrainStep(model, lr=2e-3)
    first = float(ts.step(audio.cuda(), gt.cuda())['total'])
    for _ in range(7):
        last = ts.step(audio.cuda(), gt.cuda())
    assert float(last['total']) < 0.8 * first, (first, float(last['total']))
    assert np.isfinite(float(last['grad_norm']))


@pytest.mark.parametrize('C,H,T,k,d,B', [(4, 23, 256, 3, 1, 2), (8, 17, 384, 3, 2, 1), (16, 33, 200, 3, 3, 2), (32, 9, 128, 3, 1, 3), (32, 40, 512, 1, 1, 1),
                                           (3, 11, 136, 1, 1, 2), (16, 300, 128, 3, 2, 1), (8, 20, 256, 3, 3, 1), (2, 35, 640, 3, 3, 2)])
def test_tensor_core_weight_gradient(C, H, T, k, d, B):
    """tt_conv_wgrad_same (MN-major tcgen05 GEMM over the pixel axis, deterministic two-stage reduction) against torch autograd on the
    same bf16-rounded operands; products of bf16 values are exact in fp32, so only the summation order differs."""
    import torch.nn.functional as F
    from timbre_trap_b200.framework import packing as P, train as TR
    g = torch.Generator().manual_seed(C * 100 + H)
    x = torch.randn((B, C, H, T), generator=g).to(torch.bfloat16).float()
    dz = (torch.randn((B, C, H, T), generator=g) * 0.1).to(torch.bfloat16).float()
    w = torch.zeros((C, C, k, k), requires_grad=True)
    bias = torch.zeros(C, requires_grad=True)
    pad = d if k == 3 else 0
    y = F.conv2d(x, w, bias, padding=pad, dilation=d if k == 3 else 1)
    (y * dz).sum().backward()
    dw, db = TR._wgrad_same(P.to_c8(x.cuda()), P.to_c8(dz.cuda()), C, C, k, d)
    assert dw.shape == w.grad.shape
    scale = float(w.grad.abs().max())
    assert float((dw.cpu() - w.grad).abs().max()) <= 2e-4 * scale + 1e-4, (float((dw.cpu() - w.grad).abs().max()), scale)
    assert float((db.cpu() - bias.grad).abs().max()) <= 2e-4 * float(bias.grad.abs().max()) + 1e-4
    dw2, db2 = TR._wgrad_same(P.to_c8(x.cuda()), P.to_c8(dz.cuda()), C, C, k, d)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)                  # fixed reduction order: bit-reproducible


@pytest.mark.parametrize('Cf,Cc,Hc,T,op,B', [(4, 8, 19, 256, 0, 2), (8, 16, 33, 200, 1, 1), (16, 32, 9, 384, 0, 2), (32, 64, 31, 128, 1, 2)])
def test_tensor_core_weight_gradient_strided_and_transposed(Cf, Cc, Hc, T, op, B):
    """tt_conv_wgrad_updown against torch autograd: EncoderBlock.sconv (fine = input, coarse = output gradient) and DecoderBlock.tconv
    (coarse = input, fine = output gradient, bias gradient = pixel sums of the fine tensor)."""
    from timbre_trap_b200.framework import packing as P, train as TR
    g = torch.Generator().manual_seed(Cf * 10 + Hc)
    Hf = 2 * Hc + 2 + op
    fine = torch.randn((B, Cf, Hf, T), generator=g).to(torch.bfloat16).float()
    coarse = (torch.randn((B, Cc, Hc, T), generator=g) * 0.1).to(torch.bfloat16).float()
    # strided conv: y = conv(fine, w (Cc, Cf, 4, 1), stride 2) has (Hf - 4) // 2 + 1 = Hc rows; coarse plays dL/dy
    w = torch.zeros((Cc, Cf, 4, 1), requires_grad=True)
    bias = torch.zeros(Cc, requires_grad=True)
    y = F.conv2d(fine, w, bias, stride=(2, 1))
    assert y.shape[2] == Hc
    (y * coarse).sum().backward()
    dw, db = TR._wgrad_updown(P.to_c8(fine.cuda()), P.to_c8(coarse.cuda()), Cf, Cc, False)
    assert float((dw.cpu() - w.grad).abs().max()) <= 2e-4 * float(w.grad.abs().max()) + 1e-4
    assert float((db.cpu() - bias.grad).abs().max()) <= 2e-4 * float(bias.grad.abs().max()) + 1e-4
    # transposed conv: y = convT(coarse, wt (Cc, Cf, 4, 1), stride 2, output_padding op) has Hf rows; fine plays dL/dy
    wt = torch.zeros((Cc, Cf, 4, 1), requires_grad=True)
    bt = torch.zeros(Cf, requires_grad=True)
    yt = F.conv_transpose2d(coarse, wt, bt, stride=(2, 1), output_padding=(op, 0))
    assert yt.shape[2] == Hf
    (yt * fine).sum().backward()
    dwt, dbt = TR._wgrad_updown(P.to_c8(fine.cuda()), P.to_c8(coarse.cuda()), Cf, Cc, True)
    assert float((dwt.cpu() - wt.grad).abs().max()) <= 2e-4 * float(wt.grad.abs().max()) + 1e-4
    assert float((dbt.cpu() - bt.grad).abs().max()) <= 2e-4 * float(bt.grad.abs().max()) + 2e-3


@pytest.mark.parametrize('Ct,Cf,H,T,B', [(64, 128, 31, 256, 2), (32, 32, 5, 200, 3), (64, 100, 31, 128, 1)])
def test_tensor_core_weight_gradient_tall_kernel_layers(Ct, Cf, H, T, B):
    """tt_conv_wgrad_lat against torch autograd: Encoder.convlat (Conv2d (Cf, Ct, H, 1) over the full height) and Decoder.convin
    (ConvTranspose2d (Cf + 1, Ct, H, 1) with the indicator channel: its weight row and the bias are row sums of the output gradient)."""
    from timbre_trap_b200.framework import packing as P, train as TR
    g = torch.Generator().manual_seed(Ct + H)
    tall = torch.randn((B, Ct, H, T), generator=g).to(torch.bfloat16).float()
    flat = (torch.randn((B, Cf, 1, T), generator=g) * 0.1).to(torch.bfloat16).float()
    fpad = (Cf + 15) // 16 * 16
    flat8 = TR.P.to_c8(torch.nn.functional.pad(flat, (0, 0, 0, 0, 0, fpad - Cf)).cuda())
    tall8 = P.to_c8(tall.cuda())
    # convlat: y = conv(tall, w (Cf, Ct, H, 1)); flat plays dL/dy
    w = torch.zeros((Cf, Ct, H, 1), requires_grad=True)
    bias = torch.zeros(Cf, requires_grad=True)
    (F.conv2d(tall, w, bias) * flat).sum().backward()
    dw = torch.zeros((Cf, Ct, H, 1), device='cuda')
    db = torch.zeros(Cf, device='cuda')
    TR._wgrad_lat(tall8, flat8, Ct, Cf, dw, db, None)
    assert float((dw.cpu() - w.grad).abs().max()) <= 2e-4 * float(w.grad.abs().max()) + 1e-4
    assert float((db.cpu() - bias.grad).abs().max()) <= 2e-4 * float(bias.grad.abs().max()) + 1e-4
    # decoder convin: y = convT(cat(flat, indicator), wt (Cf + 1, Ct, H, 1)); tall plays dL/dy
    wt = torch.zeros((Cf + 1, Ct, H, 1), requires_grad=True)
    bt = torch.zeros(Ct, requires_grad=True)
    full = torch.cat((flat, torch.ones_like(flat[:, :1])), dim=1)
    (F.conv_transpose2d(full, wt, bt) * tall).sum().backward()
    dwt = torch.zeros((Cf + 1, Ct, H, 1), device='cuda')
    rows = torch.zeros((Ct, H), device='cuda')
    TR._wgrad_lat(tall8, flat8, Ct, Cf, dwt, None, rows)
    dwt[Cf, :, :, 0] = rows
    assert float((dwt.cpu() - wt.grad).abs().max()) <= 2e-4 * float(wt.grad.abs().max()) + 2e-3
    assert float((rows.sum(1).cpu() - bt.grad).abs().max()) <= 2e-4 * float(bt.grad.abs().max()) + 2e-3
