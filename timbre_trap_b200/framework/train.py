"""
The reference's loss step (reference: experiments/train.py:393-500) on the CUDA kernels.

  compute_step_losses(model, audio, ground_truth)   forward half only (no graph): evaluation / validation losses
  TrainStep(model, ...).step(audio, ground_truth)   the whole step: CQT target (once), forward with consistency, the four
                                                    losses, backward, gradient all-reduce (NCCL, when a process group is
                                                    given), clip_grad_norm_(10), AdamW

Forward kernels are the inference kernels (tcgen05 strips); every layer is a torch.autograd.Function whose backward calls
the native gradient kernels of csrc/train_kernels.cu (direct-convolution dgrad / tiled wgrad on fp32 NCHW, CUDA cores); the
residual blocks recompute their intermediate and take their data gradients on the tensor cores (tt_conv_same, the gradient
operand split into bf16 hi + lo parts).  PyTorch's autograd engine only walks the graph and accumulates `.grad`; there is no cuDNN / ATen compute on the
path except layout conversions (permute / dtype copies) between the inference layouts (C8 planar / packed4 bf16) and NCHW.
Differences to the reference that do not change values: the CQT target is computed once (the reference computes it twice,
train.py:404 and modules.py:366 -> :88); no `.item()` host syncs inside the step; activation gradients travel in bf16 between
layers (the reference runs the forward under fp16 autocast).
"""

import ctypes

import torch

from .. import _lib
from . import ops
from . import packing as P
from .cqt import CQT
from .modules import _PackedCache
from .objectives import compute_consistency_loss, compute_reconstruction_loss, compute_transcription_loss

__all__ = ['compute_step_losses', 'TrainStep', 'allreduce_mean_gradients']


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _s(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# ---------------------------------------------------------------------------------------------------------------
# forward half without a graph
# ---------------------------------------------------------------------------------------------------------------
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


# ---------------------------------------------------------------------------------------------------------------
# native gradient kernels on fp32 NCHW
# ---------------------------------------------------------------------------------------------------------------
def _geom(kh, kw, sh=1, dh=1, dw=1, ph=0, pw=0):
    return dict(KH=kh, KW=kw, sh=sh, dh=dh, dw=dw, ph=ph, pw=pw)


def _conv_fwd(x, w, bias, g, act):
    B, Cin, Hin, T = x.shape
    Cout = w.size(0)
    Hout = (Hin + 2 * g['ph'] - g['dh'] * (g['KH'] - 1) - 1) // g['sh'] + 1
    y = torch.empty((B, Cout, Hout, T), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().tt_conv_fwd_f32(_p(x), _p(w), _p(bias), _p(y), B, Cin, Hin, T, Cout, g['KH'], g['KW'], g['sh'], g['dh'],
                                          g['dw'], g['ph'], g['pw'], int(act), _s(x)))
    return y


def _conv_bwd_data(dz, w, x_shape, g):
    """dz (B, Cout, Hout, T), w (Cout, Cin, KH, KW) -> dx of shape x_shape (B, Cin, Hin, T)."""
    B, Cin, Hin, T = x_shape
    Cout, Hout = dz.size(1), dz.size(2)
    dx = torch.empty(x_shape, dtype=torch.float32, device=dz.device)
    _lib.check(_lib.lib().tt_conv_bwd_data_f32(_p(dz), _p(w), _p(dx), B, Cin, Hin, T, Cout, g['KH'], g['KW'], g['sh'], g['dh'],
                                               g['dw'], g['ph'], g['pw'], Hout, _s(dz)))
    return dx


def _conv_bwd_weight(x, dz, w_shape, g, want_bias):
    """x (B, Cin, Hin, T), dz (B, Cout, Hout, T) -> (dW of shape w_shape, db (Cout) or None)."""
    B, Cin, Hin, T = x.shape
    Cout, Hout = dz.size(1), dz.size(2)
    dw = torch.zeros(w_shape, dtype=torch.float32, device=x.device)
    db = torch.zeros(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    _lib.check(_lib.lib().tt_conv_bwd_weight_f32(_p(x), _p(dz), _p(dw), _p(db), B, Cin, Hin, T, Cout, g['KH'], g['KW'], g['sh'],
                                                 g['dh'], g['dw'], g['ph'], g['pw'], Hout, _s(x)))
    return dw, db


def _elu_bwd(dy, a):
    dz = torch.empty_like(dy)
    _lib.check(_lib.lib().tt_elu_bwd(_p(dy), _p(a), _p(dz), dy.numel(), _s(dy)))
    return dz


def _channel_sum(dz):
    B, C = dz.shape[:2]
    db = torch.zeros(C, dtype=torch.float32, device=dz.device)
    _lib.check(_lib.lib().tt_channel_sum(_p(dz), _p(db), B, C, dz.numel() // (B * C), _s(dz)))
    return db


# layout plumbing: inference layouts <-> NCHW fp32
def _nchw(t, c):
    """internal activation (C8 planar 5-D, packed4 4-D bf16, or interleaved coefficients (B,F,T,2) fp32) -> (B, c, H, T) fp32."""
    if t.dtype == torch.float32:
        return t.permute(0, 3, 1, 2).contiguous()
    return (P.from_p4(t, c) if t.dim() == 4 else P.from_c8(t, c)).contiguous()


def _like(nchw, ref):
    """(B, c, H, T) fp32 gradient -> the layout / dtype of the forward tensor `ref`."""
    if ref.dtype == torch.float32:
        return nchw.permute(0, 2, 3, 1).contiguous()
    return P.to_p4(nchw) if ref.dim() == 4 else P.to_c8(nchw)


# ---------------------------------------------------------------------------------------------------------------
# autograd Functions: fast forward kernel + native backward kernels
# ---------------------------------------------------------------------------------------------------------------
class _ConvFn(torch.autograd.Function):
    """A regular conv layer (+ optional ELU): forward through `run(x)` (an inference kernel), backward generic."""

    @staticmethod
    def forward(ctx, x, weight, bias, run, geom, cin, cout, act):
        y = run(x)
        ctx.save_for_backward(x, y, weight)
        ctx.meta = (geom, cin, cout, act)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight = ctx.saved_tensors
        geom, cin, cout, act = ctx.meta
        xn, dy = _nchw(x, cin), _nchw(gy, cout)
        dz = _elu_bwd(dy, _nchw(y, cout)) if act else dy
        w = weight.detach().float().contiguous()
        dw, db = _conv_bwd_weight(xn, dz, w.shape, geom, True)
        gx = _like(_conv_bwd_data(dz, w, xn.shape, geom), x) if ctx.needs_input_grad[0] else None
        return gx, dw, db, None, None, None, None, None


class _ConvTFn(torch.autograd.Function):
    """A transposed conv layer + ELU (weight (Cin, Cout, KH, KW)); optional per-row bias table handled by the caller."""

    @staticmethod
    def forward(ctx, x, weight, bias, run, geom, cin, cout):
        y = run(x)
        ctx.save_for_backward(x, y, weight)
        ctx.meta = (geom, cin, cout)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight = ctx.saved_tensors
        geom, cin, cout = ctx.meta
        xn = _nchw(x, cin)
        dz = _elu_bwd(_nchw(gy, cout), _nchw(y, cout))
        w = weight.detach().float().contiguous()               # (cin, cout, KH, KW) = the regular conv y-space -> x-space
        # weight gradient: regular-conv roles swapped (its input is dz, its output-side gradient is x)
        dw, _ = _conv_bwd_weight(dz, xn, w.shape, geom, False)
        db = _channel_sum(dz)
        gx = _like(_conv_fwd(dz, w, None, geom, False), x)
        return gx, dw, db, None, None, None, None


def _conv_same_hi_lo(dz, wp, c, k, d):
    """
    Data-gradient conv on the tensor cores with the gradient operand split into bf16 hi + lo parts (two passes, summed in fp32):
    the weights are the bf16 values the forward pass used (so they are exact for the function being differentiated), and the split
    keeps ~16 mantissa bits of dz - weight gradients are long sums with heavy cancellation and do not tolerate 8-bit gradients.
    """
    hi = P.to_c8(dz)
    lo = P.to_c8(dz - P.from_c8(hi, c))
    return (P.from_c8(ops.conv_same(hi, wp, None, k, d), c) + P.from_c8(ops.conv_same(lo, wp, None, k, d), c)).contiguous()


# ---- residual blocks: the whole backward in the inference layouts (bf16), convolutions AND weight gradients on the tensor cores ----
def _ew(fn_name, *tensors):
    """element-wise bf16 kernel over tensors of one common layout -> new tensor like the first"""
    out = torch.empty_like(tensors[0])
    _lib.check(getattr(_lib.lib(), fn_name)(*[_p(t) for t in tensors], _p(out), out.numel(), _s(out)))
    return out


def _as_c8(t):
    """packed 4-channel (B, H, T, 4) -> C8 planar (B, 1, H, T, 8) (native re-layout kernel); C8 tensors pass through"""
    if t.dim() == 5:
        return t
    B, H, T, _ = t.shape
    out = torch.empty((B, 1, H, T, 8), dtype=torch.bfloat16, device=t.device)
    _lib.check(_lib.lib().tt_p4_to_c8(_p(t), _p(out), B * H * T, _s(t)))
    return out


def _to_layout_of(c8, ref):
    """C8 planar (B, 1, H, T, 8) -> the layout of `ref` (packed 4-channel or C8)"""
    if ref.dim() == 5:
        return c8
    out = torch.empty_like(ref)
    _lib.check(_lib.lib().tt_c8_to_p4(_p(c8), _p(out), out.numel() // 4, _s(out)))
    return out


_WGRAD_SCRATCH = {}
_BW_PACKS = {'epoch': -1}


def _scratch(pool, device, n):
    """One growing fp32 scratch buffer per device (the weight-gradient partials): reused by every layer, reallocated only to grow."""
    buf = pool.get(device)
    if buf is None or buf.numel() < n:
        buf = torch.empty(n, dtype=torch.float32, device=device)
        pool[device] = buf
    return buf


def _bw_pack(weight, tag, build):
    """Packed (kernel-layout) weights of the backward pass, built once per optimisation step per layer: the same layer runs backward
    two to four times per step (two encoder, four decoder passes), and every pack is a handful of tiny launches."""
    if _BW_PACKS['epoch'] != _PackedCache.epoch:
        _BW_PACKS.clear()
        _BW_PACKS['epoch'] = _PackedCache.epoch
    key = (weight.data_ptr(), weight._version, tag)
    if key not in _BW_PACKS:
        with torch.no_grad():
            _BW_PACKS[key] = build()
    return _BW_PACKS[key]


def _wgrad_same(x8, dz8, cin, cout, k, d):
    """(dW (cout, cin, k, k), db (cout)) fp32 of a 'same' conv from C8 planar bf16 x and dz (tt_conv_wgrad_same, tensor cores)."""
    B, CGi, H, T, _ = x8.shape
    lib = _lib.lib()
    scratch = _scratch(_WGRAD_SCRATCH, x8.device, int(lib.tt_wgrad_scratch_floats(B, H, T)))
    dw = torch.zeros((cout, cin, k, k), dtype=torch.float32, device=x8.device)
    db = torch.zeros(cout, dtype=torch.float32, device=x8.device)
    _lib.check(lib.tt_conv_wgrad_same(_p(x8), _p(dz8), _p(dw), _p(db), B, CGi * 8, dz8.size(1) * 8, cin, cout, H, T, k, d,
                                      _p(scratch), _s(x8)))
    return dw, db


class _ResFn(torch.autograd.Function):
    """ResidualConv2dBlock: fused forward kernel, which also writes the inner activation ELU(W1 * x + b1) it stages in shared memory
    (tt_res_block_rs_mid) - the backward needs it and a recompute costs a whole 3x3 conv launch per block; the backward stays in bf16
    C8 planar end to end - both data-gradient convolutions through the tile kernel (tt_conv_same), both weight gradients through the
    MN-major tcgen05 kernels (tt_conv_wgrad_same), the ELU derivatives / residual add as one-pass element-wise kernels."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, run, c, d):
        a1 = torch.empty_like(x)
        y = run(x, a1)
        ctx.save_for_backward(x, y, a1, w1, b1, w2)
        ctx.meta = (c, d)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, a1, w1, b1, w2 = ctx.saved_tensors
        c, d = ctx.meta
        w2_t = _bw_pack(w2, 'res1x1T', lambda: P.pack_res1x1(w2.detach().float().transpose(0, 1).contiguous()))
        w1_t = _bw_pack(w1, 'res3x3T', lambda: P.pack_res3x3(w1.detach().float().transpose(0, 1).flip(2, 3).contiguous()))
        gy = gy.contiguous()
        dz2 = _as_c8(_ew('tt_res_out_bwd_bf16', gy, y, x))                       # gy * ELU'(z2), activated 1x1 output = y - x
        x8 = _as_c8(x)
        a1 = _as_c8(a1)                                                          # the inner activation, as the forward staged it
        dw2, db2 = _wgrad_same(a1, dz2, c, c, 1, 1)
        dz1 = ops.conv_same(dz2, w2_t, None, 1, 1, times_elu_grad_of=a1)         # W2^T dz2, times ELU'(z1): one launch
        dw1, db1 = _wgrad_same(x8, dz1, c, c, 3, d)
        if gy.dim() == 5:
            gx = ops.conv_same(dz1, w1_t, None, 3, d, plus=gy)                   # gx = gy + W1^T (*) dz1: the residual add in the epilogue
        else:
            gx = _to_layout_of(ops.conv_same(dz1, w1_t, None, 3, d), x)          # packed 4-channel stage: re-layout, then add
            one = torch.ones((), dtype=torch.float32, device=gx.device)
            _lib.check(_lib.lib().tt_add_scaled_bf16(_p(gy), _p(gx), _p(one.reshape(1)), _p(gx), gx.numel(), _s(gx)))
        return gx, dw1, db1, dw2, db2, None, None, None


def _wgrad_updown(fine8, coarse8, cfine, ccoarse, transposed):
    """Weight / bias gradients of a (4,1) stride-(2,1) layer from C8 planar bf16 tensors (tt_conv_wgrad_updown, tensor cores):
    sconv: (fine = layer input, coarse = dz) -> dW (ccoarse, cfine, 4, 1), db (ccoarse); tconv (transposed): (fine = dz, coarse = layer
    input) -> dW (ccoarse, cfine, 4, 1) in the ConvTranspose2d layout, db (cfine)."""
    B, CGf, Hf, T, _ = fine8.shape
    Hc = coarse8.size(2)
    lib = _lib.lib()
    scratch = _scratch(_WGRAD_SCRATCH, fine8.device, int(lib.tt_wgrad_scratch_floats(B, Hc, T)))
    dw = torch.zeros((ccoarse, cfine, 4, 1), dtype=torch.float32, device=fine8.device)
    db = torch.zeros(cfine if transposed else ccoarse, dtype=torch.float32, device=fine8.device)
    _lib.check(lib.tt_conv_wgrad_updown(_p(fine8), _p(coarse8), _p(dw), _p(db), B, CGf * 8, coarse8.size(1) * 8, cfine, ccoarse, Hf, Hc, T,
                                        int(transposed), _p(scratch), _s(fine8)))
    return dw, db


class _DownFn(torch.autograd.Function):
    """EncoderBlock.sconv + ELU (modules.py:626-629).  Backward in the inference layouts: ELU derivative (element-wise), weight gradient
    (tensor-core GEMM over the pixel axis), data gradient = the transposed-conv forward kernel without activation on W seen as a
    ConvTranspose2d weight (in = Cout, out = Cin)."""

    @staticmethod
    def forward(ctx, x, weight, bias, run, cin, cout):
        y = run(x)
        ctx.save_for_backward(x, y, weight)
        ctx.meta = (cin, cout)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight = ctx.saved_tensors
        cin, cout = ctx.meta
        dz = _ew('tt_elu_bwd_bf16', gy.contiguous(), y)
        x8 = _as_c8(x)
        dw, db = _wgrad_updown(x8, dz, cin, cout, False)
        gx = None
        if ctx.needs_input_grad[0]:
            hin, hout = x8.size(2), dz.size(2)
            wt = _bw_pack(weight, 'downT', lambda: P.pack_up_strip(weight.detach().float(), torch.zeros(cin, device=weight.device)))
            gx = ops.conv_up_strip(dz, wt, P.pad8(cin), hin - 2 * hout - 2, packed4_out=x.dim() == 4, act=False)
        return gx, dw, db, None, None, None


class _UpFn(torch.autograd.Function):
    """DecoderBlock.tconv + ELU (modules.py:685-688).  Backward: data gradient = the strided-conv forward kernel without activation on
    W^T seen as a Conv2d weight (out = Cin, in = Cout)."""

    @staticmethod
    def forward(ctx, x, weight, bias, run, cin, cout):
        y = run(x)
        ctx.save_for_backward(x, y, weight)
        ctx.meta = (cin, cout)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight = ctx.saved_tensors
        cin, cout = ctx.meta
        dz = _ew('tt_elu_bwd_bf16', gy.contiguous(), y)                          # layout of y (packed 4-channel for the last stage)
        dw, db = _wgrad_updown(_as_c8(dz), x, cout, cin, True)
        # ConvTranspose2d (cin, cout, 4, 1) read as a Conv2d weight (out = cin, in = cout)
        pack = P.pack_down_pairs if dz.dim() == 4 else P.pack_down_strip
        wd = _bw_pack(weight, 'upT', lambda: pack(weight.detach().float().contiguous(), torch.zeros(cin, device=weight.device)))
        gx = ops.conv_down_strip(dz, wd, P.pad8(cin), act=False)
        return gx, dw, db, None, None, None


_LAT_SCRATCH = {}


def _wgrad_lat(tall8, flat8, ctall, cflat, dw, db, row_sums):
    """tt_conv_wgrad_lat: accumulates into dw (cflat.., ctall, H, 1) [first cflat rows], db (cflat) and row_sums (ctall, H) (either may be None)."""
    B, CGt, H, T, _ = tall8.shape
    lib = _lib.lib()
    scratch = _scratch(_LAT_SCRATCH, tall8.device, int(lib.tt_wgrad_lat_scratch_floats(B, H, T)))
    _lib.check(lib.tt_conv_wgrad_lat(_p(tall8), _p(flat8), _p(dw), _p(db), _p(row_sums), B, CGt * 8, flat8.size(1) * 8, ctall, cflat, H, T,
                                     _p(scratch), _s(tall8)))


def _lat_tc_ok(c_tall, c_flat_pad):
    """The tensor-core kernels of the two (H, 1)-kernel layers take up to 64 embedding / 128 latent channels (the base model's sizes)."""
    return P.pad8(c_tall) in (16, 32, 64) and c_flat_pad <= 128


class _LatFn(torch.autograd.Function):
    """Encoder.convlat (modules.py:446, no activation).  Backward in the inference layouts: weight gradient = tap-grouped tensor-core GEMM
    over (b, t); data gradient = the Decoder.convin forward kernel (a (H, 1) transposed conv) without activation on the same weights."""

    @staticmethod
    def forward(ctx, x, weight, bias, run, c4, d, d_pad):
        y = run(x)
        ctx.save_for_backward(x, weight)
        ctx.meta = (c4, d, d_pad)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        c4, d, d_pad = ctx.meta
        gy = gy.contiguous()
        dw = torch.zeros(weight.shape, dtype=torch.float32, device=x.device)
        db = torch.zeros(d, dtype=torch.float32, device=x.device)
        _wgrad_lat(x, gy, c4, d, dw, db, None)
        # gx[ci, h, t] = sum_co W[co, ci, h] gy[co, t]: ConvTranspose2d with weight (in = co, out = ci, H, 1)
        w, table = _bw_pack(weight, 'latT', lambda: P.pack_deconv_in_film(weight.detach().float(), torch.zeros(c4, device=x.device),
                                                                          torch.ones(d, device=x.device), torch.zeros(d, device=x.device), d_pad))
        gx = ops.deconv_in(gy, w, table, P.pad8(c4), x.size(2), act=False)
        return gx, dw, db, None, None, None, None


def _pairs_to_c8(pairs):
    """fp32 interleaved (B, F, T, 2) -> C8 planar bf16 (B, 1, F, T, 8) (tt_pairs_to_c8)."""
    pairs = pairs.contiguous()
    B, F_, T, _ = pairs.shape
    out = torch.empty((B, 1, F_, T, 8), dtype=torch.bfloat16, device=pairs.device)
    _lib.check(_lib.lib().tt_pairs_to_c8(_p(pairs), B * F_ * T, _p(out), _s(pairs)))
    return out


def _edge_tc_ok(x_like, c):
    """The native backward of the first / last 3x3 conv needs the packed 4-channel stage layout (model_complexity 1 and 2)."""
    return x_like.dim() == 4 and c <= 4


class _InFn(torch.autograd.Function):
    """Encoder.convin + ELU (modules.py:430-433) with a packed 4-channel output.  Backward: ELU derivative on the packed tensor; weight
    gradient on the tensor cores (both operands re-laid out to C8 planar bf16 by one-pass kernels); data gradient (only the consistency
    pass asks for it) = the Decoder.convout forward kernel on transposed, flipped weights."""

    @staticmethod
    def forward(ctx, coeffs, weight, bias, run, c0):
        y = run(coeffs)
        ctx.save_for_backward(coeffs, y, weight)
        ctx.c0 = c0
        return y

    @staticmethod
    def backward(ctx, gy):
        coeffs, y, weight = ctx.saved_tensors
        c0 = ctx.c0
        dz = _ew('tt_elu_bwd_bf16', gy.contiguous(), y)
        dw, db = _wgrad_same(_pairs_to_c8(coeffs), _as_c8(dz), 2, c0, 3, 1)
        gx = None
        if ctx.needs_input_grad[0]:
            wt, zb = _bw_pack(weight, 'inT', lambda: (weight.detach().float().transpose(0, 1).flip(2, 3).contiguous(),      # (2, c0, 3, 3)
                                                      torch.zeros(2, device=dz.device)))
            gx = ops.conv_out(dz, wt, zb, c0)
        return gx, dw, db, None, None


class _OutFn(torch.autograd.Function):
    """Decoder.convout (modules.py:543, no activation) on a packed 4-channel input.  Backward: weight gradient on the tensor cores;
    data gradient = the Encoder.convin forward kernel without ELU on transposed, flipped weights."""

    @staticmethod
    def forward(ctx, x, weight, bias, run, c):
        y = run(x)
        ctx.save_for_backward(x, weight)
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        c = ctx.c
        gy = gy.contiguous()
        dw, db = _wgrad_same(_as_c8(x), _pairs_to_c8(gy), c, 2, 3, 1)
        wt, zb = _bw_pack(weight, 'outT', lambda: (weight.detach().float().transpose(0, 1).flip(2, 3).contiguous(),         # (c, 2, 3, 3)
                                                   torch.zeros(c, device=gy.device)))
        B, F_, T, _ = gy.shape
        gx = torch.empty((B, F_, T, 4), dtype=torch.bfloat16, device=gy.device)
        _lib.check(_lib.lib().tt_conv_in(_p(gy), _p(gx), _p(wt), _p(zb), B, c, F_, T, 2, _s(gy)))
        return gx, dw, db, None, None


class _SkipAddFn(torch.autograd.Function):
    """Decoder skip connection (modules.py:568-589) with TimbreTrap.apply_skip_connections' learnable weight (:110-112):
    out = x + w * e on two bf16 tensors of one layout, one pass; gradients: x <- g, e <- w * g (one pass), w <- <g, e> (deterministic)."""

    @staticmethod
    def forward(ctx, x, e, w):
        scale = w.detach().float().reshape(1)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().tt_add_scaled_bf16(_p(x), _p(e), _p(scale), _p(out), x.numel(), _s(x)))
        ctx.save_for_backward(e, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        e, scale = ctx.saved_tensors
        g = g.contiguous()
        lib = _lib.lib()
        ge = torch.empty_like(g)
        _lib.check(lib.tt_add_scaled_bf16(_p(g), _p(g), _p(scale - 1.0), _p(ge), g.numel(), _s(g)))        # g + (w - 1) g = w g
        gw = torch.empty((), dtype=torch.float32, device=g.device)
        scratch = torch.empty(int(lib.tt_dot_scratch_floats()), dtype=torch.float32, device=g.device)
        _lib.check(lib.tt_dot_bf16(_p(g), _p(e), g.numel(), _p(gw), _p(scratch), _s(g)))
        return g, ge, gw


class _ActivationsFn(torch.autograd.Function):
    """TimbreTrap.to_activations (modules.py:271-289) on interleaved coefficients (B, F, T, 2) -> (B, F, T)."""

    @staticmethod
    def forward(ctx, coeffs):
        ctx.save_for_backward(coeffs)
        return CQT._magnitude(coeffs.permute(0, 3, 1, 2), True)

    @staticmethod
    def backward(ctx, gact):
        (coeffs,) = ctx.saved_tensors
        g = torch.empty_like(coeffs)
        _lib.check(_lib.lib().tt_activations_bwd(_p(coeffs), _p(gact.contiguous()), _p(g), gact.numel(), _s(coeffs)))
        return g


class _SqDiffFn(torch.autograd.Function):
    """compute_reconstruction_loss (objectives.py:11-33) on two interleaved coefficient tensors; gradients to both."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return compute_reconstruction_loss(a.permute(0, 3, 1, 2), b.permute(0, 3, 1, 2))

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        scale = 1.0 / (a.size(0) * a.size(2))
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.lib().tt_sum_sq_diff_bwd(_p(a), _p(b), _p(gout.contiguous()), scale, _p(ga), _p(gb), a.numel(), _s(a)))
        return ga, gb


class _TrnLossFn(torch.autograd.Function):
    """compute_transcription_loss(estimate, target, weight_positive_class=True) (objectives.py:36-74)."""

    @staticmethod
    def forward(ctx, est, tgt):
        ctx.save_for_backward(est, tgt)
        return compute_transcription_loss(est, tgt, True)

    @staticmethod
    def backward(ctx, gout):
        est, tgt = ctx.saved_tensors
        g = torch.empty_like(est)
        B, F, T = est.shape
        _lib.check(_lib.lib().tt_transcription_loss_bwd(_p(est), _p(tgt), _p(gout.contiguous()), B, F, T, 1, _p(g), _s(est)))
        return g, None


# ---------------------------------------------------------------------------------------------------------------
# the differentiable forward (same kernels, same order as Encoder / Decoder .forward_c8)
# ---------------------------------------------------------------------------------------------------------------
def _res(blk, x):
    c1, c2 = blk.conv1[0], blk.conv2[0]
    return _ResFn.apply(x, c1.weight, c1.bias, c2.weight, c2.bias, lambda t, mid: blk.forward_c8(t, mid_out=mid), blk.channels, blk.dilation)


def _encoder(enc, coeffs):
    """-> (latents C8, [the five embeddings in their internal layouts])"""
    ch = enc.channels
    w_in, b_in, w_lat, b_lat = enc._packed()
    ci = enc.convin[0]
    if enc.packed4:
        x = _InFn.apply(coeffs, ci.weight, ci.bias, lambda t: ops.conv_in(t, w_in, b_in, ch[0], packed4=True), ch[0])
    else:
        x = _ConvFn.apply(coeffs, ci.weight, ci.bias, lambda t: ops.conv_in(t, w_in, b_in, ch[0], packed4=enc.packed4),
                          _geom(3, 3, ph=1, pw=1), 2, ch[0], True)
    emb = [x]
    for i, blk in enumerate((enc.block1, enc.block2, enc.block3, enc.block4)):
        for rb in (blk.block1, blk.block2, blk.block3):
            x = _res(rb, x)
        sc = blk.sconv[0]

        def run_down(t, blk=blk):
            pack = P.pack_down_pairs if blk.packed4 else P.pack_down_strip
            (w,) = blk._cache.get((blk.sconv[0].weight, blk.sconv[0].bias), lambda: (pack(blk.sconv[0].weight, blk.sconv[0].bias),))
            return ops.conv_down_strip(t, w, P.pad8(blk.out_channels))
        x = _DownFn.apply(x, sc.weight, sc.bias, run_down, ch[i], ch[i + 1])
        emb.append(x)
    cl = enc.convlat
    if _lat_tc_ok(ch[4], enc.latent_pad):
        return _LatFn.apply(x, cl.weight, cl.bias, lambda t: ops.conv_lat(t, w_lat, b_lat, enc.latent_pad), ch[4], enc.latent_size, enc.latent_pad), emb
    return _ConvFn.apply(x, cl.weight, cl.bias, lambda t: ops.conv_lat(t, w_lat, b_lat, enc.latent_pad),
                         _geom(cl.weight.size(2), 1), ch[4], enc.latent_size, False), emb


class _IndicatorFn(torch.autograd.Function):
    """
    Decoder.convin on latents + the constant indicator channel (modules.py:139-142, 533-536).  Forward: tt_deconv_in with the
    indicator folded into the bias table.  Backward: transposed-conv gradients with the indicator channel written out (its weight
    row receives a gradient, its input does not).
    """

    @staticmethod
    def forward(ctx, lat, weight, bias, run, flag, d, c0):
        y = run(lat)
        ctx.save_for_backward(lat, y, weight)
        ctx.meta = (flag, d, c0)
        return y

    @staticmethod
    def backward(ctx, gy):
        lat, y, weight = ctx.saved_tensors
        flag, d, c0 = ctx.meta
        if _lat_tc_ok(c0, lat.size(1) * 8):
            # inference layouts end to end: ELU derivative, tap-grouped tensor-core weight gradient (rows 0 .. D-1; the indicator row and
            # the bias are row sums of dz), data gradient = the Encoder.convlat forward kernel on the first D weight rows
            dz = _ew('tt_elu_bwd_bf16', gy.contiguous(), y)
            h0 = y.size(2)
            dw = torch.zeros(weight.shape, dtype=torch.float32, device=lat.device)
            rows = torch.zeros((c0, h0), dtype=torch.float32, device=lat.device)
            _wgrad_lat(dz, lat, c0, d, dw, None, rows)
            dw[d, :, :, 0] = flag * rows
            db = rows.sum(dim=1)
            d_pad = lat.size(1) * 8
            wl, zl = _bw_pack(weight, 'decinT', lambda: (P.pack_lat(weight.detach().float()[:d], d_pad), torch.zeros(d_pad, device=lat.device)))
            glat = ops.conv_lat(dz, wl, zl, d_pad)
            return glat, dw, db, None, None, None, None
        geom = _geom(weight.size(2), 1)
        ln = _nchw(lat, d)                                                       # (B, D, 1, T)
        full = torch.cat((ln, torch.full_like(ln[:, :1], flag)), dim=1)          # (B, D+1, 1, T)
        dz = _elu_bwd(_nchw(gy, c0), _nchw(y, c0))
        w = weight.detach().float().contiguous()                                # (D+1, C0, H0, 1)
        dw, _ = _conv_bwd_weight(dz, full, w.shape, geom, False)
        db = _channel_sum(dz)
        gfull = _conv_fwd(dz, w, None, geom, False)                             # (B, D+1, 1, T)
        return _like(gfull[:, :d].contiguous(), lat), dw, db, None, None, None, None


def _decoder(dec, lat, reconstruct, skips=None):
    """skips: None or [(weight 0-dim tensor, embedding)] * 5 in encoder order (modules.py:568-589)"""
    ch = dec.channels
    w_in, tables, w_out, b_out = dec._packed()
    ci = dec.convin[0]
    sw = 1 if reconstruct else 0
    x = _IndicatorFn.apply(lat, ci.weight, ci.bias, lambda t: ops.deconv_in(t, w_in[sw], tables[sw], P.pad8(ch[0]), dec.embedding_size),
                           1.0 if reconstruct else 0.0, dec.latent_size, ch[0])
    for i, blk in enumerate((dec.block1, dec.block2, dec.block3, dec.block4)):
        if skips is not None:
            x = _SkipAddFn.apply(x, skips[-1 - i][1], skips[-1 - i][0])
        tc = blk.tconv[0]

        def run_up(t, blk=blk):
            (w,) = blk._cache.get((blk.tconv[0].weight, blk.tconv[0].bias), lambda: (P.pack_up_strip(blk.tconv[0].weight, blk.tconv[0].bias),))
            return ops.conv_up_strip(t, w, P.pad8(blk.out_channels), blk.out_pad, packed4_out=blk.packed4)
        x = _UpFn.apply(x, tc.weight, tc.bias, run_up, ch[i], ch[i + 1])
        for rb in (blk.block1, blk.block2, blk.block3):
            x = _res(rb, x)
    if skips is not None:
        x = _SkipAddFn.apply(x, skips[0][1], skips[0][0])
    co = dec.convout
    if dec.packed4:
        return _OutFn.apply(x, co.weight, co.bias, lambda t: ops.conv_out(t, w_out, b_out, ch[4]), ch[4])
    return _ConvFn.apply(x, co.weight, co.bias, lambda t: ops.conv_out(t, w_out, b_out, ch[4]), _geom(3, 3, ph=1, pw=1), ch[4], 2, False)


def allreduce_mean_gradients(params, group, flat=None):
    """Replicas with equal per-rank batches: the mean over ranks of the per-rank gradients is the gradient of the reference's
    global `.mean()` losses (objectives.py:31,72).  One flat bucket (base model: 614,490 fp32 = 2.46 MB), one all-reduce (NCCL on
    the GPU path; any backend works - tests/test_sharding_gloo.py runs it over gloo).  `flat`: the bucket the `.grad` tensors are
    views of (TrainStep) - reduced in place; without it the gradients are gathered into a bucket and copied back."""
    import torch.distributed as dist
    if flat is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()


class TrainStep:
    """
    One optimisation step of experiments/train.py:393-500 for TimbreTrap (with or without skip connections):
    losses as in compute_step_losses, backward, optional NCCL all-reduce of one flat gradient bucket (mean over ranks),
    clip_grad_norm_(max_norm) and AdamW (torch defaults: betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2; train.py:334).
    """

    def __init__(self, model, lr=1e-3, max_norm=10.0, multipliers=None, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, group=None):
        if getattr(model, 'HEAD_MODE', None) is not None or hasattr(model, 'film_layer'):
            raise NotImplementedError('TrainStep covers TimbreTrap with or without skip connections (experiments/train.py:155-161); '
                                      'the FiLM / magnitude variants run forward only')
        self.model = model
        self.params = [p for p in model.parameters()]
        self.lr, self.max_norm, self.betas, self.eps, self.wd = lr, max_norm, betas, eps, weight_decay
        self.mult = dict(reconstruction=1, transcription=1, consistency=1)
        self.mult.update(multipliers or {})
        self.group = group
        self.t = 0
        # ONE flat fp32 bucket each for parameters, gradients and the two AdamW moments: every parameter (and its .grad) is a view
        # into it, so autograd accumulates straight into the bucket, the NCCL all-reduce runs on it in place, and clip + AdamW are
        # two launches over 614,490 floats instead of 2 x 120 (train.py:493-496)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._grad_views = []
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + n].view(p.shape)
                self._grad_views.append(self.flat_g[off:off + n].view(p.shape))
                off += n
        _PackedCache.epoch += 1

    def losses(self, audio, ground_truth, late_start=False):
        """The four losses and their total, with the autograd graph attached."""
        model = self.model
        with torch.no_grad():
            coeffs = model.sliCQ.encode_interleaved(audio)
        def skips_of(emb):
            if model.skip_weights is None:
                return None
            return [(model.skip_weights[i], e) for i, e in enumerate(emb)]
        lat, emb = _encoder(model.encoder, coeffs)
        sk = skips_of(emb)
        rec = _decoder(model.decoder, lat, True, sk)
        trn = _decoder(model.decoder, lat, False, sk)
        act = _ActivationsFn.apply(trn)
        n = ground_truth.size(0)
        out = dict(reconstruction=_SqDiffFn.apply(rec, coeffs), transcription=_TrnLossFn.apply(act[:n].contiguous(), ground_truth.float().contiguous()))
        total = self.mult['reconstruction'] * out['reconstruction']
        if self.mult['consistency']:
            lat_t, emb_t = _encoder(model.encoder, trn)
            sk_t = skips_of(emb_t)
            trn_rec = _decoder(model.decoder, lat_t, True, sk_t)
            trn_scr = _decoder(model.decoder, lat_t, False, sk_t)
            tgt = trn[:n].contiguous()
            out['consistency_spectral'] = _SqDiffFn.apply(trn_rec[:n].contiguous(), tgt)
            out['consistency_score'] = _SqDiffFn.apply(trn_scr[:n].contiguous(), tgt)
        if not late_start:
            total = total + self.mult['transcription'] * out['transcription']
            if self.mult['consistency']:
                total = total + self.mult['consistency'] * (out['consistency_spectral'] + out['consistency_score'])
        out['total'] = total
        return out

    def backward(self, total):
        self.flat_g.zero_()
        for p, g in zip(self.params, self._grad_views):
            p.grad = g                                  # autograd accumulates in place: the gradients land in the flat bucket
        total.backward()
        if self.group is not None:
            self._sync_grad_views()
            allreduce_mean_gradients(self.params, self.group, flat=self.flat_g)

    def _sync_grad_views(self):
        """Gradients that were assigned from outside (p.grad = tensor) instead of accumulated by backward(): copy into the bucket."""
        for p, g in zip(self.params, self._grad_views):
            if p.grad is None:
                g.zero_()
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def optimizer_step(self):
        self.t += 1
        self._sync_grad_views()
        acc = torch.zeros((), dtype=torch.float64, device=self.flat_p.device)
        lib = _lib.lib()
        n = self.flat_p.numel()
        _lib.check(lib.tt_grad_sumsq(_p(self.flat_g), n, _p(acc), _s(self.flat_g)))
        _lib.check(lib.tt_adamw_step(_p(self.flat_p), _p(self.flat_g), _p(self.flat_m), _p(self.flat_v), n, _p(acc), self.max_norm, self.lr,
                                     self.betas[0], self.betas[1], self.eps, self.wd, self.t, _s(self.flat_p)))
        _PackedCache.epoch += 1                                                  # weights changed behind torch's back: repack lazily
        return acc.sqrt()

    def step(self, audio, ground_truth, late_start=False):
        """Returns the losses (detached 0-dim tensors) and the pre-clip gradient norm."""
        out = self.losses(audio, ground_truth, late_start)
        self.backward(out['total'])
        norm = self.optimizer_step()
        res = {k: v.detach() for k, v in out.items()}
        res['grad_norm'] = norm
        return res
