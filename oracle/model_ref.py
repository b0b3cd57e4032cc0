"""
CPU oracle for the Timbre-Trap autoencoder, the CQT wrapper conveniences, the chunked
inference loop and the three objectives - a functional, state_dict-driven restatement in
plain torch fp32 (torch.nn.functional library calls only).

TEST INFRASTRUCTURE ONLY (see oracle/nsgt_ref.py for the rule).  Pinned: every function
here is checked against the reference's own classes, imported unmodified from
/root/reference with `cqt_pytorch`/`librosa` stubbed, by scripts/make_golden.py; the
resulting vectors live in tests/golden/ and tests/test_oracle_golden.py replays them.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""

import math

import numpy as np
import torch
import torch.nn.functional as F

from .nsgt_ref import NSGTOracle

__all__ = ['CQTRef', 'encoder_ref', 'decoder_ref', 'decode_ref', 'inference_ref', 'chunked_inference_ref',
           'transcribe_ref', 'reconstruct_ref', 'forward_ref', 'reconstruction_loss_ref',
           'transcription_loss_ref', 'consistency_loss_ref', 'channel_plan', 'feature_sizes', 'init_state_dict',
           'to_decibels_ref', 'hann_sym', 'features_ref', 'decode_variant_ref', 'activations_variant_ref',
           'forward_variant_ref', 'chunked_inference_variant_ref']


# ---------------------------------------------------------------------------------------
# geometry of the conv stack
# ---------------------------------------------------------------------------------------

def channel_plan(model_complexity):
    """Encoder channel widths; the decoder uses them reversed (timbre_trap/framework/modules.py:417-424,503-510)."""
    return tuple(int(round(c * 2 ** (model_complexity - 1))) for c in (2, 4, 8, 16, 32))


def feature_sizes(n_bins):
    """Heights after each strided conv, h -> h//2 - 1, and the decoder's output paddings (modules.py:437-446,514-531)."""
    sizes, pads = [n_bins], []
    for _ in range(4):
        pads.append(sizes[-1] % 2)
        sizes.append(sizes[-1] // 2 - 1)
    return sizes, pads[::-1]


def init_state_dict(n_bins, latent_size=None, model_complexity=1, seed=0, variant='base'):
    """
    A deterministic, torch-version-independent random state_dict with the reference's 120
    names and shapes.  Values follow the fan-in-uniform recipe of
    torch's conv default init in spirit (U(-1/sqrt(fan_in), 1/sqrt(fan_in))), but are drawn
    from numpy's PCG64 so fixtures do not depend on torch's RNG stream.
    """
    ch = channel_plan(model_complexity)
    if latent_size is None:
        latent_size = 32 * 2 ** (model_complexity - 1)
    sizes, _ = feature_sizes(n_bins)
    rng = np.random.default_rng(seed)
    sd = {}

    def add(name, shape, fan_in, transposed=False):
        bound = 1.0 / math.sqrt(fan_in)
        n_out = shape[1] if transposed else shape[0]
        sd[name + '.weight'] = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))
        sd[name + '.bias'] = torch.from_numpy(rng.uniform(-bound, bound, size=(n_out,)).astype(np.float32))

    def add_res(prefix, c):
        add(prefix + '.conv1.0', (c, c, 3, 3), c * 9)
        add(prefix + '.conv2.0', (c, c, 1, 1), c)

    # variant: 'base' | 'film' (modules.py:780-840) | 'mag' / 'magdb' (modules.py:892-1075: one-channel convin / convout)
    cin = 1 if variant in ('mag', 'magdb') else 2
    add('encoder.convin.0', (ch[0], cin, 3, 3), cin * 9)
    for i in range(4):
        for j in range(3):
            add_res(f'encoder.block{i + 1}.block{j + 1}', ch[i])
        add(f'encoder.block{i + 1}.sconv.0', (ch[i + 1], ch[i], 4, 1), ch[i] * 4)
    add('encoder.convlat', (latent_size, ch[4], sizes[4], 1), ch[4] * sizes[4])

    dch = ch[::-1]
    # ConvTranspose2d weights are (C_in, C_out, kh, kw); torch derives fan_in from dim 1
    add('decoder.convin.0', (latent_size + (0 if variant == 'film' else 1), dch[0], sizes[4], 1), dch[0] * sizes[4], True)
    for i in range(4):
        add(f'decoder.block{i + 1}.tconv.0', (dch[i], dch[i + 1], 4, 1), dch[i + 1] * 4, True)
        for j in range(3):
            add_res(f'decoder.block{i + 1}.block{j + 1}', dch[i + 1])
    add('decoder.convout', (cin, dch[4], 3, 3), dch[4] * 9)
    if variant == 'film':
        add('film_layer.gamma', (latent_size, 2), 2)
        add('film_layer.beta', (latent_size, 2), 2)
    return sd


# ---------------------------------------------------------------------------------------
# conv stack
# ---------------------------------------------------------------------------------------

def _res(x, sd, p, d):
    """ResidualConv2dBlock.forward (modules.py:743-777): 3x3 dilated 'same' conv, ELU, 1x1 conv, ELU, + x."""
    y = F.elu(F.conv2d(x, sd[p + '.conv1.0.weight'], sd[p + '.conv1.0.bias'], padding=d, dilation=d))
    y = F.elu(F.conv2d(y, sd[p + '.conv2.0.weight'], sd[p + '.conv2.0.bias']))
    return y + x


def _enc_block(x, sd, p):
    """EncoderBlock.forward (modules.py:632-655): dilations 1,2,3 then (4,1)/(2,1) strided conv + ELU."""
    for j, d in enumerate((1, 2, 3)):
        x = _res(x, sd, f'{p}.block{j + 1}', d)
    return F.elu(F.conv2d(x, sd[p + '.sconv.0.weight'], sd[p + '.sconv.0.bias'], stride=(2, 1)))


def _dec_block(x, sd, p, out_pad):
    """DecoderBlock.forward (modules.py:695-718): (4,1)/(2,1) transposed conv (+output_padding) + ELU, then dilations 1,2,3."""
    x = F.elu(F.conv_transpose2d(x, sd[p + '.tconv.0.weight'], sd[p + '.tconv.0.bias'], stride=(2, 1),
                                 output_padding=(out_pad, 0)))
    for j, d in enumerate((1, 2, 3)):
        x = _res(x, sd, f'{p}.block{j + 1}', d)
    return x


def encoder_ref(coefficients, sd):
    """Encoder.forward (modules.py:448-483) -> (latents (B,D,T), [5 embeddings])."""
    emb = [F.elu(F.conv2d(coefficients, sd['encoder.convin.0.weight'], sd['encoder.convin.0.bias'], padding=1))]
    for i in range(4):
        emb.append(_enc_block(emb[-1], sd, f'encoder.block{i + 1}'))
    latents = F.conv2d(emb[-1], sd['encoder.convlat.weight'], sd['encoder.convlat.bias']).squeeze(-2)
    return latents, emb


def decoder_ref(latents_plus, sd, n_bins, skips=None):
    """Decoder.forward (modules.py:545-594); `latents_plus` already carries the indicator channel."""
    _, pads = feature_sizes(n_bins)
    x = F.elu(F.conv_transpose2d(latents_plus.unsqueeze(-2), sd['decoder.convin.0.weight'], sd['decoder.convin.0.bias']))
    for i in range(4):
        if skips is not None:
            x = x + skips[-1 - i]
        x = _dec_block(x, sd, f'decoder.block{i + 1}', pads[i])
    if skips is not None:
        x = x + skips[0]
    return F.conv2d(x, sd['decoder.convout.weight'], sd['decoder.convout.bias'], padding=1)


def decode_ref(latents, sd, n_bins, transcribe=False, skips=None):
    """TimbreTrap.decode (modules.py:119-147): indicator channel 1 = reconstruct, 0 = transcribe, appended last."""
    flag = torch.full_like(latents[..., :1, :], 0.0 if transcribe else 1.0)
    return decoder_ref(torch.cat((latents, flag), dim=-2), sd, n_bins, skips)


def _skips(sd, emb):
    """TimbreTrap.apply_skip_connections (modules.py:95-117)."""
    if 'skip_weights' not in sd:
        return None
    return [sd['skip_weights'][i] * e for i, e in enumerate(emb)]


# ---------------------------------------------------------------------------------------
# CQT wrapper (cqtwrapper.py) around the NSGT oracle
# ---------------------------------------------------------------------------------------

def to_decibels_ref(magnitude, rescale=True):
    """
    CQT.to_decibels (cqtwrapper.py:143-182) with torchaudio's AmplitudeToDB(stype='amplitude',
    top_db=80) written out: 20*log10(clamp(x, 1e-10)) (multiplier 20, amin 1e-10, ref 1), floored
    per item at max-80; then ceiling moved to 0 dB and mapped to [0, 1].
    """
    out = []
    for m in magnitude:
        d = 20.0 * torch.log10(torch.clamp(m, min=1e-10))
        d = torch.maximum(d, d.max() - 80.0)
        if rescale:
            d = 1 + (d - d.max()) / 80
        out.append(d.unsqueeze(0))
    return torch.cat(out, dim=0)


class CQTRef:
    """Restates timbre_trap.framework.CQT (cqtwrapper.py:10-308) on top of NSGTOracle."""

    def __init__(self, n_octaves, bins_per_octave, sample_rate, secs_per_block, dtype=np.complex128):
        self.nsgt = NSGTOracle(n_octaves, bins_per_octave, sample_rate, int(secs_per_block * sample_rate), True, dtype)
        self.block_length = self.nsgt.block_length
        self.max_window_length = self.nsgt.max_window_length
        self.sample_rate = sample_rate
        self.hop_length = self.block_length / self.max_window_length            # cqtwrapper.py:40
        self.n_bins = n_octaves * bins_per_octave                               # :43
        fmin = 12.0 * (math.log2((sample_rate / 2) / (2 ** n_octaves)) - math.log2(440.0)) + 69.0   # librosa.hz_to_midi, :45
        self.midi_freqs = fmin + np.arange(self.n_bins) / (bins_per_octave / 12)   # :48

    def encode(self, audio):
        return torch.from_numpy(self.nsgt.encode(audio.detach().cpu().numpy()).astype(np.complex64))

    def __call__(self, audio):
        """CQT.forward (cqtwrapper.py:50-72)."""
        return self.to_real(self.encode(audio))

    @staticmethod
    def to_real(c):
        """cqtwrapper.py:74-97 - (B,1,F,T) complex -> (B,2,F,T) real view."""
        return torch.view_as_real(c.squeeze(-3)).permute(0, 3, 1, 2)

    @staticmethod
    def to_complex(c):
        """cqtwrapper.py:99-120."""
        return torch.view_as_complex(c.permute(0, 2, 3, 1).contiguous())

    @staticmethod
    def to_magnitude(c):
        """cqtwrapper.py:122-141."""
        return torch.sqrt((c ** 2).sum(dim=-3))

    def decode(self, coefficients):
        """CQT.decode (cqtwrapper.py:184-213): synthesis, then global infinity-norm normalise (skipped if the peak is 0)."""
        if not coefficients.is_complex():
            coefficients = self.to_complex(coefficients).unsqueeze(-3)
        audio = torch.from_numpy(self.nsgt.decode(coefficients.detach().cpu().numpy()).astype(np.float32))
        peak = audio.abs().max()
        if peak:
            audio = audio / peak
        return audio

    def decode_raw(self, coefficients):
        """Synthesis without the peak normalise (exposes the transform's own scale)."""
        if not coefficients.is_complex():
            coefficients = self.to_complex(coefficients).unsqueeze(-3)
        return torch.from_numpy(self.nsgt.decode(coefficients.detach().cpu().numpy()).astype(np.float32))

    def pad_to_block_length(self, audio):
        """cqtwrapper.py:215-233."""
        return F.pad(audio, (0, -audio.size(-1) % self.block_length))

    def get_expected_samples(self, t):
        """cqtwrapper.py:235-253."""
        return int(max(0, t) * self.sample_rate)

    def get_expected_frames(self, num_samples):
        """cqtwrapper.py:255-273."""
        return math.ceil((num_samples / self.block_length) * self.max_window_length)

    def get_times(self, n_frames):
        """cqtwrapper.py:275-293."""
        return np.arange(n_frames) * self.hop_length / self.sample_rate


# ---------------------------------------------------------------------------------------
# inference paths (modules.py:149-336)
# ---------------------------------------------------------------------------------------

def hann_sym(n):
    """torch.signal.windows.hann(n) (symmetric): 0.5 - 0.5 cos(2 pi i / (n-1)) (modules.py:239)."""
    return torch.from_numpy((0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))).astype(np.float32))


def inference_ref(audio, sd, cqt, transcribe=False):
    """TimbreTrap._inference (modules.py:149-177) on audio that is already a whole number of blocks."""
    latents, emb = encoder_ref(cqt(audio), sd)
    return decode_ref(latents, sd, cqt.n_bins, transcribe, _skips(sd, emb))


def chunked_inference_ref(audio, sd, cqt, transcribe=False, prepadded=False):
    """TimbreTrap.chunked_inference (modules.py:204-269): 50 % overlapped blocks, Hann cross-fade of coefficients, trim.
    prepadded: `audio` already is whole blocks plus half a block on either side (a shard with its neighbourhood)."""
    B, nb = audio.size(0), cqt.n_bins
    hop = cqt.block_length // 2
    if not prepadded:
        audio = cqt.pad_to_block_length(audio)
        audio = F.pad(audio, [hop, hop])
    n_chunks = (audio.size(-1) - hop) // hop
    M = cqt.max_window_length
    window = hann_sym(M)
    out = torch.zeros((B, 2, nb, cqt.get_expected_frames(audio.size(-1))))
    for i in range(n_chunks):
        piece = audio[..., i * hop: i * hop + cqt.block_length]
        f0 = i * M // 2
        out[..., f0: f0 + M] += window * inference_ref(piece, sd, cqt, transcribe)
    return out[..., M // 2: -M // 2]


def transcribe_ref(audio, sd, cqt):
    """TimbreTrap.transcribe (modules.py:292-313): tanh of the magnitude of the cross-faded coefficients."""
    return torch.tanh(cqt.to_magnitude(chunked_inference_ref(audio, sd, cqt, True)))


def reconstruct_ref(audio, sd, cqt):
    """TimbreTrap.reconstruct (modules.py:315-336)."""
    return cqt.decode(chunked_inference_ref(audio, sd, cqt, False))


def forward_ref(audio, sd, cqt, consistency=False):
    """TimbreTrap.forward (modules.py:338-393) -> (reconstruction, latents, transcription, transcription_rec, transcription_scr)."""
    latents, emb = encoder_ref(cqt(audio), sd)
    sk = _skips(sd, emb)
    rec = decode_ref(latents, sd, cqt.n_bins, False, sk)
    trn = decode_ref(latents, sd, cqt.n_bins, True, sk)
    trn_rec = trn_scr = None
    if consistency:
        lat2, emb2 = encoder_ref(trn, sd)
        sk2 = _skips(sd, emb2)
        trn_rec = decode_ref(lat2, sd, cqt.n_bins, False, sk2)
        trn_scr = decode_ref(lat2, sd, cqt.n_bins, True, sk2)
    return rec, latents, trn, trn_rec, trn_scr


# ---------------------------------------------------------------------------------------
# model variants (modules.py:780-1075)
# ---------------------------------------------------------------------------------------

def features_ref(variant, audio, cqt):
    """The encoder's input: TimbreTrap.encode :88, TimbreTrapMag.encode :947, TimbreTrapMagDB.encode :1024-1027."""
    c = cqt(audio)
    if variant in ('base', 'film'):
        return c
    mag = cqt.to_magnitude(c)
    return (to_decibels_ref(mag) if variant == 'magdb' else mag).unsqueeze(-3)


def decode_variant_ref(variant, latents, sd, n_bins, transcribe=False, skips=None):
    """TimbreTrap.decode :119-147 / TimbreTrapFiLM.decode :811-840 / TimbreTrapMag.decode :952-978 / TimbreTrapMagDB.decode :1033-1054."""
    if variant == 'film':
        cond = torch.tensor([float(transcribe), float(not transcribe)])
        gamma = F.linear(cond, sd['film_layer.gamma.weight'], sd['film_layer.gamma.bias'])
        beta = F.linear(cond, sd['film_layer.beta.weight'], sd['film_layer.beta.bias'])
        return decoder_ref((latents.transpose(-1, -2) * gamma + beta).transpose(-1, -2), sd, n_bins, skips)
    out = decode_ref(latents, sd, n_bins, transcribe, skips)
    if variant == 'mag':
        return F.relu(out)
    if variant == 'magdb':
        return torch.sigmoid(out)
    return out


def activations_variant_ref(variant, coefficients):
    """to_activations: TimbreTrap :271-289, TimbreTrapMag :980-1001, TimbreTrapMagDB :1056-1075."""
    if variant == 'mag':
        return torch.tanh(coefficients.squeeze(-3))
    if variant == 'magdb':
        return coefficients.squeeze(-3)
    return torch.tanh(CQTRef.to_magnitude(coefficients))


def forward_variant_ref(variant, audio, sd, cqt, consistency=False):
    """TimbreTrap.forward :338-393 with the variant's encode / decode."""
    latents, emb = encoder_ref(features_ref(variant, audio, cqt), sd)
    sk = _skips(sd, emb)
    rec = decode_variant_ref(variant, latents, sd, cqt.n_bins, False, sk)
    trn = decode_variant_ref(variant, latents, sd, cqt.n_bins, True, sk)
    trn_rec = trn_scr = None
    if consistency:
        lat2, emb2 = encoder_ref(trn, sd)
        sk2 = _skips(sd, emb2)
        trn_rec = decode_variant_ref(variant, lat2, sd, cqt.n_bins, False, sk2)
        trn_scr = decode_variant_ref(variant, lat2, sd, cqt.n_bins, True, sk2)
    return rec, latents, trn, trn_rec, trn_scr


def chunked_inference_variant_ref(variant, audio, sd, cqt, transcribe=False):
    """The inherited TimbreTrap.chunked_inference (modules.py:204-269) as a variant runs it: the buffer has two channels whatever
    the variant's output has (:244), a one-channel chunk output broadcasts into both (:259-263)."""
    B, nb = audio.size(0), cqt.n_bins
    hop = cqt.block_length // 2
    audio = F.pad(cqt.pad_to_block_length(audio), [hop, hop])
    n_chunks = (audio.size(-1) - hop) // hop
    M = cqt.max_window_length
    window = hann_sym(M)
    out = torch.zeros((B, 2, nb, cqt.get_expected_frames(audio.size(-1))))
    for i in range(n_chunks):
        piece = audio[..., i * hop: i * hop + cqt.block_length]
        latents, emb = encoder_ref(features_ref(variant, piece, cqt), sd)
        out[..., i * M // 2: i * M // 2 + M] += window * decode_variant_ref(variant, latents, sd, nb, transcribe, _skips(sd, emb))
    return out[..., M // 2: -M // 2]


# ---------------------------------------------------------------------------------------
# objectives (objectives.py)
# ---------------------------------------------------------------------------------------

def reconstruction_loss_ref(reconstructed, target):
    """objectives.py:11-33 - squared error summed over (C,F), averaged over (B,T)."""
    return ((reconstructed - target) ** 2).sum(-3).sum(-2).mean()


def transcription_loss_ref(estimate, target, weight_positive_class=False):
    """objectives.py:36-74 - per frame, bins with target == 1 are up-weighted by (#neg mass)/(#pos mass + eps)."""
    err = (estimate - target) ** 2
    if weight_positive_class:
        pos = target.sum(dim=-2, keepdim=True)
        neg = (1 - target).sum(dim=-2, keepdim=True)
        scale = (neg / (pos + torch.finfo(torch.float32).eps)) * (target == 1)
        scale = torch.where(scale == 0, torch.ones_like(scale), scale)
        err = err * scale
    return err.sum(-2).mean()


def consistency_loss_ref(spectral, transcription, target):
    """objectives.py:77-104."""
    return reconstruction_loss_ref(spectral, target), reconstruction_loss_ref(transcription, target)
