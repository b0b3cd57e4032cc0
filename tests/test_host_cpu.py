"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the product's filter-bank
tables equal the oracle's, module/state_dict structure matches the reference, chunk slicing matches the reference loop."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from timbre_trap_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()                                   # a fresh checkout: the .so is not in the history (nvcc cross-compiles without a GPU)
    header = open(os.path.join(ROOT, 'include', 'timbre_trap_b200.h')).read()
    declared = set(re.findall(r'\b(tt_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.lib().tt_version() == 1


@pytest.mark.parametrize('cfg', [(9, 60, 22050, 3), (6, 12, 8000, 0.5), (8, 24, 16000, 1.0), (7, 36, 44100, 1.5)])
def test_filter_bank_matches_oracle(cfg):
    from oracle.nsgt_ref import make_tables
    from timbre_trap_b200.nsgt_tables import FilterBank
    a, b = FilterBank(cfg[0], cfg[1], cfg[2], int(cfg[3] * cfg[2])), make_tables(*cfg)
    assert a.max_window_length == b.max_window_length and a.n_taps == b.n_taps
    for k in ('start', 'length', 'first', 'offset'):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.abs(a.win - b.win_packed).max() < 1e-7
    assert np.abs(a.dual - b.dual_packed).max() <= 1e-7 * b.dual_packed.max()


def test_geometry_and_state_dict_match_reference(golden_dir):
    from oracle import model_ref as R
    from timbre_trap_b200.framework import CQT, TimbreTrap
    geo = json.load(open(os.path.join(golden_dir, 'geometry.json')))
    for g in geo.values():
        cfg = g['cfg']
        c = CQT(cfg['n_octaves'], cfg['bins_per_octave'], cfg['sample_rate'], cfg['secs_per_block'])
        assert (c.block_length, c.max_window_length, c.n_bins, c.hop_length) == (g['block_length'], g['max_window_length'], g['n_bins'], g['hop_length'])
        assert abs(c.midi_freqs[0] - g['midi_first']) < 1e-9 and abs(c.get_midi_freqs()[-1] - g['midi_last']) < 1e-9
        for p, v in g['expected_frames'].items():
            assert c.get_expected_frames(int(p)) == v
        for t, v in g['expected_samples'].items():
            assert c.get_expected_samples(float(t)) == v
        for p, v in g['padded_len'].items():
            assert c.pad_to_block_length(torch.zeros(1, 1, int(p))).size(-1) == v
        np.testing.assert_allclose(c.get_times(6), g['times_head'], rtol=1e-12)
    m = TimbreTrap(22050, 9, 60, 3, latent_size=128, model_complexity=2)
    sd = R.init_state_dict(540, 128, 2, 0)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert sum(p.numel() for p in m.parameters()) == 614490
    # reference checkpoints carry cqt_pytorch's buffers under sliCQ.*: accepted and dropped
    sd2 = dict(sd)
    sd2['sliCQ.windows'] = torch.zeros(540, 1024)
    sd2['sliCQ.windows_range_indices'] = torch.zeros(540, 1024, dtype=torch.long)
    m.load_state_dict(sd2)
    ms = TimbreTrap(22050, 9, 60, 3, skip_connections=True)
    assert 'skip_weights' in ms.state_dict() and ms.sliCQ.n_bins == 540


def test_models_with_live_plans_pickle_and_deepcopy():
    """The reference checkpoints WHOLE modules (experiments/train.py:511 torch.save(model, path); every script torch.load()s
    them): a model that has already run on a device (= holds a plan with raw device handles) must still pickle / deep-copy."""
    import copy
    import io
    from timbre_trap_b200.framework import TimbreTrap

    class FakePlan:                                        # what CQT._plan() caches after the first GPU call
        handle = ctypes.c_void_p(0x1234)
        device = 'cuda:0'

    m = TimbreTrap(8000, 6, 12, 0.5, latent_size=16, model_complexity=1, skip_connections=True)
    m.sliCQ._plans[('cuda', 0)] = FakePlan()
    m._windows[('cuda', 0)] = torch.ones(4)
    m.encoder.block1.block1._packed('planar')              # a live packed-weight cache entry
    with pytest.raises(Exception):
        import pickle
        pickle.dumps(FakePlan.handle)                      # the thing that used to break torch.save(model)
    c = copy.deepcopy(m)
    assert c.sliCQ._plans == {} and c._windows == {} and m.sliCQ._plans          # the copy starts without plans, the original keeps its own
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert r.sliCQ._plans == {} and r.sliCQ.block_length == m.sliCQ.block_length
    assert np.array_equal(r.sliCQ._bank.win, m.sliCQ._bank.win)
    for (k, a), (_, b) in zip(m.state_dict().items(), r.state_dict().items()):
        assert torch.equal(a, b), k
    m.sliCQ._plans.clear()


def test_filter_bank_from_checkpoint_buffers():
    """CQT.load_state_dict drives the kernels with the tables a reference checkpoint carries (sliCQ.windows /
    windows_range_indices / windows_inverse) - here emitted by the oracle in the dense upstream shape, in two variants."""
    import oracle.nsgt_ref as N
    from timbre_trap_b200.framework import CQT
    from timbre_trap_b200.nsgt_tables import FilterBank
    cfg = (6, 12, 8000, 0.5)
    for periodic in (True, False):
        old = N.U4_HANN_PERIODIC
        N.U4_HANN_PERIODIC = periodic
        try:
            t = N.make_tables(*cfg)
        finally:
            N.U4_HANN_PERIODIC = old
        for inverse in (t.win_inv, np.where(t.frame_diagonal > 0, 1.0 / np.where(t.frame_diagonal > 0, t.frame_diagonal, 1.0), 0.0)):
            c = CQT(*cfg)
            c.load_state_dict({'windows': torch.from_numpy(t.win).float(), 'windows_range_indices': torch.from_numpy(t.idx),
                               'windows_inverse': torch.from_numpy(inverse).float()})
            b = c._bank
            assert b.source == 'checkpoint buffers' and list(c.state_dict()) == []
            # expand the packed tables again: they must reproduce the dense ones
            for k in range(0, t.n_bins, 5):
                dense_w, dense_d = np.zeros(t.max_window_length), np.zeros(t.max_window_length)
                sl = slice(b.first[k], b.first[k] + b.length[k])
                dense_w[sl] = b.win[b.offset[k]:b.offset[k + 1]]
                dense_d[sl] = b.dual[b.offset[k]:b.offset[k + 1]]
                assert np.abs(dense_w - t.win[k]).max() < 1e-7 and np.abs(dense_d - t.win_inv[k]).max() <= 1e-6 * t.win_inv.max()
                assert b.start[k] == t.idx[k, b.first[k]]
    # incomplete / placeholder buffers keep the constructed tables; a wrong geometry is an error, not a silent fallback
    c = CQT(*cfg)
    c.load_state_dict({'windows': torch.zeros(72, 256)})
    assert c._bank.source.startswith('constructor')
    with pytest.raises(ValueError):
        c.load_state_dict({'windows': torch.ones(72, 128), 'windows_range_indices': torch.zeros(72, 128, dtype=torch.long),
                           'windows_inverse': torch.ones(72, 128)})
    with pytest.raises(ValueError):
        FilterBank.from_buffers(4000, np.ones((2, 8)), np.stack([np.arange(8) * 2, np.arange(8)]), np.ones((2, 8)))   # taps not consecutive


def test_unsupported_channel_plans_fail_at_construction():
    from timbre_trap_b200.framework import Decoder, Encoder, TimbreTrap
    for cls in (Encoder, Decoder):
        with pytest.raises(ValueError, match='model_complexity 1 and 2'):
            cls(feature_size=540, latent_size=None, model_complexity=3)
    with pytest.raises(ValueError):
        TimbreTrap(22050, 9, 60, 3, model_complexity=3)
    TimbreTrap(22050, 9, 60, 3, model_complexity=1)
    TimbreTrap(22050, 9, 60, 3, latent_size=128, model_complexity=2)
    # latent sizes: padded to the GEMM widths the (H, 1)-kernel layers are built for, the state_dict keeps the reference's shapes
    for latent, pad in ((1, 16), (16, 16), (24, 32), (40, 64), (100, 128), (129, 256), (256, 256)):
        m = TimbreTrap(8000, 6, 12, 0.5, latent_size=latent)
        assert m.encoder.latent_pad == m.decoder.latent_pad == pad and m.encoder.convlat.weight.size(0) == latent
    with pytest.raises(ValueError, match='latent_size'):
        TimbreTrap(8000, 6, 12, 0.5, latent_size=257)


def test_cpu_tensors_are_rejected_loudly():
    from timbre_trap_b200._lib import TimbreTrapB200Error
    from timbre_trap_b200.framework import TimbreTrap
    m = TimbreTrap(8000, 6, 12, 0.5)
    with pytest.raises(TimbreTrapB200Error):
        m.transcribe(torch.zeros(1, 1, 4000))
    with pytest.raises(TimbreTrapB200Error):
        m.sliCQ(torch.zeros(1, 1, 4000))


def test_chunk_slicing_matches_reference_loop():
    """TimbreTrap._chunks must produce exactly the slices of modules.py:226-253."""
    from timbre_trap_b200.framework import TimbreTrap
    m = TimbreTrap(8000, 6, 12, 0.5)
    L = m.sliCQ.block_length
    audio = torch.arange(2 * int(2.3 * L), dtype=torch.float32).reshape(2, 1, -1)
    chunks, n_chunks = m._chunks(audio)
    padded = torch.nn.functional.pad(m.sliCQ.pad_to_block_length(audio), [L // 2] * 2)
    assert n_chunks == (padded.size(-1) - L // 2) // (L // 2) == 7
    for b in range(2):
        for i in range(n_chunks):
            assert torch.equal(chunks[b * n_chunks + i, 0], padded[b, 0, i * (L // 2): i * (L // 2) + L])


def _emulate_rs(x_rows, w1p, bias, d, shifts, step, halo, NC):
    """The GEMMs csrc/res_rs.cu issues for the 3x3 stage, in torch: x_rows (H, Tr, K) GEMM rows, w1p (KG, 3 NC, 8) packed weights with
    K groups ordered (shift, 8-wide chunk); input row r, shift s reads rows t + s * step - halo and feeds output rows r-d, r, r+d."""
    H, Tr, K = x_rows.shape
    kc = K // 8
    w = w1p.float().reshape(shifts, -1, 3, NC, 8)[:, :kc]                         # [shift][chunk][j][n][k]
    xp = torch.nn.functional.pad(x_rows, (0, 0, halo, halo))
    acc = bias[0].float().view(1, 1, NC).repeat(H, Tr, 1)
    for r in range(H):
        for s_ in range(shifts):
            a = xp[r, s_ * step: s_ * step + Tr].reshape(Tr, kc, 8)
            for j in range(3):
                o = r + (j - 1) * d
                if 0 <= o < H:
                    acc[o] += torch.einsum('tck,cnk->tn', a, w[s_, :, j])
    return acc


def _emulate_rs_fold(x_rows, w1p, bias, d, mmas, NC):
    """The folded-row schedule of csrc/res_rs.cu (RsPlan::fold_s / fold_adj): MMA (s, adj) multiplies the first 16-byte half of GEMM
    row t + s and the second half of row t + s - adj with the K groups (mma, half) of w1p."""
    H, Tr, K = x_rows.shape
    assert K == 16
    halo = 1 + max(abs(s_) + a for s_, a in mmas)
    w = w1p.float().reshape(len(mmas), 2, 3, NC, 8)                              # [mma][half][j][n][k]
    xp = torch.nn.functional.pad(x_rows, (0, 0, halo, halo))
    acc = bias[0].float().view(1, 1, NC).repeat(H, Tr, 1)
    for r in range(H):
        for m, (s_, adj) in enumerate(mmas):
            first = xp[r, halo + s_: halo + s_ + Tr, :8]
            second = xp[r, halo + s_ - adj: halo + s_ - adj + Tr, 8:]
            for j in range(3):
                o = r + (j - 1) * d
                if 0 <= o < H:
                    acc[o] += first @ w[m, 0, j].t() + second @ w[m, 1, j].t()
    return acc


def test_weight_packing_row_stationary():
    """packing.pack_res_rs / _pairs / _fold against F.conv2d through a torch emulation of the kernel's GEMM schedule."""
    import torch.nn.functional as F
    from timbre_trap_b200.framework import packing as P
    torch.manual_seed(0)
    H, T = 7, 24
    for C, d, mode in ((16, 2, 'planar'), (32, 1, 'planar'), (4, 1, 'fold4'), (2, 2, 'fold4'), (3, 3, 'fold4'), (7, 1, 'fold2'), (8, 2, 'fold2'),
                       (5, 3, 'fold2'), (4, 3, 'pairs'), (8, 1, 'planar8')):
        x = torch.randn(1, C, H, T).to(torch.bfloat16).float()
        w1, b1 = torch.randn(C, C, 3, 3).to(torch.bfloat16).float(), torch.randn(C)
        w2, b2 = torch.randn(C, C, 1, 1).to(torch.bfloat16).float(), torch.randn(C)
        want = F.conv2d(x, w1, b1, padding=d, dilation=d)[0]                   # (C, H, T)
        if mode.startswith('fold'):
            fold = int(mode[-1]); Cw = 16 // fold
            w1p, w2p, bias = P.pack_res_rs_fold(w1, b1, w2, b2, d, fold)
            mmas = P.res_rs_fold_mmas(d, fold)
            # two MMAs per input row where the taps' frames fit four 16-byte halves, three otherwise (never the 5 whole-row shifts of d = 3)
            assert len(mmas) == (2 if (fold == 4 and d <= 2) or (fold == 2 and d == 1) else 3)
            assert tuple(w1p.shape) == (2 * len(mmas), 48, 8) and tuple(w2p.shape) == (2, 16, 8) and tuple(bias.shape) == (2, 16)
            rows = torch.zeros(H, T // fold, fold, Cw)
            rows[..., :C] = x[0].permute(1, 2, 0).reshape(H, T // fold, fold, C)
            acc = _emulate_rs_fold(rows.reshape(H, T // fold, 16), w1p, bias, d, mmas, 16)
            got = acc.reshape(H, T // fold, fold, Cw)[..., :C].reshape(H, T, C).permute(2, 0, 1)
        elif mode == 'pairs':
            w1p, w2p, bias = P.pack_res_rs_pairs(w1, b1, w2, b2, d)
            hp = (d + 1) // 2
            assert tuple(w1p.shape) == (2 * hp + 2, 48, 8)
            rows = torch.zeros(H, T // 2, 2, 4)
            rows[..., :C] = x[0].permute(1, 2, 0).reshape(H, T // 2, 2, C)
            # K groups are single 8-wide chunks per shift here (the kernel pairs neighbouring shifts into one K = 16 MMA)
            acc = _emulate_rs(rows.reshape(H, T // 2, 8), w1p[:2 * hp + 1], bias, d, 2 * hp + 1, 1, hp, 16)
            got = acc.reshape(H, T // 2, 16)[..., :8].reshape(H, T // 2, 2, 4)[..., :C].reshape(H, T, C).permute(2, 0, 1)
        else:
            w1p, w2p, bias = P.pack_res_rs(w1, b1, w2, b2)
            Cp, NC = P.pad8(C), (32 if C >= 32 else 16)
            assert tuple(w1p.shape) == (3 * (Cp // 8) + (1 if Cp == 8 else 0), 3 * NC, 8) and w1p.dtype == torch.bfloat16
            rows = torch.zeros(H, T, Cp)
            rows[..., :C] = x[0].permute(1, 2, 0)
            acc = _emulate_rs(rows, w1p[:3 * (Cp // 8)], bias, d, 3, d, d, NC)
            got = acc[..., :C].permute(2, 0, 1)
        assert float((got - want).abs().max()) <= 2e-2 * float(want.abs().max()), (C, d, mode)
        # 1x1 stage: block-diagonal over the frames of a folded row, biases in the accumulator-column order
        n2 = w2p.float().permute(1, 0, 2).reshape(w2p.shape[1], -1)              # [n][k over the row]
        Cw = {'fold4': 4, 'fold2': 8, 'pairs': 4}.get(mode, None)
        if Cw:
            assert torch.allclose(n2[:C, :C], w2.reshape(C, C)) and torch.allclose(n2[Cw:Cw + C, Cw:Cw + C], w2.reshape(C, C))
            assert float(n2[:Cw, Cw:2 * Cw].abs().max()) == 0.0 and torch.allclose(bias[1, Cw:Cw + C], b2)
        else:
            assert torch.allclose(n2[:C, :C], w2.reshape(C, C)) and torch.allclose(bias[1, :C], b2)


def test_weight_packing_shapes():
    from timbre_trap_b200.framework import packing as P
    w, b = torch.randn(129, 64, 31, 1), torch.randn(64)
    packed, tables = P.pack_deconv_in(w, b, 128)
    assert tuple(packed.shape) == (31, 16, 64, 8) and tuple(tables.shape) == (2, 31, 64)
    assert torch.allclose(tables[1] - tables[0], w[128, :, :, 0].t(), atol=1e-6)
    x = torch.randn(2, 5, 7, 16)
    assert torch.equal(P.from_c8(P.to_c8(x), 5), x.to(torch.bfloat16).float())


def _unpack_b(packed):
    """(KG, N, 8) bf16 B operand -> (N, K) fp32."""
    return packed.float().permute(1, 0, 2).reshape(packed.shape[1], -1)


def _ones_k(n_rows):
    """The kernels' constant "ones" K group: (1, 1, 0, ..., 0) per GEMM row - meets the bias group (bf16 hi + lo)."""
    o = torch.zeros(n_rows, 8)
    o[:, :2] = 1.0
    return o


def test_weight_packing_strided_transposed_latent():
    """pack_down / pack_up (+ _strip, _pairs variants with their bias groups), pack_lat, pack_deconv_in against torch's convs,
    through the GEMMs the kernels issue (K order, polyphase rows, bias-as-K-group, indicator-as-bias-table)."""
    import torch.nn.functional as F
    from timbre_trap_b200.framework import packing as P
    torch.manual_seed(1)
    bf = lambda t: t.to(torch.bfloat16).float()
    T = 6
    for Ci, Co, H in ((4, 8, 11), (8, 16, 10), (16, 32, 9), (3, 5, 8)):
        x = bf(torch.randn(1, Ci, H, T))
        Cip, Cop = P.pad8(Ci), P.pad8(Co)
        xp = torch.zeros(H, T, Cip)
        xp[..., :Ci] = x[0].permute(1, 2, 0)
        # ---- strided conv (4,1)/(2,1): K = (kh, ci) over input rows 2q .. 2q+3
        w, b = bf(torch.randn(Co, Ci, 4, 1)), torch.randn(Co)
        want = F.conv2d(x, w, b, stride=(2, 1))[0]                                   # (Co, Hq, T)
        Hq = want.shape[1]
        a = torch.stack([xp[2 * q: 2 * q + 4].permute(1, 0, 2).reshape(T, 4 * Cip) for q in range(Hq)])      # (Hq, T, K)
        got = (a @ _unpack_b(P.pack_down(w)).t())[..., :Co].permute(2, 0, 1) + b.view(-1, 1, 1)
        assert torch.allclose(got, want, atol=1e-4)
        ws = _unpack_b(P.pack_down_strip(w, b))                                      # K groups: rows (+ per-row filler / bias groups)
        if Cip == 8:
            a_s = torch.stack([torch.cat([torch.cat([xp[2 * q + kh], _ones_k(T) if kh == 0 else torch.zeros(T, 8)], -1) for kh in range(4)], -1)
                               for q in range(Hq)])
        else:
            a_s = torch.cat([a, _ones_k(T).expand(Hq, T, 8), torch.zeros(Hq, T, 8)], -1)
        got = (a_s @ ws.t())[..., :Co].permute(2, 0, 1)
        assert torch.allclose(got, want, atol=2e-3)                                  # bias as bf16 hi + lo
        if Ci <= 4 and Co <= 8:                                                      # packed 4-channel input: a GEMM row is a frame pair
            wp = _unpack_b(P.pack_down_pairs(w, b))
            x4 = torch.zeros(H, T // 2, 2, 4)
            x4[..., :Ci] = x[0].permute(1, 2, 0).reshape(H, T // 2, 2, Ci)
            x4 = x4.reshape(H, T // 2, 8)
            a_p = torch.stack([torch.cat([torch.cat([x4[2 * q + kh], _ones_k(T // 2) if kh == 0 else torch.zeros(T // 2, 8)], -1) for kh in range(4)], -1)
                               for q in range(Hq)])
            got = (a_p @ wp.t()).reshape(Hq, T // 2, 2, 8)[..., :Co].reshape(Hq, T, Co).permute(2, 0, 1)
            assert torch.allclose(got, want, atol=2e-3)
        # ---- transposed conv (4,1)/(2,1) as a polyphase GEMM: rows q-1, q -> output rows 2q, 2q+1
        wt, bt = bf(torch.randn(Ci, Co, 4, 1)), torch.randn(Co)
        for op in (0, 1):
            want = F.conv_transpose2d(x, wt, bt, stride=(2, 1), output_padding=(op, 0))[0]                   # (Co, 2H + 2 + op, T)
            Ho = want.shape[1]
            xz = torch.cat([torch.zeros(1, T, Cip), xp, torch.zeros(2, T, Cip)])    # rows -1 .. H+1
            rows = []
            for q in range((Ho + 1) // 2):
                a_q = torch.cat([xz[q], xz[q + 1]], -1)                              # input rows q-1, q
                y = a_q @ _unpack_b(P.pack_up(wt)).t() + P.pack_up_bias(bt, Co)
                if Cip > 8:
                    a_s = torch.cat([a_q, _ones_k(T), torch.zeros(T, 8)], -1)
                else:
                    a_s = torch.cat([xz[q], _ones_k(T), xz[q + 1], torch.zeros(T, 8)], -1)
                ys = a_s @ _unpack_b(P.pack_up_strip(wt, bt)).t()
                assert torch.allclose(y, ys, atol=2e-3)
                rows += [y[:, :Co], y[:, Cop:Cop + Co]]
            got = torch.stack(rows[:Ho]).permute(2, 0, 1)
            assert torch.allclose(got, want, atol=1e-4), (Ci, Co, op)
    # ---- Encoder.convlat: one GEMM over the full height, K = (kh, ci)
    C4, H4, D = 16, 5, 24
    x = bf(torch.randn(1, C4, H4, T))
    w = bf(torch.randn(D, C4, H4, 1))
    want = F.conv2d(x, w)[0, :, 0]                                                   # (D, T)
    a = x[0].permute(2, 1, 0).reshape(T, H4 * C4)
    assert torch.allclose((a @ _unpack_b(P.pack_lat(w, 32)).t())[:, :D].t(), want, atol=1e-4)
    # ---- Decoder.convin: per output row its own (C0 x D) operand; the indicator channel is a bias table
    C0 = 16
    wd, bd = bf(torch.randn(D + 1, C0, H4, 1)), torch.randn(C0)
    lat = bf(torch.randn(1, D, 1, T))
    packed, tables = P.pack_deconv_in(wd, bd, 32)
    for ind in (0, 1):
        want = F.conv_transpose2d(torch.cat([lat, torch.full((1, 1, 1, T), float(ind))], 1), wd, bd)[0]      # (C0, H4, T)
        latp = torch.zeros(T, 32)
        latp[:, :D] = lat[0, :, 0].t()
        got = torch.stack([latp @ packed[h].float().permute(1, 0, 2).reshape(packed.shape[2], -1).t()[:, :C0] + tables[ind, h, :C0] for h in range(H4)])
        assert torch.allclose(got.permute(2, 0, 1), want, atol=1e-4)


def test_pipeline_numa_helpers(monkeypatch, tmp_path):
    """framework.pipeline: the GPU-local CPU list is parsed from sysfs; near_gpu narrows the affinity to it and restores it, and is a
    no-op where the kernel gives no answer (no GPU needed: the device properties and the sysfs path are stubbed)."""
    import builtins
    import os
    from timbre_trap_b200.framework import pipeline as PL

    class Props:
        pci_domain_id, pci_bus_id, pci_device_id = 0, 0x1b, 0

    monkeypatch.setattr(PL.torch.cuda, 'get_device_properties', lambda d: Props())
    real_open = builtins.open
    listing = {'text': '0-1,3\n'}

    def fake_open(path, *a, **k):
        if str(path).endswith('0000:1b:00.0/local_cpulist'):
            f = tmp_path / 'cpulist'
            f.write_text(listing['text'])
            return real_open(f, *a, **k)
        return real_open(path, *a, **k)

    monkeypatch.setattr(builtins, 'open', fake_open)
    assert PL.gpu_local_cpus('cuda:0') == {0, 1, 3}
    listing['text'] = '\n'
    assert PL.gpu_local_cpus('cuda:0') is None
    if hasattr(os, 'sched_getaffinity'):
        before = os.sched_getaffinity(0)
        with PL.near_gpu('cuda:0') as moved:                     # no list: nothing changes
            assert moved is False and os.sched_getaffinity(0) == before
        listing['text'] = ','.join(str(c) for c in sorted(before)[:1]) + '\n'
        with PL.near_gpu('cuda:0') as moved:
            assert os.sched_getaffinity(0) == set(sorted(before)[:1]) and moved == (len(before) > 1)
        assert os.sched_getaffinity(0) == before
    monkeypatch.setattr(PL.torch.cuda, 'get_device_properties', lambda d: (_ for _ in ()).throw(RuntimeError('no device')))
    assert PL.gpu_local_cpus('cuda:0') is None
