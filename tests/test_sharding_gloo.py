"""
World-size-2 (gloo, CPU) check of the sharded `reconstruct` normalisation: every rank synthesises its shard WITHOUT
normalising, the per-rank peaks are MAX-all-reduced, each rank scales locally - the result must equal the reference's
batch-global `audio /= audio.abs().max()` (cqtwrapper.py:209-211).  Uses the CPU oracle for the synthesis; the GPU
path runs the identical three steps (TimbreTrap._decode_shared_peak) over NCCL.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle.model_ref import CQTRef
    from tests.helpers import tonal_clip
    cqt = CQTRef(6, 12, 8000, 0.5)
    audio = tonal_clip(2 * cqt.block_length, 8000, seed=3, n_batch=4) * torch.tensor([0.2, 1.0, 0.5, 0.7]).view(4, 1, 1)
    coeffs = cqt(audio)
    shard = coeffs[rank * 2:(rank + 1) * 2]                        # block sharding: 2 items per rank
    raw = cqt.decode_raw(shard)
    peak = raw.abs().max().reshape(1)
    dist.all_reduce(peak, op=dist.ReduceOp.MAX)
    local = raw / peak if float(peak) else raw
    want = cqt.decode(coeffs)[rank * 2:(rank + 1) * 2]             # the reference semantics on the whole batch
    np.save(os.path.join(out_dir, f'err{rank}.npy'), np.array([float((local - want).abs().max()), float(local.abs().max())]))
    dist.destroy_process_group()


def test_sharded_peak_normalise_equals_global(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    errs = [np.load(os.path.join(tmp_path, f'err{r}.npy')) for r in range(2)]
    assert max(e[0] for e in errs) < 1e-6
    assert abs(max(e[1] for e in errs) - 1.0) < 1e-6               # exactly one shard holds the global peak


def _grad_worker(rank, world, port, out_dir):
    """Loss step sharding: each rank back-propagates the reference's losses (oracle, CPU fp32) on ITS half of the batch, then
    framework.train.allreduce_mean_gradients averages the flat bucket - the result must be the gradient of the whole batch."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from timbre_trap_b200.framework.train import allreduce_mean_gradients
    torch.manual_seed(0)
    # a small stand-in with the same structure of the problem: per-item losses averaged over the batch
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    data = torch.randn(4, 3)

    def loss_of(batch):
        h = torch.tanh(batch @ params[0].t())                        # (b, 5)
        return ((h.sum(-1, keepdim=True) * params[1]) ** 2).sum(-1).mean()
    for p in params:
        p.grad = None
    loss_of(data[rank * 2:(rank + 1) * 2]).backward()
    allreduce_mean_gradients(params, dist.group.WORLD)
    got = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    loss_of(data).backward()
    err = max(float((g - p.grad).abs().max()) for g, p in zip(got, params))
    np.save(os.path.join(out_dir, f'gerr{rank}.npy'), np.array([err]))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_equals_global_batch(tmp_path):
    port = 31500 + os.getpid() % 2000
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    errs = [float(np.load(os.path.join(tmp_path, f'gerr{r}.npy'))[0]) for r in range(2)]
    assert max(errs) < 1e-5


def _clip_worker(rank, world, port, out_dir):
    """One long clip over several ranks (BASELINE.json configs[4]): every rank runs the reference's chunk loop (oracle) on ITS
    blocks plus the half-block halos cut by framework.modules.shard_audio, the product's all_gather helper reassembles the
    frames - the result must equal the unsharded loop on the whole clip."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import model_ref as R
    from tests.helpers import tonal_clip
    from timbre_trap_b200.framework.modules import _gather_frames, shard_audio
    cqt = R.CQTRef(6, 12, 8000, 0.5)
    sd = R.init_state_dict(cqt.n_bins, 16, 1, seed=0)
    L, M = cqt.block_length, cqt.max_window_length
    audio = tonal_clip(5 * L - 123, 8000, seed=9)                       # 5 blocks after padding: ranks own 2 and 3 blocks
    sub, b0, b1 = shard_audio(cqt.pad_to_block_length(audio), L, rank, world)
    local = R.chunked_inference_ref(sub, sd, cqt, True, prepadded=True)             # (1, 2, F, (b1 - b0) * M)
    whole = R.chunked_inference_ref(audio, sd, cqt, True)
    err_local = float((local - whole[..., b0 * M: b1 * M]).abs().max())
    # the product resolves `group=None` to the default process group under torch.distributed (a call that shards by the
    # default group's ranks must gather over the same ranks)
    from timbre_trap_b200.framework.modules import _resolve_group
    group, r_, w_ = _resolve_group(None)
    assert group is dist.group.WORLD and (r_, w_) == (rank, world)
    assert _resolve_group(None, 1, 3) == (None, 1, 3)                  # explicit rank / world: emulation, no collective
    gathered = _gather_frames(local.contiguous(), -1, M, audio, L, group, w_)
    err_gather = float((gathered - whole).abs().max()) if gathered.shape == whole.shape else 1e9
    np.save(os.path.join(out_dir, f'cerr{rank}.npy'), np.array([err_local, err_gather, b1 - b0]))
    dist.destroy_process_group()


def test_long_clip_sharded_by_blocks_equals_unsharded(tmp_path):
    port = 33500 + os.getpid() % 2000
    mp.spawn(_clip_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [np.load(os.path.join(tmp_path, f'cerr{r}.npy')) for r in range(2)]
    assert sorted(int(r[2]) for r in res) == [2, 3]
    assert max(r[0] for r in res) < 1e-5 and max(r[1] for r in res) < 1e-5
