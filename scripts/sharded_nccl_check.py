"""
Two (or more) NCCL ranks run TimbreTrap.transcribe_sharded / reconstruct_sharded on one long clip every rank holds
(BASELINE.json configs[4]) with `group=None` (resolved to the default process group) and compare with the unsharded
calls on the same GPU.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1
--master-port 29611 scripts/sharded_nccl_check.py [n_blocks]   (on the GPU box: `gpurun --gpus 2 -- ...`).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from timbre_trap_b200.framework import TimbreTrap

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 21
torch.manual_seed(0)
model = TimbreTrap(22050, 9, 60, 3, latent_size=128, model_complexity=2).cuda().eval()
g = torch.Generator().manual_seed(5)
L = model.sliCQ.block_length
audio = ((torch.rand((1, 1, n_blocks * L - 1000), generator=g) * 2 - 1) * 0.5).cuda()

act = model.transcribe(audio)
wav = model.reconstruct(audio)
act_s = model.transcribe_sharded(audio)                 # group=None -> WORLD: sharded by rank AND gathered
wav_s = model.reconstruct_sharded(audio)
act_b, wav_b = model.transcribe_and_reconstruct_sharded(audio)
assert torch.equal(act_b, act_s) and torch.equal(wav_b, wav_s)
torch.cuda.synchronize()
res = dict(rank=rank, world=world, n_blocks=n_blocks, act_shape=list(act_s.shape), wav_shape=list(wav_s.shape),
           act_equal=bool(torch.equal(act_s, act)), act_max_diff=float((act_s - act).abs().max()),
           wav_max_diff=float((wav_s - wav).abs().max()), wav_peak=float(wav_s.abs().max()))
# timing of the sharded pair (device time, max over ranks)
for _ in range(2):
    model.transcribe_sharded(audio); model.reconstruct_sharded(audio)
dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    model.transcribe_sharded(audio); model.reconstruct_sharded(audio)
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 3], device='cuda')
dist.all_reduce(t, op=dist.ReduceOp.MAX)
res['ms_sharded_pair'] = float(t)
res['audio_s_per_s'] = audio.size(-1) / 22050 / (float(t) * 1e-3)
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    ok = all(r['act_equal'] and r['wav_max_diff'] <= 2e-6 and abs(r['wav_peak'] - 1.0) < 1e-5 for r in allres)
    print(json.dumps(dict(ok=ok, ranks=allres)))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(dict(ok=ok, ranks=allres), open(f'gpurun_out/sharded_nccl_check_w{world}.json', 'w'), indent=1)
    assert ok
dist.destroy_process_group()
