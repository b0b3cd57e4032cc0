"""
Where does the end-to-end time of `HostPipeline` go when N ranks read back at once?  Run under torch.distributed.run:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/e2e_readback_check.py

Per rank and for (buffer sets, early hand-over of the activations) = (2, no), (3, no), (3, yes): device-resident step,
end-to-end step through the pipeline, and the bare device-to-host copy of one step's results with all ranks copying at once.
Rank 0 prints one JSON line per arm with the per-rank figures.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
from timbre_trap_b200.framework import HostPipeline, TimbreTrap  # noqa: E402
from timbre_trap_b200.framework.pipeline import gpu_local_cpus  # noqa: E402


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        group = dist.group.WORLD
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.manual_seed(0)
    model = TimbreTrap(bench.SR, bench.N_OCT, bench.BPO, bench.SECS, latent_size=bench.LATENT, model_complexity=bench.COMPLEXITY).to(device).eval()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(x):
        t = torch.tensor([x], device=device, dtype=torch.float64)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(o), 3) for o in out]

    local_cpus = gpu_local_cpus(device)
    here = sorted(os.sched_getaffinity(0))
    info = dict(rank=rank, local_cpus=None if local_cpus is None else [min(local_cpus), max(local_cpus), len(local_cpus)], allowed=[here[0], here[-1], len(here)])
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps(dict(topology=infos)), flush=True)

    for depth, early in ((2, False), (3, False), (3, True), (2, False), (3, True)):
        numa_local = True
        pipe = HostPipeline(model, device, depth=depth, group=group, early=early)
        host_audio = pipe.pinned_empty((bench.N_BLOCKS if hasattr(bench, 'N_BLOCKS') else 256, 1, bench.L))
        g = torch.Generator().manual_seed(1000 + rank)
        host_audio.copy_(torch.rand(host_audio.shape, generator=g) * 2 - 1)
        dev_audio = host_audio.to(device)
        for _ in range(3):
            pipe.collect(pipe.submit(host_audio))
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            model.transcribe_and_reconstruct(dev_audio, group=group)
        b.record()
        barrier()
        dev_ms = a.elapsed_time(b) / steps
        barrier()
        t0 = time.perf_counter()
        in_flight = []
        for _ in range(steps):
            in_flight.append(pipe.submit(host_audio))
            if len(in_flight) > depth - 1:
                pipe.collect(in_flight.pop(0))
        for k in in_flight:
            act, wav = pipe.collect(k)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        barrier()
        # the bare read-back, all ranks at once
        d_act, d_wav = model.transcribe_and_reconstruct(dev_audio, group=group)
        barrier()
        a.record()
        for _ in range(4):
            act.copy_(d_act, non_blocking=True)
            wav.copy_(d_wav, non_blocking=True)
        b.record()
        barrier()
        d2h_ms = a.elapsed_time(b) / 4
        nbytes = act.numel() * 4 + wav.numel() * 4
        # ... and one rank at a time, for the uncontended figure
        solo = 0.0
        for r in range(world):
            barrier()
            if r == rank:
                a.record()
                act.copy_(d_act, non_blocking=True)
                wav.copy_(d_wav, non_blocking=True)
                b.record()
                torch.cuda.synchronize()
                solo = a.elapsed_time(b)
        barrier()
        rows = dict(depth=depth, early=early, steps=steps, device_ms=gather(dev_ms), e2e_ms=gather(e2e_ms), d2h_all_ms=gather(d2h_ms), d2h_solo_ms=gather(solo),
                    d2h_all_GBps=[round(nbytes / (m * 1e6), 1) for m in gather(d2h_ms)], d2h_bytes=nbytes)
        if rank == 0:
            print(json.dumps(rows), flush=True)
        del pipe, host_audio, act, wav, d_act, d_wav
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
