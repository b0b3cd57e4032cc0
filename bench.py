"""
Benchmark of the Timbre-Trap hot path on B200 (contract: see the round prompt / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--blocks 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the audio-s/s metric is quoted on): transcribe + reconstruct of a
batch of 256 synthetic 3 s blocks (768 audio-seconds) per GPU with the base model (9 oct x 60 bpo, 22.05 kHz, latent
128, complexity 2, random init, bf16 convs / fp32 CQT).  One step = TimbreTrap.transcribe_and_reconstruct on the
whole batch: chunking (50 % overlap -> 3 chunks per block), CQT analysis, encoder (once), decoder (twice), Hann
cross-fade, tanh|.| activations, CQT synthesis + peak normalise.  Multi-GPU: blocks are sharded across ranks (weak
scaling, no data-path collective; one scalar MAX all-reduce for the reference's global peak normalise).

Printed by rank 0: ONE JSON line (metric/value/e2e/roofline/cpu_baseline/clocks/...).  `--impl reference` times the CPU
restatement of the reference path (oracle/, the reference itself is pure Python + a missing third-party CQT and does not
exist on the GPU box) on the host cores, on a bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR, N_OCT, BPO, SECS, LATENT, COMPLEXITY = 22050, 9, 60, 3, 128, 2
L, F, M = 66150, 540, 1024
METRIC = 'audio-sec/sec transcribe+reconstruct'
UNIT = 'audio-s/s'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], bf16_burst=p['bf16_tflops'], bf16_sustained=p['bf16_tflops_sustained'], source='measured')
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=int(self.rows[0][1]) if self.rows else None,
                    samples=len(sm), reasons=sorted(reasons))


def synthetic_audio(n_blocks, seed, device='cpu', pin=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n_blocks, 1, L), generator=g) * 2 - 1
    if pin:
        x = x.pin_memory()
    return x.to(device) if device != 'cpu' else x


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (model.transcribe(audio); model.reconstruct(audio), modules.py:292-336)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_step(state, n_blocks):
    from oracle import model_ref as R
    sd, cqt = state
    audio = synthetic_audio(n_blocks, seed=0)
    t0 = time.perf_counter()
    with torch.no_grad():
        R.transcribe_ref(audio, sd, cqt)
        R.reconstruct_ref(audio, sd, cqt)
    return time.perf_counter() - t0


def cpu_state():
    from oracle import model_ref as R
    import numpy as np
    torch.set_num_threads(os.cpu_count() or 1)
    return R.init_state_dict(F, LATENT, COMPLEXITY, seed=0), R.CQTRef(N_OCT, BPO, SR, SECS, dtype=np.complex64)


WORKLOAD = ('BASELINE.json configs[2]: transcribe+reconstruct, 256 x 3 s blocks per GPU, base model '
            '(9 oct x 60 bpo, 22.05 kHz, latent 128, complexity 2), random init')


def run_reference(args, rank):
    if rank != 0:
        return
    state = cpu_state()
    n_blocks = 1                              # bounded sample: one 3 s block per step (= BASELINE.json configs[0])
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_step(state, n_blocks)
    times = [cpu_reference_step(state, n_blocks) for _ in range(args.steps)]
    per_step = sum(times) / len(times)
    value = n_blocks * SECS / per_step
    cores = torch.get_num_threads()
    line = dict(metric=METRIC, value=value, unit=UNIT, impl='reference', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per_step * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD, sample=f'bounded sample of that workload: {n_blocks} x 3 s block per step on the host CPU '
                                                      '(the reference\'s sequential chunk loop)'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind='port',
                                  sample=f'{args.steps} steps x {n_blocks} block(s) of 3 s; oracle/model_ref.py + oracle/nsgt_ref.py (the reference is '
                                         'pure Python with an un-vendored CQT dependency; /root/reference is absent on the GPU box)'),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters          # ms


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from timbre_trap_b200 import _lib
    from timbre_trap_b200.framework import TimbreTrap, ops, packing as P

    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        group = dist.group.WORLD

    torch.manual_seed(0)
    model = TimbreTrap(SR, N_OCT, BPO, SECS, latent_size=LATENT, model_complexity=COMPLEXITY).to(device).eval()
    n_blocks = args.blocks
    host_audio = synthetic_audio(n_blocks, seed=1000 + rank, pin=True)
    dev_audio = host_audio.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return model.transcribe_and_reconstruct(dev_audio, group=group)

    host_act = torch.empty((n_blocks, F, M), dtype=torch.float32).pin_memory()
    host_wav = torch.empty((n_blocks, 1, L), dtype=torch.float32).pin_memory()

    def step_e2e():
        x = host_audio.to(device, non_blocking=True)
        act, wav = model.transcribe_and_reconstruct(x, group=group)
        host_act.copy_(act, non_blocking=True)
        host_wav.copy_(wav, non_blocking=True)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.lib().tt_launch_count(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step_resident()
    b.record()
    barrier()
    launches = int(_lib.lib().tt_launch_count(1))
    ms_total = max_over_ranks(a.elapsed_time(b))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * n_blocks * SECS / (ms_step * 1e-3)

    # ---- end-to-end arm (pinned host buffers in, host buffers out, every step) ----
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()
    e2e = dict(value=world * n_blocks * SECS / (e2e_ms * 1e-3), unit=UNIT, ms_per_step=e2e_ms,
               h2d_bytes_per_step=host_audio.numel() * 4, d2h_bytes_per_step=(host_act.numel() + host_wav.numel()) * 4)

    line = None
    if rank == 0:
        pk = peaks()
        # ---- roofline of the dominant kernel: the fused residual-block kernel (res_rs_kernel, ~60 % of the step).  Every
        # instance is HBM-bound (arithmetic intensity 17..136 FLOP/B against a ridge of ~210); the line reports the first-stage
        # instance (C = 4, packed layout, the largest tensors) against the measured copy bandwidth, and the C = 32 instance
        # against the measured bf16 peak as `roofline_tensor` (the north star's tensor-pipe view).
        n_chunks = min(3 * n_blocks, model.MAX_CHUNKS_PER_BATCH)
        blk4 = model.encoder.block1.block1
        x4 = torch.randn((n_chunks, F, M, 4), device=device).to(torch.bfloat16)
        y4 = torch.empty_like(x4)
        ms_4 = time_kernel(lambda: blk4.forward_c8(x4, out=y4), iters=10)
        bytes_4 = 2.0 * x4.numel() * 2                      # read x + write y, un-padded 4 channels, bf16
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r01_res_rs_c4_traffic.json')
        if os.path.exists(tpath):
            t = json.load(open(tpath))
            traffic = t['dram_bytes_per_launch'] * (n_chunks / t['chunks'])
        roofline = dict(kernel='res_rs_kernel<2,16,1,4> (fused ResidualConv2dBlock, C=4 packed layout folded to 4 frames per GEMM row, dilation 1, H=540)', bound='hbm',
                        achieved=bytes_4 / (ms_4 * 1e-3) / 1e9, peak=pk['hbm'], unit='GB/s', frac=bytes_4 / (ms_4 * 1e-3) / 1e9 / pk['hbm'],
                        traffic=traffic, peak_source=f"{pk['source']} HBM copy bandwidth", us_per_launch=ms_4 * 1e3,
                        bytes_per_launch=bytes_4, chunks_per_launch=n_chunks)
        del x4, y4
        blk = model.encoder.block4.block2
        x32 = torch.randn((n_chunks, 4, 65, M, 8), device=device).to(torch.bfloat16)
        y32 = torch.empty_like(x32)
        ms_k = time_kernel(lambda: blk.forward_c8(x32, out=y32), iters=10)
        flops = 2.0 * (9 * 32 * 32 + 32 * 32) * 65 * M * n_chunks
        achieved = flops / (ms_k * 1e-3) / 1e12
        roofline_tensor = dict(kernel='res_rs_kernel<4,32,2,0> (fused ResidualConv2dBlock, C=32, dilation 2, H=65)', bound='tensor',
                               achieved=achieved, peak=pk['bf16_burst'], unit='TFLOP/s', frac=achieved / pk['bf16_burst'],
                               peak_source=f"{pk['source']} bf16 burst (kernel timed alone)", us_per_launch=ms_k * 1e3,
                               flops_per_launch=flops, hbm_gbs=2.0 * x32.numel() * 2 / (ms_k * 1e-3) / 1e9)
        del x32, y32
        # ---- CQT forward / inverse against the HBM roofline (BASELINE.json configs[1]: 1024 blocks) ----
        cq = {}
        nb = 1024
        big = synthetic_audio(nb, seed=7, device=device)
        coeffs = model.sliCQ.encode_interleaved(big)
        bytes_dir = nb * (4 * L + 8 * F * M)
        for name, fn in (('forward', lambda: model.sliCQ.encode_interleaved(big)),
                         ('inverse', lambda: model.sliCQ.decode_raw(coeffs.permute(0, 3, 1, 2), normalise=True))):
            ms = time_kernel(fn, iters=5)
            cq[name] = dict(ms=ms, achieved=bytes_dir / (ms * 1e-3) / 1e9, peak=pk['hbm'], unit='GB/s',
                            frac=bytes_dir / (ms * 1e-3) / 1e9 / pk['hbm'], bytes=bytes_dir)
        cq['round_trip_audio_s_per_s'] = nb * SECS / ((cq['forward']['ms'] + cq['inverse']['ms']) * 1e-3)
        cq['workload'] = 'BASELINE.json configs[1]: 1024 x 3 s blocks, 540 bins, fp32; whole path (all kernels of the direction)'
        del big, coeffs
        torch.cuda.empty_cache()

        cpu = None
        if world == 1 and not args.no_cpu:
            state = cpu_state()
            cpu_reference_step(state, 1)
            times = []
            t_start = time.perf_counter()
            while len(times) < 3 or (time.perf_counter() - t_start < 10 and len(times) < 10):
                times.append(cpu_reference_step(state, 1))
            v = SECS / (sum(times) / len(times))
            cpu = dict(value=v, unit=UNIT, cores=torch.get_num_threads(), kind='port',
                       sample=f'{len(times)} x (transcribe + reconstruct of one 3 s block, sequential 3-chunk loops) with oracle/model_ref.py')

        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                    higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16 convs (fp32 accumulate), fp32 CQT',
                    data='synthetic', impl='ours',
                    config=dict(workload=WORKLOAD,
                                blocks_per_gpu=n_blocks, chunks_per_gpu=3 * n_blocks, parallelism=f'dp{world} (block sharding)',
                                l2='working set per step ~40 GB >> 126 MB L2; no explicit flush'),
                    e2e=e2e, gpu_launches=launches, roofline=roofline, roofline_tensor=roofline_tensor, cqt=cq, cpu_baseline=cpu, clocks=clocks)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--blocks', type=int, default=256, help='3 s blocks per GPU per step')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device: timbre_trap_b200 has no CPU path')
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
