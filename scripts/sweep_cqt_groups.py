"""Sweep of the CQT plan's group size (blocks per launch) and lane count (internal streams) at BASELINE configs[1] (1024 blocks)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import CQT
n_blocks = 1024
g = torch.Generator(device='cuda').manual_seed(0)
audio = torch.rand((n_blocks, 1, 66150), device='cuda', generator=g) * 2 - 1
bpb = 4 * 66150 + 8 * 540 * 1024
def timeit(fn, iters=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for lanes in (1, 2, 3, 4):
    for mb in (24, 32, 48, 64, 96):
        CQT.BLOCKS_PER_LAUNCH, CQT.LANES = mb, lanes
        cqt = CQT(9, 60, 22050, 3)
        coeffs = cqt.encode_interleaved(audio)
        f = timeit(lambda: cqt.encode_interleaved(audio)); i = timeit(lambda: cqt.decode_raw(coeffs.permute(0, 3, 1, 2), normalise=True))
        print('lanes', lanes, 'blocks', mb, 'fwd %.3f ms %.1f%%  inv %.3f ms %.1f%%' % (f, n_blocks*bpb/f/1e6/6538.9*100, i, n_blocks*bpb/i/1e6/6538.9*100), flush=True)
        del cqt, coeffs
