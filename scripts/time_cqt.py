"""Quick device timing of the CQT forward / inverse paths (CUDA events), used while tuning."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import CQT

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cqt = CQT(9, 60, 22050, 3)
g = torch.Generator(device='cuda').manual_seed(0)
audio = torch.rand((n_blocks, 1, 66150), device='cuda', generator=g) * 2 - 1
bytes_per_block = 4 * 66150 + 8 * 540 * 1024

def timeit(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]

coeffs = cqt.encode_interleaved(audio)
res = {}
for name, fn in (('forward', lambda: cqt.encode_interleaved(audio)), ('inverse', lambda: cqt.decode_raw(coeffs.permute(0, 3, 1, 2)))):
    best, med = timeit(fn)
    res[name] = dict(ms_best=best, ms_median=med, GBps=n_blocks * bytes_per_block / (med * 1e-3) / 1e9,
                     frac_of_6538=n_blocks * bytes_per_block / (med * 1e-3) / 1e9 / 6538.9)
print(json.dumps(dict(n_blocks=n_blocks, **res), indent=1))
