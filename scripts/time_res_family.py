"""Device time of every (stage, dilation) instance of the fused residual-block kernel at the bench shapes (256 chunks per launch):
python scripts/time_res_family.py [iters].  One line per instance: ms and GB/s of algorithmic bytes (read x + write y)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from timbre_trap_b200.framework import TimbreTrap

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device('cuda')
torch.manual_seed(0)
model = TimbreTrap(bench.SR, bench.N_OCT, bench.BPO, bench.SECS, latent_size=bench.LATENT, model_complexity=bench.COMPLEXITY).to(dev).eval()
F, M, n = model.sliCQ.n_bins, model.sliCQ.max_window_length, 256
shapes = {4: (n, F, M, 4), 8: (n, 1, 269, M, 8), 16: (n, 2, 133, M, 8), 32: (n, 4, 65, M, 8)}
total = 0.0
for blk, c in ((model.encoder.block1, 4), (model.encoder.block2, 8), (model.encoder.block3, 16), (model.encoder.block4, 32)):
    x = torch.randn(shapes[c], device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    for rb in (blk.block1, blk.block2, blk.block3):
        best = []
        for _ in range(3):
            best.append(bench.time_kernel(lambda: rb.forward_c8(x, out=y), iters=iters))
        ms = sorted(best)[1]
        h = x.shape[1] if c == 4 else x.shape[2]
        nbytes = 2.0 * n * c * h * M * 2
        total += ms
        print(f'C={c:2d} d={rb.dilation}: {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s  (min {min(best):.4f} max {max(best):.4f})')
print(f'family: {total:.3f} ms')
