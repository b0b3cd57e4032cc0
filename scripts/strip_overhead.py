"""Per-strip set-up / fill / drain cost of the residual kernel: with 148 chunks of the C = 4 stage (296 tiles = one CTA per resident
slot) a launch with s strips per tile is exactly s waves, so time(s) = (H + s * overhead_rows) * cycles_per_row: the slope over s is
the overhead of one strip in row-equivalents."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from timbre_trap_b200 import _lib
from timbre_trap_b200.framework import TimbreTrap

dev = torch.device('cuda')
model = TimbreTrap(bench.SR, bench.N_OCT, bench.BPO, bench.SECS, latent_size=bench.LATENT, model_complexity=bench.COMPLEXITY).to(dev).eval()
F, M = model.sliCQ.n_bins, model.sliCQ.max_window_length
for c, blk, shape in ((4, model.encoder.block1, (148, F, M, 4)), (16, model.encoder.block3, (37, 2, 133, M, 8))):
    x = torch.randn(shape, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    H = shape[1] if c == 4 else shape[2]
    rb = blk.block1
    rows_list, ms_list = [], []
    for s in (1, 2, 3, 4, 5, 6, 8, 10, 12):
        rows = (H + s - 1) // s
        _lib.lib().tt_set_strip_rows(rows)
        ms = sorted(bench.time_kernel(lambda: rb.forward_c8(x, out=y), iters=20) for _ in range(3))[1]
        n_strips = (H + rows - 1) // rows
        rows_list.append(n_strips)
        ms_list.append(ms)
        print(f'C={c} H={H}: {n_strips:2d} strips of {rows:3d} rows: {ms * 1e3:8.1f} us')
    _lib.lib().tt_set_strip_rows(0)
    slope, icpt = np.polyfit(rows_list, ms_list, 1)
    print(f'  fit: {icpt * 1e3:.1f} us + {slope * 1e3:.2f} us per strip  ->  one strip costs {slope / (icpt / H):.1f} row-equivalents '
          f'({100 * slope / (icpt / H) / (H / 4):.1f} % at 4 strips)')
