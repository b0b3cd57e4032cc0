"""Device timing of the first / last conv (conv_in, conv_out) on the packed 4-channel layout at the bench shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F, T = 540, 1024
torch.manual_seed(0)
coeffs = torch.randn(B, F, T, 2, device='cuda')
w_in, b_in = torch.randn(4, 2, 3, 3, device='cuda') * 0.2, torch.randn(4, device='cuda') * 0.1
w_out, b_out = torch.randn(2, 4, 3, 3, device='cuda') * 0.2, torch.randn(2, device='cuda') * 0.1
x = ops.conv_in(coeffs, w_in, b_in, 4, packed4=True)

def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

by = B * F * T * 16.0
for name, fn in (('conv_in', lambda: ops.conv_in(coeffs, w_in, b_in, 4, packed4=True)), ('conv_out', lambda: ops.conv_out(x, w_out, b_out, 4))):
    ms = timeit(fn)
    print(f'{name}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s')
