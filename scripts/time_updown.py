"""Device timing of the strided / transposed conv kernels at the bench shapes (256 chunks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import ops, packing as P

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 1024
torch.manual_seed(0)

def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def c8(C, H): return torch.randn((B, (C + 7) // 8, H, T, 8), device='cuda').to(torch.bfloat16)
for Ci, Co, H in ((4, 8, 540), (8, 16, 269), (16, 32, 133), (32, 64, 65)):
    w, b = torch.randn(Co, Ci, 4, 1, device='cuda') * 0.2, torch.randn(Co, device='cuda') * 0.1
    if Ci == 4:
        x = torch.randn((B, H, T, 4), device='cuda').to(torch.bfloat16)
        wp = P.pack_down_pairs(w, b)
    else:
        x = c8(Ci, H)
        wp = P.pack_down_strip(w, b)
    ms = timeit(lambda: ops.conv_down_strip(x, wp, P.pad8(Co)))
    print(f'down {Ci:2d}->{Co:2d} H={H}: {ms:.3f} ms')
    del x
for Ci, Co, H, op in ((64, 32, 31, 1), (32, 16, 65, 1), (16, 8, 133, 1), (8, 4, 269, 0)):
    w, b = torch.randn(Ci, Co, 4, 1, device='cuda') * 0.2, torch.randn(Co, device='cuda') * 0.1
    x = c8(Ci, H)
    wp = P.pack_up_strip(w, b)
    ms = timeit(lambda: ops.conv_up_strip(x, wp, P.pad8(Co), op, packed4_out=(Co == 4)))
    print(f'up   {Ci:2d}->{Co:2d} H={H}: {ms:.3f} ms')
    del x
