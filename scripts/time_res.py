"""Device timing of one residual block at a bench shape: python scripts/time_res.py C H d [B] [kernel: rs|fold] [iters]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import ops, packing as P

C, H, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
kernel = sys.argv[5] if len(sys.argv) > 5 else 'rs'
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
T = 1024
torch.manual_seed(0)
w1, b1 = torch.randn(C, C, 3, 3, device='cuda') * 0.2, torch.randn(C, device='cuda') * 0.1
w2, b2 = torch.randn(C, C, 1, 1, device='cuda') * 0.3, torch.randn(C, device='cuda') * 0.1
packed = C <= 4
x = (torch.randn((B, H, T, 4), device='cuda') if packed else torch.randn((B, (C + 7) // 8, H, T, 8), device='cuda')).to(torch.bfloat16)
y = torch.empty_like(x)
if kernel == 'fold':
    w1p, w2p, bias = P.pack_res_rs_fold(w1, b1, w2, b2, d, 4 if packed else 2)
    fn = lambda: ops.res_block_rs(x, w1p, w2p, bias, C, d, out=y, fold=True)
else:
    w1p, w2p, bias = P.pack_res_rs_pairs(w1, b1, w2, b2, d) if packed else P.pack_res_rs(w1, b1, w2, b2)
    fn = lambda: ops.res_block_rs(x, w1p, w2p, bias, C, d, out=y)
for _ in range(2): fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters): fn()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
by = 2.0 * B * H * T * C * 2
print(f'{kernel} C={C} H={H} d={d} B={B}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s')
