"""Device timing of each conv kernel at the bench shapes (256 chunks), used while tuning."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import TimbreTrap

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = TimbreTrap(22050, 9, 60, 3, 128, 2).cuda().eval()
T = 1024

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def c8(C, H): return torch.randn((B, (C + 7) // 8, H, T, 8), device='cuda').to(torch.bfloat16)
rows = []
enc, dec = m.encoder, m.decoder
shapes = [(4, 540), (8, 269), (16, 133), (32, 65)]
for i, (C, H) in enumerate(shapes):
    blk = getattr(enc, f'block{i+1}')
    x = torch.randn((B, H, T, 4), device='cuda').to(torch.bfloat16) if (i == 0 and enc.packed4) else c8(C, H)
    y = torch.empty_like(x)
    for d, rb in ((1, blk.block1), (2, blk.block2), (3, blk.block3)):
        ms = timeit(lambda: rb.forward_c8(x, out=y))
        fl = 2.0 * (9 * C * C + C * C) * H * T * B
        by = 2.0 * B * H * T * C * 2          # algorithmic: un-padded channels, read + write
        rows.append((f'res C={C} d={d}', ms, fl / ms / 1e9, by / ms / 1e6))
    ms = timeit(lambda: blk.forward_c8(x))
    rows.append((f'enc block{i+1} (3 res + down)', ms, 0, 0))
    del x, y
coeffs = torch.randn((B, 540, T, 2), device='cuda')
rows.append(('encoder total', timeit(lambda: enc.forward_c8(coeffs)), 2 * 4.686e9 * B / 1e9, 0))
lat, _ = enc.forward_c8(coeffs)
rows.append(('decoder total', timeit(lambda: dec.forward_c8(lat, True)), 2 * 4.688e9 * B / 1e9, 0))
for name, ms, gf, mb in rows:
    extra = ''
    if name.endswith('total'): extra = f' {gf / ms:8.1f} TFLOP/s'
    elif gf: extra = f' {gf:8.1f} TFLOP/s {mb:8.1f} GB/s'
    print(f'{name:32s} {ms:9.3f} ms{extra}')
