"""BASELINE.json configs[4] on ONE GPU: transcribe + reconstruct of a single 1-hour clip (1200 blocks -> 2401 overlapping chunks)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import TimbreTrap

secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
torch.manual_seed(0)
model = TimbreTrap(22050, 9, 60, 3, 128, 2).cuda().eval()
n = secs * 22050
g = torch.Generator(device='cuda').manual_seed(1)
audio = (torch.rand((1, 1, n), device='cuda', generator=g) * 2 - 1) * 0.5
with torch.no_grad():
    for _ in range(2):
        act, wav = model.transcribe_and_reconstruct(audio)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    act, wav = model.transcribe_and_reconstruct(audio)
    b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(json.dumps({'workload': f'{secs} s clip, one GPU, transcribe + reconstruct', 'ms': ms, 'audio_s_per_s': secs / (ms * 1e-3),
                  'activations': list(act.shape), 'audio_out': list(wav.shape), 'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}))
