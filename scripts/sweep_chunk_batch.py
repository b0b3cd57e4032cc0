"""Device time of transcribe_and_reconstruct (256 x 3 s blocks, base model) against TimbreTrap.MAX_CHUNKS_PER_BATCH."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from timbre_trap_b200.framework import TimbreTrap
torch.manual_seed(0)
model = TimbreTrap(22050, 9, 60, 3, latent_size=128, model_complexity=2).cuda().eval()
g = torch.Generator(device='cuda').manual_seed(1)
audio = torch.rand((256, 1, 66150), device='cuda', generator=g) * 2 - 1
for mc in (128, 192, 256, 384, 768):
    model.MAX_CHUNKS_PER_BATCH = mc
    for _ in range(3): model.transcribe_and_reconstruct(audio)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8): model.transcribe_and_reconstruct(audio)
    b.record(); torch.cuda.synchronize()
    print(mc, 'chunks per batch: %.2f ms/step' % (a.elapsed_time(b) / 8), 'peak mem %.1f GB' % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
