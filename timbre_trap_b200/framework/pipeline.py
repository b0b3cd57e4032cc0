"""
Host-to-host streaming of `TimbreTrap.transcribe_and_reconstruct` (not in the reference, whose scripts move one batch at a time
with blocking `.to(device)` / `.cpu()` calls, e.g. experiments/comparison.py:209-255).

A production caller holds audio in pinned host memory and wants the activations / audio back on the host.  Issued naively on the
compute stream, the device-to-host read of step k (0.63 GB for 256 blocks) delays step k+1 and, with one process per GPU, all the
ranks' reads contend for host memory at the same moment.  `HostPipeline` keeps three streams - copy-in, compute, copy-out - and
`depth` sets of pinned output buffers, so the read-back of step k overlaps the compute of step k+1:

    pipe = HostPipeline(model, depth=2)
    tickets = [pipe.submit(batch) for batch in pinned_batches]      # returns immediately (blocks only when `depth` are in flight)
    activations, audio = pipe.collect(ticket)                       # pinned host tensors, valid until the slot is reused
"""

import torch

__all__ = ['HostPipeline']


class _Slot:
    def __init__(self):
        self.dev_in = None
        self.host_act = self.host_wav = None
        self.in_done = torch.cuda.Event()
        self.comp_done = torch.cuda.Event()
        self.out_done = torch.cuda.Event()
        self.busy = False
        self.ticket = -1


class HostPipeline:
    def __init__(self, model, device=None, depth=2, group=None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != 'cuda':
            raise ValueError('HostPipeline needs the model on a CUDA device')
        self.group = group
        with torch.cuda.device(self.device):
            self.s_in = torch.cuda.Stream()
            self.s_out = torch.cuda.Stream()
            self.slots = [_Slot() for _ in range(max(2, depth))]
        self.next_ticket = 0

    def submit(self, host_audio):
        """host_audio (B, 1, N) fp32, ideally pinned.  Returns a ticket for collect()."""
        k = self.next_ticket
        self.next_ticket += 1
        slot = self.slots[k % len(self.slots)]
        if slot.busy:                                   # the caller never collected it: its host buffers are about to be overwritten
            slot.out_done.synchronize()
        slot.busy, slot.ticket = True, k
        comp = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            if slot.dev_in is None or slot.dev_in.shape != host_audio.shape:
                slot.dev_in = torch.empty(host_audio.shape, dtype=torch.float32, device=self.device)
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(slot.comp_done)    # the previous user of this input buffer has been consumed
                slot.dev_in.copy_(host_audio, non_blocking=True)
                slot.in_done.record(self.s_in)
            comp.wait_event(slot.in_done)
            act, wav = self.model.transcribe_and_reconstruct(slot.dev_in, group=self.group)
            slot.comp_done.record(comp)
            if slot.host_act is None or slot.host_act.shape != act.shape or slot.host_wav.shape != wav.shape:
                slot.host_act = torch.empty(act.shape, dtype=act.dtype).pin_memory()
                slot.host_wav = torch.empty(wav.shape, dtype=wav.dtype).pin_memory()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot.comp_done)
                slot.host_act.copy_(act, non_blocking=True)
                slot.host_wav.copy_(wav, non_blocking=True)
                slot.out_done.record(self.s_out)
            # the results are read on another stream: keep the allocator from handing their memory to the next step early
            act.record_stream(self.s_out)
            wav.record_stream(self.s_out)
        return k

    def collect(self, ticket):
        """Blocks until the results of `ticket` are on the host; returns (activations (B, F, T), audio (B, 1, N')) pinned tensors
        that stay valid until `depth` further submissions."""
        slot = self.slots[ticket % len(self.slots)]
        if slot.ticket != ticket:
            raise ValueError(f'ticket {ticket} has been overwritten (at most {len(self.slots)} submissions may be in flight)')
        slot.out_done.synchronize()
        slot.busy = False
        return slot.host_act, slot.host_wav

    def bytes_per_step(self, host_audio):
        """(host-to-device, device-to-host) bytes one submission moves."""
        slot = self.slots[0]
        d2h = 0 if slot.host_act is None else (slot.host_act.numel() * slot.host_act.element_size() + slot.host_wav.numel() * slot.host_wav.element_size())
        return host_audio.numel() * host_audio.element_size(), d2h
