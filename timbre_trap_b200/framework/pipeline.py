"""
Host-to-host streaming of `TimbreTrap.transcribe_and_reconstruct` (not in the reference, whose scripts move one batch at a time
with blocking `.to(device)` / `.cpu()` calls, e.g. experiments/comparison.py:209-255).

A production caller holds audio in pinned host memory and wants the activations / audio back on the host.  Issued naively on the
compute stream, the device-to-host read of step k (0.63 GB for 256 blocks) delays step k+1 and, with one process per GPU, all the
ranks' reads contend for host memory at the same moment.  `HostPipeline` keeps three streams - copy-in, compute, copy-out - and
`depth` sets of pinned output buffers, so the read-back of step k overlaps the compute of step k+1:

    pipe = HostPipeline(model, depth=3)
    tickets = [pipe.submit(batch) for batch in pinned_batches]      # returns immediately (blocks only when `depth` are in flight)
    activations, audio = pipe.collect(ticket)                       # pinned host tensors, valid until the slot is reused

Two details matter once eight ranks share one host (measured, profiles/r02_e2e_readback_check_n8.txt: a rank's 0.63 GB read-back takes
11 ms alone and 35-55 ms when all eight copy at once):
  * `early`: the model hands over the activations of the clips whose chunks are through the decoder after every chunk batch
    (TimbreTrap.transcribe_and_reconstruct(on_activations=...)); they start their way to the host while the next chunk batch
    computes, so only the last batch's share and the audio are left to copy when the step's last kernel ends - the drain of the
    pipeline after the final step shrinks from the whole read-back to about a third of it;
  * keep TWO submissions in flight (depth >= 3, collect ticket k - 2 after submitting k): the host then never waits for a
    read-back before it may enqueue the next step's ~2000 launches, and the GPU never runs dry.
"""

import contextlib
import os

import torch

__all__ = ['HostPipeline', 'gpu_local_cpus', 'near_gpu']


def gpu_local_cpus(device):
    """The CPUs of the NUMA node the GPU's PCIe root port hangs off (sysfs `local_cpulist`), or None when the kernel does not say."""
    try:
        p = torch.cuda.get_device_properties(device)
        path = f'/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/local_cpulist'
        with open(path) as f:
            text = f.read().strip()
    except (OSError, AttributeError, RuntimeError):
        return None
    cpus = set()
    for part in filter(None, text.split(',')):
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus or None


@contextlib.contextmanager
def near_gpu(device):
    """Run the body on the GPU-local CPUs, so that pinned host memory allocated inside lands on the GPU's own NUMA node (first-touch
    policy).  With one process per GPU and every rank reading back 0.63 GB per step, buffers that all sit on the node the launcher
    happened to start on send half of the ranks' traffic across the socket interconnect.  A no-op where sysfs gives no answer."""
    if not hasattr(os, 'sched_getaffinity'):
        yield False
        return
    before = os.sched_getaffinity(0)
    local = gpu_local_cpus(device)
    want = (local & before) if local else None
    if not want or want == before:
        yield False
        return
    os.sched_setaffinity(0, want)
    try:
        yield True
    finally:
        os.sched_setaffinity(0, before)


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
    def __init__(self, model, device=None, depth=3, group=None, numa_local=True, early=True):
        self.model = model
        self.numa_local = numa_local
        self.early = bool(early)
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
            self._run(slot, comp)
        return k

    def _host_buffers(self, slot, act_shape, wav_shape, dtype):
        if slot.host_act is None or tuple(slot.host_act.shape) != tuple(act_shape) or tuple(slot.host_wav.shape) != tuple(wav_shape):
            slot.host_act = self.pinned_empty(act_shape, dtype)
            slot.host_wav = self.pinned_empty(wav_shape, dtype)

    def _run(self, slot, comp):
        B = slot.dev_in.size(0)
        sent = []                                          # clip ranges whose activations are already on their way

        def on_activations(lo, hi, act):
            if slot.host_act is None or tuple(slot.host_act.shape) != (B,) + tuple(act.shape[1:]):
                return                                     # first use of this slot / new shape: everything goes at the end
            ready = torch.cuda.Event()
            ready.record(comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ready)
                slot.host_act[lo:hi].copy_(act, non_blocking=True)
            sent.append((lo, hi))

        act, wav = self.model.transcribe_and_reconstruct(slot.dev_in, group=self.group, on_activations=on_activations if self.early else None)
        slot.comp_done.record(comp)
        self._host_buffers(slot, act.shape, wav.shape, act.dtype)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot.comp_done)
            done = sent[-1][1] if sent else 0              # the hook is called with consecutive ranges starting at clip 0
            if done < B:
                slot.host_act[done:].copy_(act[done:], non_blocking=True)
            slot.host_wav.copy_(wav, non_blocking=True)
            slot.out_done.record(self.s_out)
        # the results are read on another stream: keep the allocator from handing their memory to the next step early
        act.record_stream(self.s_out)
        wav.record_stream(self.s_out)

    def pinned_empty(self, shape, dtype=torch.float32):
        """A pinned host tensor on the GPU's NUMA node (for the caller's input batches as well)."""
        with near_gpu(self.device) if self.numa_local else contextlib.nullcontext():
            return torch.empty(tuple(shape), dtype=dtype, pin_memory=True)

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
