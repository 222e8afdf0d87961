"""Flat gradient storage and the overlapped data-parallel gradient all-reduce of the fused training steps.

The reference trains data-parallel with torch DDP (tools/train_pl.py:479) or nn.DataParallel (tools/train.py:185): autograd produces one
gradient tensor per parameter and DDP all-reduces them in buckets while the backward pass is still running.  Here the engine's own BPTT
writes every gradient during the fused step, so the equivalent structure is:

  * ``FlatGrads``: ONE persistent fp32 buffer per (model, device) holding every parameter gradient, laid out in the order the engine
    finishes them (gradient *groups*, include/capb200.h: capb200_*_set_grad_events).  The engine writes straight into it -- no per-step
    allocation, no torch.cat, no copy-back; ``param.grad`` becomes a view of it.
  * ``GradSync``: one NCCL all-reduce (ReduceOp.AVG: the division by the world size is folded into the collective) per group, issued on a
    communication stream that waits for the group's "complete" event, so the transfer of the logit / decoder gradients overlaps the rest
    of the backward pass and only the last group's transfer is exposed.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

ALIGN = 64          # floats: every gradient view starts on a 256-byte boundary (128-bit stores in the GEMM epilogues, NCCL alignment)


class FlatGrads:
    def __init__(self, groups: Sequence[Sequence[torch.nn.Parameter]], device):
        self.groups = [list(g) for g in groups]
        self.device = device
        offs, total = [], 0
        self.ranges = []
        for g in self.groups:
            start = total
            for p in g:
                offs.append((p, total))
                total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            self.ranges.append((start, total))
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.views = {id(p): self.flat[o:o + p.numel()].view(p.shape) for p, o in offs}
        self.params = [p for p, _ in offs]
        # one event per group, recorded by the engine when the group's last gradient has been written (a CPU buffer -- the layout tests --
        # has none)
        self.events = []
        if torch.device(device).type == 'cuda':
            self.events = [torch.cuda.Event() for _ in self.groups]
            with torch.cuda.device(device):
                for ev in self.events:
                    ev.record()          # creates the underlying cudaEvent_t (torch creates it lazily on first record)
        self._event_table = (ctypes.c_void_p * len(self.events))(*[ev.cuda_event for ev in self.events])

    def event_table(self):
        return self._event_table, len(self.events)

    def view(self, p):
        return self.views[id(p)]

    @property
    def nbytes(self):
        return self.flat.numel() * 4


class GradSync:
    """Chunked all-reduce of a FlatGrads buffer, overlapped with the backward pass that fills it."""

    def __init__(self, process_group=None):
        self.group = process_group
        self.stream: Optional[torch.cuda.Stream] = None
        self.done = None
        self.bytes = 0
        self._wait = (None, None)

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def launch(self, fg: FlatGrads):
        """Call right after the fused step returned (all of its kernels are enqueued, the group events recorded)."""
        self.bytes = 0
        self.done = None
        if self.world <= 1:
            return
        if self.stream is None or self.stream.device != fg.flat.device:
            self.stream = torch.cuda.Stream(device=fg.flat.device)
        op = dist.ReduceOp.AVG if dist.get_backend(self.group) == 'nccl' else dist.ReduceOp.SUM
        with torch.cuda.stream(self.stream):
            for (s, e), ev in zip(fg.ranges, fg.events):
                if e <= s:
                    continue
                self.stream.wait_event(ev)
                chunk = fg.flat[s:e]
                dist.all_reduce(chunk, op=op, group=self.group)
                if op != dist.ReduceOp.AVG:
                    chunk.div_(self.world)
                self.bytes += (e - s) * 4
            self.done = torch.cuda.Event()
            self.done.record(self.stream)

    def wait(self):
        """Makes the current stream wait for the reduced gradients; records events around the wait so the exposed time can be read."""
        if self.done is None:
            return
        cur = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cur)
        cur.wait_event(self.done)
        b.record(cur)
        self._wait = (a, b)
        self.done = None

    def exposed_ms(self):
        a, b = self._wait
        if a is None:
            return 0.0
        b.synchronize()
        return a.elapsed_time(b)


def allreduce_flat(fg: FlatGrads, process_group=None) -> int:
    """Non-overlapped variant (one collective over the whole buffer, then the average)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) <= 1:
        return 0
    dist.all_reduce(fg.flat, op=dist.ReduceOp.SUM, group=process_group)
    fg.flat.div_(dist.get_world_size(process_group))
    return fg.nbytes
