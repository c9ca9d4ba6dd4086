"""Host <-> device staging for the hot path's drivers.

The reference moves every batch with a blocking `.to(device)` on the compute stream (trainer/trainer.py:186-187,
trainer/r_evaler.py:88-90) and reads results back the same way.  `HostPrefetcher` keeps those copies off the
critical path: pinned host batches are copied on a dedicated H2D stream into a two-slot device buffer (batch i+1
travels while batch i is computed) and results leave on a dedicated D2H stream (result i travels while batch i+1 is
computed).  Ordering is by CUDA events only; nothing here synchronises the host.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class HostPrefetcher:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("HostPrefetcher stages pinned host memory onto a CUDA device")
        self.depth = depth
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._bufs: List[Optional[List[torch.Tensor]]] = [None] * depth
        self._ready: List[Optional[torch.cuda.Event]] = [None] * depth
        self._free: List[Optional[torch.cuda.Event]] = [None] * depth
        self._n_put = 0
        self._n_get = 0
        self._cur = -1

    def put(self, *host: torch.Tensor) -> None:
        """Enqueue the H2D copy of one (pinned) host batch; returns immediately."""
        if self._n_put - self._n_get >= self.depth:
            raise RuntimeError("HostPrefetcher: more batches in flight than buffer slots")
        s = self._n_put % self.depth
        self._n_put += 1
        bufs = self._bufs[s]
        if bufs is None or len(bufs) != len(host) or any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host)):
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host]   # allocated on the compute stream
            self._bufs[s] = bufs
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._free[s] = ev
        with torch.cuda.stream(self.h2d):
            if self._free[s] is not None:
                self.h2d.wait_event(self._free[s])            # the previous user of this slot has been computed
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.h2d)
            self._ready[s] = ev

    def get(self) -> Sequence[torch.Tensor]:
        """Device tensors of the oldest batch in flight; the current stream waits for its copy (the host does not)."""
        if self._n_get >= self._n_put:
            raise RuntimeError("HostPrefetcher.get() without a matching put()")
        s = self._n_get % self.depth
        self._n_get += 1
        torch.cuda.current_stream(self.device).wait_event(self._ready[s])
        self._cur = s
        return self._bufs[s]

    def done(self) -> None:
        """Call once the work reading the batch returned by the last get() is enqueued: frees its slot for reuse."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[self._cur] = ev

    def download(self, dev_t: torch.Tensor, host_t: torch.Tensor) -> None:
        """Enqueue the D2H copy of a result on the D2H stream (ordered after the work enqueued so far)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        dev_t.record_stream(self.d2h)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(ev)
            host_t.copy_(dev_t, non_blocking=True)

    def join(self) -> None:
        """Make the current stream wait for every copy enqueued so far (no host synchronisation)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.h2d)
        cur.wait_stream(self.d2h)
