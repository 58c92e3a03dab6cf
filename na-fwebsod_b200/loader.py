"""Device-side input queue and result fetch for the head's training loop.

The reference never feeds the net synchronously: ``RoIDataLoader`` keeps host minibatches in a queue and
``enqueue_blobs`` copies them into a per-GPU ``BlobsQueue`` from loader threads while the net of the previous
iteration runs (detectron/roi_data/loader_wsl.py:98-127, 215-238; the net dequeues with ``DequeueBlobs``,
detectron/modeling/detector.py:85-105).  :class:`BlobsQueue` is that queue for this path: ``enqueue_blobs``
issues the host->device copies of one minibatch on a copy stream (pinned host tensors -> the copy engine runs
beside the compute stream's kernels), ``dequeue_blobs`` makes the compute stream wait for exactly that copy and
hands the blobs to ``WeblyHeadModel.FeedBlobs``.  :class:`LossFetcher` is the matching read side: the step's
loss blob is copied device->host into pinned memory behind the step (``workspace.FetchBlob`` in the reference's
``TrainingStats.UpdateIterStats``, detectron/utils/training_stats.py), and the host only blocks on the copy of
the PREVIOUS step, so it keeps one step of launches queued ahead of the GPU.

PyTorch is the plumbing here (streams, events, pinned memory); no arithmetic happens in this file.
"""
from __future__ import annotations

import collections

import torch


class BlobsQueue:
    """FIFO of minibatches resident on the device (capacity = minibatches that may be in flight).

    The device buffers are a RING of ``capacity + 1`` slots allocated on first use and reused for every later
    minibatch of the same shapes: the steady state performs no allocation at all (a per-step ``torch.empty`` on the
    copy stream + ``record_stream`` makes the caching allocator fall back to ``cudaMalloc`` every so often -- seen as
    one 50-130 ms step per few hundred in the round-1 end-to-end timings).  A slot is rewritten only after the compute
    stream has passed the dequeue that REPLACED it (an event recorded there)."""

    BLOB_NAMES = ("data_conv5", "rois", "obn_scores", "labels_oh", "roi_offsets")

    def __init__(self, model, capacity: int = 2, x_layout: str = "NCHW"):
        if capacity < 1:
            raise RuntimeError("BlobsQueue capacity must be >= 1")
        if model.device.type != "cuda":
            raise RuntimeError("BlobsQueue needs a CUDA model (libnawsod has no CPU path)")
        self.model, self.capacity, self.x_layout = model, capacity, x_layout
        self.device = model.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._ready = collections.deque()       # (slot index, blobs dict, event recorded behind the copies)
        self._slots = [dict() for _ in range(capacity + 1)]     # name -> device tensor, per ring slot
        self._free_after = [None] * (capacity + 1)              # compute-stream event after which the slot may be rewritten
        self._next_slot = 0
        self._in_use = None                     # slot currently fed to the model (dead once the next dequeue is enqueued)
        self.h2d_bytes = 0

    def __len__(self):
        return len(self._ready)

    def enqueue_blobs(self, data_conv5, rois, obn_scores, labels_oh=None, roi_offsets=None):
        """Start the host->device copy of one minibatch (host tensors; pinned ones copy asynchronously)."""
        if len(self._ready) >= self.capacity:
            raise RuntimeError("BlobsQueue is full (capacity %d): dequeue before enqueueing more" % self.capacity)
        host = dict(zip(self.BLOB_NAMES, (data_conv5, rois, obn_scores, labels_oh, roi_offsets)))
        slot = self._next_slot
        self._next_slot = (slot + 1) % len(self._slots)
        bufs = self._slots[slot]
        blobs = {}
        with torch.cuda.stream(self.copy_stream):
            if self._free_after[slot] is not None:
                self.copy_stream.wait_event(self._free_after[slot])     # the step that read this slot has been passed
                self._free_after[slot] = None
            for name, t in host.items():
                if t is None:
                    blobs[name] = None
                    continue
                if t.is_cuda:
                    raise RuntimeError("enqueue_blobs takes HOST tensors; %s is already on the device" % name)
                d = bufs.get(name)
                if d is None or d.shape != t.shape or d.dtype != t.dtype:
                    d = torch.empty(t.shape, dtype=t.dtype, device=self.device)     # first use of the slot (or a new shape)
                    d.record_stream(torch.cuda.current_stream(self.device))
                    bufs[name] = d
                d.copy_(t, non_blocking=True)
                blobs[name] = d
                self.h2d_bytes += t.numel() * t.element_size()
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready.append((slot, blobs, ev))

    def dequeue_blobs(self):
        """Feed the oldest minibatch to the model; the compute stream waits for its copy only."""
        if not self._ready:
            raise RuntimeError("BlobsQueue is empty: enqueue_blobs first")
        slot, blobs, ev = self._ready.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        if self._in_use is not None:            # everything that read the previous minibatch is already enqueued on `cur`
            done = torch.cuda.Event()
            done.record(cur)
            self._free_after[self._in_use] = done
        self._in_use = slot
        self.model.FeedBlobs(blobs["data_conv5"], blobs["rois"], blobs["obn_scores"], blobs["labels_oh"],
                             blobs["roi_offsets"], x_layout=self.x_layout)
        return blobs


class LossFetcher:
    """Device->host reads of per-step results with a lag: push() enqueues the copy behind the step, the host
    blocks (wait_lagged) only until the copy of `lag` steps ago has landed.  The pinned host buffers are a ring of
    ``lag + 2`` (allocated once per shape): no ``cudaHostAlloc`` in the steady state; values are cloned out."""

    def __init__(self, lag: int = 1):
        self.lag = lag
        self._pending = collections.deque()     # (pinned host tensor, event)
        self._ring, self._next = [None] * (lag + 2), 0
        self.values = []
        self.d2h_bytes = 0

    def push(self, t: torch.Tensor):
        h = self._ring[self._next]
        if h is None or h.shape != t.shape or h.dtype != t.dtype:
            h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            self._ring[self._next] = h
        self._next = (self._next + 1) % len(self._ring)
        h.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((h, ev))
        self.d2h_bytes += h.numel() * h.element_size()
        self.wait_lagged()

    def wait_lagged(self):
        while len(self._pending) > self.lag:
            h, ev = self._pending.popleft()
            ev.synchronize()
            self.values.append(h.clone())

    def wait_all(self):
        while self._pending:
            h, ev = self._pending.popleft()
            ev.synchronize()
            self.values.append(h.clone())
        return self.values
