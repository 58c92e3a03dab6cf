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
    """FIFO of minibatches resident on the device (capacity = minibatches that may be in flight)."""

    BLOB_NAMES = ("data_conv5", "rois", "obn_scores", "labels_oh", "roi_offsets")

    def __init__(self, model, capacity: int = 2, x_layout: str = "NCHW"):
        if capacity < 1:
            raise RuntimeError("BlobsQueue capacity must be >= 1")
        if model.device.type != "cuda":
            raise RuntimeError("BlobsQueue needs a CUDA model (libnawsod has no CPU path)")
        self.model, self.capacity, self.x_layout = model, capacity, x_layout
        self.device = model.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._ready = collections.deque()       # (blobs dict, event recorded behind the copies)
        self._retired = collections.deque()     # events: compute-stream position after which a minibatch's blobs are dead
        self._in_use = None                     # blobs currently fed to the model (kept alive until the next dequeue)
        self.h2d_bytes = 0

    def __len__(self):
        return len(self._ready)

    def enqueue_blobs(self, data_conv5, rois, obn_scores, labels_oh=None, roi_offsets=None):
        """Start the host->device copy of one minibatch (host tensors; pinned ones copy asynchronously)."""
        if len(self._ready) >= self.capacity:
            raise RuntimeError("BlobsQueue is full (capacity %d): dequeue before enqueueing more" % self.capacity)
        host = dict(zip(self.BLOB_NAMES, (data_conv5, rois, obn_scores, labels_oh, roi_offsets)))
        cur = torch.cuda.current_stream(self.device)
        blobs = {}
        with torch.cuda.stream(self.copy_stream):
            for name, t in host.items():
                if t is None:
                    blobs[name] = None
                    continue
                if t.is_cuda:
                    raise RuntimeError("enqueue_blobs takes HOST tensors; %s is already on the device" % name)
                d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                d.copy_(t, non_blocking=True)
                d.record_stream(cur)            # allocated on the copy stream, consumed on the compute stream
                blobs[name] = d
                self.h2d_bytes += t.numel() * t.element_size()
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready.append((blobs, ev))

    def dequeue_blobs(self):
        """Feed the oldest minibatch to the model; the compute stream waits for its copy only."""
        if not self._ready:
            raise RuntimeError("BlobsQueue is empty: enqueue_blobs first")
        blobs, ev = self._ready.popleft()
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._in_use = blobs
        self.model.FeedBlobs(blobs["data_conv5"], blobs["rois"], blobs["obn_scores"], blobs["labels_oh"],
                             blobs["roi_offsets"], x_layout=self.x_layout)
        return blobs


class LossFetcher:
    """Device->host reads of per-step results with a lag: push() enqueues the copy behind the step, the host
    blocks (wait_lagged) only until the copy of `lag` steps ago has landed."""

    def __init__(self, lag: int = 1):
        self.lag = lag
        self._pending = collections.deque()     # (pinned host tensor, event)
        self.values = []
        self.d2h_bytes = 0

    def push(self, t: torch.Tensor):
        h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
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
            self.values.append(h)

    def wait_all(self):
        while self._pending:
            h, ev = self._pending.popleft()
            ev.synchronize()
            self.values.append(h)
        return self.values
