"""BlobsQueue / LossFetcher (nafwebsod_b200/loader.py): feeding the head through the device-side
queue (host->device copies on a copy stream, prefetched one step ahead, lagged loss reads) must give
the results of feeding it synchronously -- the queue moves bytes, it does no arithmetic.  (Compared to fp32
rounding noise rather than bit for bit: bias-gradient column sums and the MIL partial sums use atomics whose
order may differ between two runs; a stale or torn minibatch would change the losses at O(1).)"""
import numpy as np
import pytest
import torch

from oracle import nawsod_oracle as O

pytestmark = pytest.mark.gpu


def _batches(n, Cc, Hh, Ww, R, ncls):
    out = []
    for k in range(n):
        X = O.synth_conv5(1, Cc, Hh, Ww, seed=100 + k)
        rois = O.synth_rois(R + 8 * k, Hh * 16, Ww * 16, seed=200 + k)        # ragged: R differs per minibatch
        obn = (np.random.default_rng(300 + k).random(rois.shape[0]) + 1).astype(np.float32)
        L = np.zeros((1, ncls - 1), np.float32)
        L[0, k % (ncls - 1)] = 1
        offs = np.asarray([0, rois.shape[0]], np.int32)
        out.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (X, rois, obn, L, offs)))
    return out


def _model(ncls, Cc, Hd):
    from nafwebsod_b200.heads import WeblyHeadModel
    m = WeblyHeadModel(ncls, Cc, 7, Hd, dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(11)
    m.flat_param[:m.n_weights].normal_(0.0, 0.05, generator=g)
    m.sync_shadow()
    m.UpdateWorkspaceLr(1e-2)
    return m


def test_queue_feed_matches_synchronous_feed():
    from nafwebsod_b200.dp import DataParallelHead
    from nafwebsod_b200.loader import BlobsQueue, LossFetcher
    ncls, Cc, Hd = 7, 64, 256
    batches = _batches(5, Cc, 20, 25, 64, ncls)

    # synchronous: FeedBlobs from device copies, blocking loss read
    m = _model(ncls, Cc, Hd)
    dp = DataParallelHead(m)
    ref = []
    for i, b in enumerate(batches):
        m.FeedBlobs(*[x.cuda() for x in b], x_layout="NCHW")
        ref.append(dp.step(dropout_seed=i + 1)["loss"].cpu().clone())
    dp.flush()
    p_ref = m.flat_param.cpu().clone()

    # queued: prefetch one minibatch ahead, lagged loss fetch
    m = _model(ncls, Cc, Hd)
    dp = DataParallelHead(m)
    q, fetch = BlobsQueue(m, capacity=2, x_layout="NCHW"), LossFetcher(lag=1)
    q.enqueue_blobs(*batches[0])
    for i in range(len(batches)):
        q.dequeue_blobs()
        if i + 1 < len(batches):
            q.enqueue_blobs(*batches[i + 1])
        fetch.push(dp.step(dropout_seed=i + 1)["loss"])
        assert len(fetch.values) == i                       # the host has seen every loss but the newest
    got = fetch.wait_all()
    dp.flush()
    torch.cuda.synchronize()
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-8), (a, b)
    assert len({tuple(r.flatten().tolist()) for r in ref}) == len(ref)      # the minibatches really differ
    dpar = (m.flat_param.cpu() - p_ref).abs().max().item()
    assert dpar <= 1e-6 * p_ref.abs().max().item(), dpar
    assert q.h2d_bytes == sum(t.numel() * t.element_size() for b in batches for t in b)
    assert fetch.d2h_bytes == sum(r.numel() * r.element_size() for r in ref)


def test_queue_errors():
    from nafwebsod_b200.loader import BlobsQueue
    m = _model(7, 64, 256)
    q = BlobsQueue(m, capacity=1)
    with pytest.raises(RuntimeError):
        q.dequeue_blobs()                                   # empty
    b = _batches(1, 64, 20, 25, 32, 7)[0]
    q.enqueue_blobs(*b)
    with pytest.raises(RuntimeError):
        q.enqueue_blobs(*b)                                 # full
    with pytest.raises(RuntimeError):
        BlobsQueue(m, capacity=1).enqueue_blobs(b[0].cuda(), *b[1:])   # device tensor where a host one is due
