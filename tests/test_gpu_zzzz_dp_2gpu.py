"""Multi-GPU parity of the data-parallel step (needs >= 2 GPUs; skipped on a 1-GPU box).

Both exchange schedules of na-fwebsod_b200/dp.py run the same two training steps from the same
initial state on rank-specific images: the sharded schedule (reduce-scatter -> SGD on the rank's
slice -> all-gather of the bf16 operands, overlapped with the fc6 weight-gradient panels), over NCCL
and over the peer-mapped copy-engine path (sync="p2p"), must leave the parameters, momenta and GEMM operands the reference schedule leaves (bucketed
all-reduce + full ACMWeightDecayMomentumSGDUpdate on every rank,
detectron/modeling/optimizer_wsl.py:52-137).  Tolerance: the two collectives may add the ranks'
fp32 gradients in a different order and the bias-gradient column sums use atomics (run-to-run
order), so updated parameters may differ by fp32 rounding noise (<= 1e-3 of the largest update;
a dropped or doubled rank contribution would be O(1) of it); ranks must agree bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NAWSOD_COMM_SMS="16")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from nafwebsod_b200.heads import WeblyHeadModel
        from nafwebsod_b200.dp import DataParallelHead
        from oracle import nawsod_oracle as O
        res = {}
        for variant in ("allreduce", "sharded", "p2p"):
            sync = variant
            m = WeblyHeadModel(7, 64, 7, 256, noise=True, dtype=torch.bfloat16, device=dev)
            g = torch.Generator(device=dev).manual_seed(5)
            m.flat_param[:m.n_weights].normal_(0.0, 0.02, generator=g)
            m.sync_shadow()
            m.UpdateWorkspaceLr(1e-2)
            dp = DataParallelHead(m, fc6_panels=4, sync=sync)
            dp.broadcast_parameters()
            X = O.synth_conv5(1, 64, 20, 25, seed=10 + rank)
            rois = O.synth_rois(256, 320, 400, seed=20 + rank)
            obn = (np.random.default_rng(30 + rank).random(256) + 1).astype(np.float32)
            L = np.zeros((1, 6), np.float32)
            L[0, rank % 6] = 1
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            m.FeedBlobs(t(X), t(rois), t(obn), t(L), x_layout="NCHW")
            for it in range(2):
                dp.step(dropout_seed=it + 1)
            dp.gather_master_state()
            torch.cuda.synchronize()
            if sync == "p2p":
                dp.exchange.check()
            res[variant] = (m.flat_param.cpu().numpy().copy(), m.flat_mom.cpu().numpy().copy(),
                            m.flat_lp.float().cpu().numpy().copy(), m.blobs["loss"].cpu().numpy().copy())
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_exchange_matches_allreduce_schedule():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for sync in ("allreduce", "sharded", "p2p"):
        for a, b in zip(r0[sync][:3], r1[sync][:3]):
            assert np.array_equal(a, b), "%s: ranks disagree" % sync
    pa, ma, la, _ = r0["allreduce"]
    ps, ms, ls, _ = r0["sharded"]
    upd = np.abs(ma).max()
    assert upd > 0
    dp_, dm_ = np.abs(pa - ps), np.abs(ma - ms)
    assert dp_.max() <= 1e-3 * upd, (dp_.max(), upd, int(dp_.argmax()), int((dp_ > 1e-5 * upd).sum()), pa.size)
    assert dm_.max() <= 1e-3 * upd, (dm_.max(), upd, int(dm_.argmax()), int((dm_ > 1e-5 * upd).sum()))
    assert np.abs(la - ls).max() <= 2e-2 * np.abs(la).max()     # bf16 operands: at most one rounding step apart
    # each rank saw different images: the losses must differ between ranks but match between schedules
    assert np.allclose(r0["allreduce"][3], r0["sharded"][3], rtol=1e-3)
    # the peer-mapped exchange: same bar against the reference schedule (its owner-side sum runs in rank order)
    pp, mp_, lp, _ = r0["p2p"]
    assert np.abs(pa - pp).max() <= 1e-3 * upd and np.abs(ma - mp_).max() <= 1e-3 * upd
    assert np.abs(la - lp).max() <= 2e-2 * np.abs(la).max()
    assert np.allclose(r0["allreduce"][3], r0["p2p"][3], rtol=1e-3)
