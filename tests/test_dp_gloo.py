"""world_size-2 gloo tests (CPU) of the data-parallel host logic: image sharding, the bucket plan,
and both gradient-exchange schedules of na-fwebsod_b200/dp.py.  The sharded schedule
(reduce-scatter -> SGD on the rank's slice -> all-gather of the GEMM operands) must produce the
parameters the reference's schedule produces (all-reduce, then the full
ACMWeightDecayMomentumSGDUpdate on every rank; detectron/modeling/optimizer_wsl.py:52-137), with
the update arithmetic taken from the oracle's restatement of the reference op."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nawsod_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


# a miniature of the flat layout [W6 | other weights | biases] (heads.WeblyHeadModel._alloc_params)
W6_ROWS, W6_COLS = 16, 24
N_W6 = W6_ROWS * W6_COLS
N_WEIGHTS = N_W6 + 128
N_TOTAL = N_WEIGHTS + 64
LR, MOM, WD = 1e-2, 0.9, 5e-4


def _grads(rank, step):
    rng = np.random.default_rng(100 * step + rank)
    return rng.standard_normal(N_TOTAL).astype(np.float32)


def _reference(world, steps):
    """all-reduce + the full SGD op, in one process."""
    rng = np.random.default_rng(7)
    p = rng.standard_normal(N_TOTAL).astype(np.float32)
    m = np.zeros(N_TOTAL, np.float32)
    for it in range(steps):
        g = np.zeros(N_TOTAL, np.float32)
        for r in range(world):                       # gloo sums in rank order
            g = (g + _grads(r, it)).astype(np.float32)
        for lo, hi, wd, mult in ((0, N_WEIGHTS, WD, 1.0), (N_WEIGHTS, N_TOTAL, 0.0, 2.0)):
            m[lo:hi], p[lo:hi], _, _ = O.acm_sgd_update(g[lo:hi], m[lo:hi], LR, p[lo:hi], np.zeros(hi - lo, np.float32),
                                                        momentum=MOM, weight_decay=wd, lr_mult=mult, gpu_num=world, iter_count=it)
    return p, m


def _worker(rank, world, port, sync, panels, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nafwebsod_b200 import dp
        rng = np.random.default_rng(7)
        p = torch.from_numpy(rng.standard_normal(N_TOTAL).astype(np.float32))
        m = torch.zeros(N_TOTAL)
        g = torch.zeros(N_TOTAL)
        shadow = p.clone()                            # stands for the bf16 / TF32 GEMM-operand copy
        plan = dp.bucket_plan(N_W6, W6_ROWS, W6_COLS, N_WEIGHTS, N_TOTAL, panels, align_rows=4)
        assert sum(n for _, n, _ in plan) == N_TOTAL and plan[-1][2] == "biases"
        state = {"it": 0}

        def update(off, length, tag, so, sn):
            bias = tag == "biases"
            mm, pp, _, _ = O.acm_sgd_update(g[so:so + sn].numpy(), m[so:so + sn].numpy(), LR, p[so:so + sn].numpy(),
                                            np.zeros(sn, np.float32), momentum=MOM, weight_decay=0.0 if bias else WD,
                                            lr_mult=2.0 if bias else 1.0, gpu_num=world, iter_count=state["it"])
            m[so:so + sn] = torch.from_numpy(mm)
            p[so:so + sn] = torch.from_numpy(pp)
            shadow[so:so + sn] = p[so:so + sn]

        ex = dp.GradientExchange(g, shadow, sharded=(sync == "sharded"), update_fn=update)
        for it in range(3):
            state["it"] = it
            g.copy_(torch.from_numpy(_grads(rank, it)))
            for off, n, tag in plan:                  # the order the backward pass completes the buckets
                ex.launch(off, n, tag)
            ex.finish()
            if sync == "allreduce":
                update(0, N_WEIGHTS, "weights", 0, N_WEIGHTS)
                update(N_WEIGHTS, N_TOTAL - N_WEIGHTS, "biases", N_WEIGHTS, N_TOTAL - N_WEIGHTS)
        # what the NEXT forward pass would read on this rank, with no gather in between: the operand shadow of the
        # weights and the fp32 MASTERS of the biases (heads.py reads b6 / b7 / b8 from flat_param)
        visible = np.concatenate([shadow[:N_WEIGHTS].numpy(), p[N_WEIGHTS:].numpy()]).copy()
        if sync == "sharded":
            # operands are complete everywhere; masters only on their owner until gathered
            for off, n, tag in plan:
                if not dp.bucket_is_sliced(n, tag, world):
                    continue                          # replicated bucket (biases, non-divisible): every rank updated all of it
                so, sn = dp.rank_slice(off, n, world, rank)
                for flat in (p, m):
                    dist.all_gather_into_tensor(flat[off:off + n], flat[so:so + sn])
        out[rank] = (p.numpy().copy(), m.numpy().copy(), shadow.numpy().copy(), visible)
    finally:
        dist.destroy_process_group()


def _run(sync, panels, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), sync, panels, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


@pytest.mark.parametrize("sync,panels", [("sharded", 4), ("sharded", 1), ("allreduce", 2)])
def test_exchange_matches_reference_schedule(sync, panels):
    res = _run(sync, panels)
    p_ref, m_ref = _reference(2, 3)
    for p, m, shadow, visible in res:
        # the same fp32 operations in the same order: bit-exact
        assert np.array_equal(p, p_ref)
        assert np.array_equal(m, m_ref)
        assert np.array_equal(shadow, p_ref)
        # ... and already BEFORE any gather of the master state, everything a forward pass reads (operand shadow of
        # the weights, fp32 masters of the biases) is the reference schedule's on every rank: no stale biases
        assert np.array_equal(visible, p_ref)
    assert np.array_equal(res[0][0], res[1][0])


@pytest.mark.parametrize("world,panels", [(4, 4), (3, 2)])
def test_sharded_exchange_other_world_sizes(world, panels):
    """4 ranks (every bucket divides: reduce-scatter / all-gather path) and 3 ranks (the 128- and 64-element
    buckets do not divide: whole-bucket all-reduce + redundant update path).  gloo's reduction order over more
    than two ranks is not the sequential rank order of the one-process reference, so sums may differ in the
    last bit: parameters agree to 1e-6 relative, and bit-exactly ACROSS ranks (replicas must not diverge)."""
    res = _run("sharded", panels, world=world)
    p_ref, m_ref = _reference(world, 3)
    for p, m, shadow, visible in res:
        np.testing.assert_allclose(p, p_ref, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m, m_ref, rtol=1e-5, atol=1e-7)
        assert np.array_equal(shadow, p)
        assert np.array_equal(visible, p)             # forward-visible state current on every rank without a gather
    for r in range(1, world):
        assert np.array_equal(res[0][0], res[r][0]) and np.array_equal(res[0][2], res[r][2])


def test_shard_images_and_plan():
    from nafwebsod_b200 import dp
    assert dp.shard_images(16, 8, 3) == [6, 7]
    with pytest.raises(RuntimeError):
        dp.shard_images(5, 2, 0)
    # the real layout: NA head, 20 classes
    n_w6 = 8192 * 25088
    n_weights = n_w6 + 2 * 4096 * 4096 + 2 * 40 * 4096
    n_total = n_weights + 8192 + 2 * 4096 + 2 * 64
    plan = dp.bucket_plan(n_w6, 8192, 25088, n_weights, n_total, 4)
    covered = 0
    for off, n, tag in plan:
        assert off == covered, "buckets must tile the flat buffer in order"
        covered += n
        for world in (2, 4, 8):
            so, sn = dp.rank_slice(off, n, world, world - 1)
            assert so + sn == off + n and sn % 4 == 0      # 16-byte aligned slices for the float4 SGD kernel
            # ... and for the bf16 operand shadow (8 elements): every bucket of the real layout is exchanged in slices
            assert dp.slices_aligned(n, world)
    assert covered == n_total
    with pytest.raises(RuntimeError):
        dp.rank_slice(0, 10, 4, 0)
    # an unpadded bias block (8192 + 8192 + 80 floats) splits evenly over 4 and 8 ranks but into slices that do not
    # start on 16-byte boundaries of the bf16 shadow: such a bucket takes the whole-bucket all-reduce path
    assert dp.slices_aligned(16464, 2) and not dp.slices_aligned(16464, 4) and not dp.slices_aligned(16464, 8)
    assert not dp.slices_aligned(10, 4)


# ---------------------------------------------------------------------------------------------------------------
# P2PExchange.self_test(): the dry run that lets sync="auto" fall back to NCCL on a world size the peer path was
# never run on.  The transport (peer-mapped memory + flag kernels) needs CUDA; here it is replaced by gloo
# collectives with the same contract (launch: the owner's update_fn sees the W contributions for its slice in rank
# order and its operand slice is published to every rank; finish: everything has landed), so that the CHECK logic
# and the cross-rank agreement are what is tested.
# ---------------------------------------------------------------------------------------------------------------
def _selftest_worker(rank, world, port, corrupt_rank, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nafwebsod_b200 import dp

        class GlooTransport(dp.P2PExchange):
            def __init__(self, flat, outbuf, plan):
                self.flat, self.out, self.group = flat, outbuf, None
                self.world, self.rank = world, rank
                self.plan = list(plan)
                self.update_fn, self.timeout_ms = None, 20000
                self.status = torch.zeros(1, dtype=torch.int32)
                self.seq = 0

            def launch(self, offset, length, tag):
                if not dp.bucket_is_sliced(length, tag, world):           # replicated (the biases): everyone gets everything
                    parts = [torch.empty(length) for _ in range(world)]
                    dist.all_gather(parts, self.flat[offset: offset + length].clone())
                    self.update_fn(offset, length, tag, offset, length, parts)
                    return
                n = length // world
                mine = [self.flat[offset + k * n: offset + (k + 1) * n].clone() for k in range(world)]
                if rank == corrupt_rank and tag == "small_weights":
                    mine[(rank + 1) % world][3] += 1.0                     # one wrong element sent to one owner
                got = None
                for k in range(world):                                     # scatter leg: owner k receives every rank's part
                    parts = [torch.empty(n) for _ in range(world)] if rank == k else None
                    dist.gather(mine[k], parts, dst=k)
                    if rank == k:
                        got = parts
                so = offset + rank * n
                self.update_fn(offset, length, tag, so, n, got)
                outs = [torch.empty(n, dtype=self.out.dtype) for _ in range(world)]
                dist.all_gather(outs, self.out[so: so + n].clone())        # operand leg
                for k in range(world):
                    self.out[offset + k * n: offset + (k + 1) * n] = outs[k]

            def finish(self):
                pass

        n_total = 3 * 64 * world
        plan = [(0, 64 * world, "fc6_panel"), (64 * world, 64 * world, "small_weights"), (128 * world, 64 * world, "biases")]
        ex = GlooTransport(torch.zeros(n_total), torch.zeros(n_total, dtype=torch.bfloat16), plan)
        marker = lambda *a: None
        ex.update_fn = marker
        ok, why = ex.self_test(timeout_ms=1000)
        assert ex.update_fn is marker and ex.timeout_ms == 20000            # restored
        assert float(ex.flat.abs().sum()) == 0.0                            # the scratch gradients are cleared again
        out[rank] = (ok, why)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,corrupt_rank", [(2, -1), (3, -1), (3, 1)])
def test_p2p_self_test_agrees_across_ranks(world, corrupt_rank):
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_selftest_worker, args=(world, _free_port(), corrupt_rank, out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    if corrupt_rank < 0:
        assert all(ok and why == "ok" for ok, why in res), res
    else:
        assert not any(ok for ok, _ in res), res                            # every rank falls back, not only the one that saw it
        assert any("data checks failed" in why for _, why in res), res
