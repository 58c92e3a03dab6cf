"""Multi-GPU parity of the data-parallel step against the ORACLE's reference schedule, on every GPU of the box
(world = 2, 4 or 8; skipped on a 1-GPU box).

The reference (detectron/modeling/optimizer_wsl.py:52-72, 96-137) sums every parameter gradient over the GPUs with
``NCCLAllreduce`` and then runs ``ACMWeightDecayMomentumSGDUpdate`` with ``gpu_num = NUM_GPUS`` on every replica
(ops/acm_weightdecay_momentum_sgd_op.h:79-84).  Here that schedule is evaluated ON THE CPU from the oracle alone --
per-rank oracle gradients of the whole head on the rank's own image, added in rank order, fed to the oracle's
restatement of the update op -- for three steps, and every exchange schedule of na-fwebsod_b200/dp.py
(``allreduce``, NCCL ``sharded``, peer-mapped ``p2p`` in pull and in push mode with the SM and the copy engines) must land on the
same parameters and momenta.  Unlike tests/test_gpu_zzzz_dp_2gpu.py, which compares the schedules with each other, a
bug common to all of them (bucket plan, slice ownership, the 1/gpu_num factor, bias hyper-parameters) fails here.

Also asserted, per schedule and WITHOUT gathering the master state first: after every step the state a forward pass
reads (operand shadow of the weights, fp32 masters of the biases) is bit-identical on all ranks -- the biases start
non-zero and the run is three steps long, so a rank training on stale biases outside its slice shows up.

Two bars.  (i) Against the oracle: the model runs the fp32 / TF32 path and the oracle uses its OWN ReLU pattern (it is not
conditioned on the device's, unlike tests/test_gpu_head.py), on a deliberately small head (K = 3136, 256 hidden units)
where a handful of flipped ReLU boundary elements weigh more than at full size; the ranks train on different labels, so
their gradients partly cancel in the sum while their rounding errors do not.  Measured on B200, identical for every
schedule: 2.5e-3 (world 2) and 6.0e-3 (world 8) after step 1, 2.5e-3 / 3.8e-3 after step 3; the bar is 1e-2 relative L2
per blob for parameter change and momenta.  A dropped, doubled or misrouted rank contribution is >= 1/world of a blob's
gradient: >= 0.1.  (ii) Between the schedules: every peer / NCCL variant must reproduce the reference schedule
(``allreduce``) to 5e-3 of each blob's own largest update (floored at 1e-3 of the largest update of any blob of its kind:
fc8d_b's update is pure rounding noise and differs by 100 % of itself between any two runs; measured 1e-3 on that floor, and
~1e-6 for the real blobs) -- the schedules differ in nothing but the order the ranks' fp32 gradients are added in and the
atomics' order inside the bias-gradient column sums."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STEPS, LR, MOM, WD = 3, 1e-2, 0.9, 5e-4
NCLS, CC, HD, R, MH, MW = 7, 64, 256, 256, 20, 25
# (schedule, engine of the peer copies, reduce-scatter mode of the peer exchange)
VARIANTS = (("allreduce", "-", "-"), ("sharded", "-", "-"), ("p2p", "sm", "pull"), ("p2p", "ce", "pull"), ("p2p", "sm", "push"),
            ("p2p", "ce", "push"))
TOL_FIRST, TOL_LAST, TOL_BETWEEN = 1e-2, 1e-2, 5e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_inputs(rank):
    from oracle import nawsod_oracle as O
    X = O.synth_conv5(1, CC, MH, MW, seed=10 + rank)
    rois = O.synth_rois(R, MH * 16, MW * 16, seed=20 + rank)
    obn = (np.random.default_rng(30 + rank).random((R, 1)) + 1).astype(np.float32)
    L = np.zeros((1, NCLS - 1), np.float32)
    L[0, rank % (NCLS - 1)] = 1
    return X, rois, obn, L


def _initial_params():
    from oracle import nawsod_oracle as O
    p = O.synth_params(NCLS - 1, CC * 49, HD, noise=True, seed=3)
    rng = np.random.default_rng(4)
    for k in p:
        if k.endswith("fc6_w"):
            p[k] = (p[k] * np.float32(np.sqrt(25088.0 / (CC * 49)))).astype(np.float32)
        if k.endswith("fc7_w"):
            p[k] = (p[k] * np.float32(np.sqrt(4096.0 / HD))).astype(np.float32)
        if k.endswith("_b"):                    # non-zero biases: a stale copy on a non-owner rank must be visible
            p[k] = (rng.standard_normal(p[k].shape) * 0.05).astype(np.float32)
    return p


def _is_bias(k):
    return k.endswith("_b")


def _oracle_schedule(world):
    """all-reduce (rank-order sum) + ACMWeightDecayMomentumSGDUpdate(gpu_num = world), from the oracle only."""
    from oracle import nawsod_oracle as O
    p = _initial_params()
    m = {k: np.zeros_like(v) for k, v in p.items()}
    inputs = [_rank_inputs(r) for r in range(world)]
    snaps = []
    for it in range(STEPS):
        total = None
        for r in range(world):
            X, rois, obn, L = inputs[r]
            g = O.head_forward_backward(X, rois, obn, L, p)["grads"]
            total = {k: v.astype(np.float32) for k, v in g.items()} if total is None else \
                {k: (total[k] + g[k]).astype(np.float32) for k in total}
        for k in p:                              # optimizer_wsl.py:106-123: weights wd, lr_mult 1; biases no decay, lr_mult 2
            m[k], p[k], _, _ = O.acm_sgd_update(total[k], m[k], LR, p[k], np.zeros_like(p[k]), momentum=MOM,
                                                weight_decay=0.0 if _is_bias(k) else WD, lr_mult=2.0 if _is_bias(k) else 1.0,
                                                gpu_num=world, iter_count=it)
        snaps.append(({k: v.copy() for k, v in p.items()}, {k: v.copy() for k, v in m.items()}))
    return snaps


def _to_reference_names(d):
    """oracle keys (noisy_fc6_w) -> the reference's blob names (_[noisy]_fc6_w; noisy_fc8c_w stays)."""
    out = {}
    for k, v in d.items():
        if k.startswith("noisy_fc6") or k.startswith("noisy_fc7"):
            out["_[noisy]_" + k[len("noisy_"):]] = v
        else:
            out[k] = v
    return out


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NAWSOD_COMM_SMS="16")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from nafwebsod_b200.heads import WeblyHeadModel
        from nafwebsod_b200.dp import DataParallelHead
        res = {}
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        X, rois, obn, L = _rank_inputs(rank)
        for sync, engine, rs in VARIANTS:
            os.environ["NAWSOD_P2P_ENGINE"] = engine if engine != "-" else "sm"
            os.environ["NAWSOD_P2P_RS"] = rs if rs != "-" else "pull"
            m = WeblyHeadModel(NCLS, CC, 7, HD, noise=True, dtype=torch.float32, device=dev)
            m.load_reference_params(_initial_params())
            m.UpdateWorkspaceLr(LR)
            dp = DataParallelHead(m, fc6_panels=4, sync=sync)
            assert dp.sync == sync
            m.FeedBlobs(t(X), t(rois), t(obn), t(L), x_layout="NCHW")
            nw = m.n_weights
            same = True
            entry = {"snaps": []}
            for it in range(STEPS):
                dp.step(dropout=False, momentum=MOM, weight_decay=WD)
                dp.flush()
                torch.cuda.synchronize()
                # forward-visible state, NO gather of the masters: weights' operand shadow + biases' fp32 masters
                visible = torch.cat([m.flat_lp[:nw].float(), m.flat_param[nw:]])
                ref = visible.clone()
                dist.broadcast(ref, src=0)
                same = same and bool(torch.equal(ref, visible))
                if it in (0, STEPS - 1):
                    bias_before_gather = m.flat_param[nw:].clone()
                    dp.gather_master_state()
                    torch.cuda.synchronize()
                    # the biases every rank trained on ARE the masters (nothing for the gather to repair)
                    same = same and bool(torch.equal(bias_before_gather, m.flat_param[nw:]))
                    if rank == 0:
                        entry["snaps"].append((
                            {k: v.detach().cpu().numpy().copy() for k, v in m.export_reference_params().items()},
                            {k: v for k, v in m.weights_file_blobs().items() if k.endswith("_momentum")}))
            if sync == "p2p":
                dp.exchange.check()
            entry["ranks_identical"] = same
            if rank == 0:
                entry["shadow_is_rounded_master"] = bool(torch.equal(
                    m.flat_lp[:nw], _round_tf32(m.flat_param[:nw])))
            res["%s/%s/%s" % (sync, engine, rs)] = entry
            del dp, m
        out[rank] = res
    finally:
        dist.destroy_process_group()


def _round_tf32(x):
    from nafwebsod_b200 import ops
    return ops.round_to_tf32(x.contiguous().view(1, -1)).view(-1)


def rel_l2(a, b, floor=0.0):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor, 1e-300))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_every_exchange_schedule_matches_the_oracle_reference_schedule():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = {r: out[r] for r in range(world)}
    snaps = _oracle_schedule(world)
    p0 = _to_reference_names(_initial_params())
    report, failures = {}, []
    for sync, engine, rs in VARIANTS:
        key = "%s/%s/%s" % (sync, engine, rs)
        for r in range(world):
            assert res[r][key]["ranks_identical"], "%s: rank %d's forward-visible state differs from rank 0's" % (key, r)
        e = res[0][key]
        assert e["shadow_is_rounded_master"], "%s: operand shadow is not the TF32-rounded master" % key
        report[key] = []
        for (step, tol), (p_gpu, m_gpu) in zip(((0, TOL_FIRST), (STEPS - 1, TOL_LAST)), e["snaps"]):
            p_ref, m_ref = _to_reference_names(snaps[step][0]), _to_reference_names(snaps[step][1])
            # fc8d_b's gradient vanishes analytically (the RoI-softmax gradient sums to zero over the RoIs): such a blob
            # is measured against 1e-3 of the largest blob of its kind instead of against its own (rounding-noise) norm
            floor_p = {b: 1e-3 * max(np.linalg.norm(p_ref[k] - p0[k]) for k in p_ref if _is_bias(k) == b) for b in (False, True)}
            floor_m = {b: 1e-3 * max(np.linalg.norm(m_ref[k]) for k in p_ref if _is_bias(k) == b) for b in (False, True)}
            worst = 0.0
            for k in p_ref:
                ep = rel_l2(p_gpu[k] - p0[k], p_ref[k] - p0[k], floor_p[_is_bias(k)])
                em = rel_l2(m_gpu[k + "_momentum"], m_ref[k], floor_m[_is_bias(k)])
                worst = max(worst, ep, em)
                if ep > tol or em > tol:
                    failures.append("%s after step %d: %s parameter change off by %.3g, momentum by %.3g (bar %.0e)" % (key, step + 1, k, ep, em, tol))
            report[key].append("%.2e" % worst)
    print("world %d, worst relative L2 error vs the oracle schedule after step 1 / step %d: %s" % (world, STEPS, report))
    # (ii) between the schedules, final state
    base_p, base_m = res[0]["allreduce/-/-"]["snaps"][-1]
    between = {}
    for sync, engine, rs in VARIANTS[1:]:
        key = "%s/%s/%s" % (sync, engine, rs)
        p_v, m_v = res[0][key]["snaps"][-1]
        worst = 0.0
        # scale: the largest update of any blob of the same kind (fc8d_b's own update is rounding noise, see above)
        kind_scale = {b: max(np.abs(base_m[k + "_momentum"]).max() for k in base_p if _is_bias(k) == b) for b in (False, True)}
        for k in base_p:
            scale = max(np.abs(base_m[k + "_momentum"]).max(), 1e-3 * kind_scale[_is_bias(k)])
            worst = max(worst, np.abs(p_v[k] - base_p[k]).max() / scale, np.abs(m_v[k + "_momentum"] - base_m[k + "_momentum"]).max() / scale)
        between[key] = "%.1e" % worst
        if worst > TOL_BETWEEN:
            failures.append("%s differs from the reference schedule by %.3g of the largest update (bar %.0e)" % (key, worst, TOL_BETWEEN))
    print("world %d, largest deviation from the all-reduce schedule (fraction of the blob's largest update): %s" % (world, between))
    assert not failures, "\n".join(failures)
