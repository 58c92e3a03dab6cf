"""Data-parallel training step of the head: one process per GPU, images sharded by rank, weights
replicated, ONE gradient exchange per step.

The reference (detectron/modeling/optimizer_wsl.py:18-137) issues one ``NCCLAllreduce`` (sum, in
place) per parameter blob after the whole backward graph (``_add_allreduce_graph``, :52-72), then
runs the identical ``ACMWeightDecayMomentumSGDUpdate`` on every replica; the
``1/(iter_size*gpu_num)`` averaging lives in that op (acm_weightdecay_momentum_sgd_op.h:79-84).

Here the same arithmetic is scheduled for NVLink 5 / NVSwitch (``sync="sharded"``, the default):

    all-reduce(g); every rank: SGD(all of p)          (reference)
 == reduce-scatter(g); rank k: SGD(slice k of p); all-gather(p)

with the parameters' gradients in one flat float32 buffer cut into ordered buckets
(fc6 weight-gradient row panels, then the fc7/fc8 weights, then the biases).  Each bucket is split
evenly over the ranks, so per GPU the exchange moves (W-1)/W * 4 B/param of fp32 gradients out and
(W-1)/W * 2 B/param of updated bf16 GEMM operands back (instead of 2 * (W-1)/W * 4 B/param for an
all-reduce), and the SGD update touches 1/W of the 239 M parameters per rank.  The bucket pipeline
``reduce-scatter -> SGD(slice) -> all-gather(shadow)`` runs on a high-priority side stream and is
launched as soon as the bucket's producer GEMMs are enqueued: the fc6 weight gradient (86 % of the
bytes) is produced in row panels right after the activation-gradient chain, so the transfer of
panel p overlaps the tensor-core GEMM of panel p+1 and the fc7/fc8 weight-gradient GEMMs.  While a
transfer may be in flight the persistent GEMMs leave ``comm_sms`` SMs to the NCCL kernels
(``gemm_max_ctas`` tuning) -- a persistent CTA-per-SM grid would otherwise serialise behind them.
The compute stream joins the side stream only at the start of the next step (or ``flush()``).
On ONE GPU the same pipeline runs without a collective: each bucket's SGD update (HBM-bound) goes to
the side stream behind its producer GEMM and overlaps the remaining tensor-bound weight-gradient GEMMs.

``sync="p2p"`` is the same pipeline with the collectives replaced by the box's own hardware paths:
every rank maps its peers' gradient / flag / operand (and, in push mode, staging) buffers (CUDA IPC over NVSwitch).  Reduce-scatter
leg: the owner's SGD kernel reads the other ranks' gradient slices IN PLACE over NVLink and sums the W contributions in rank
order -- deterministic -- while it updates (``nawsod_sgd_update_reduce``; "pull", the default beyond two ranks), or the copy
engines push the slices into the owner's staging area first ("push").  Operand leg: the updated bf16 slice goes back to every
peer with the COPY ENGINES (no SM is taken from the GEMMs).  Cross-GPU ordering is a sequence number per (bucket, rank)
published / awaited by one-warp kernels (``nawsod_p2p_signal`` / ``nawsod_p2p_wait``, watchdog-bounded); the next step's fc6
meets its weight panels inside the GEMM (``nawsod_fc_fwd_gated``), and consecutive buckets alternate between two update streams.

``sync="allreduce"`` keeps the reference's schedule (bucketed all-reduce, full SGD on every rank)
for comparison.  torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is the plumbing;
the path has no other collective.  Inference shards by image with no collective (replicas only).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_images(num_images_total: int, world_size: int, rank: int):
    """Rank g gets images {g*B .. g*B+B-1} with all their RoIs; B must divide evenly."""
    if num_images_total % world_size:
        raise RuntimeError("global image batch %d is not divisible by world size %d" % (num_images_total, world_size))
    b = num_images_total // world_size
    return list(range(rank * b, (rank + 1) * b))


def n_weights_w6_padded(n: int, align: int = 64) -> int:
    return (n + align - 1) // align * align


def bucket_plan(n_weights_w6: int, w6_rows: int, w6_cols: int, n_weights: int, n_total: int, panels: int, align_rows: int = 256,
                n_bias_fc6: int = 0):
    """Exchange buckets (offset, length, tag) tiling the flat gradient buffer laid out as [W6 | other weights | biases]:
    fc6 row panels, the fc7/fc8 weights, then the biases -- in two buckets when ``n_bias_fc6`` (the padded length of b6,
    which leads the bias block) is given: "biases_fc6" is complete with the last fc6 panel and is, with the panels, all
    the NEXT step's fc6 needs, so that step can start while "small_weights" and the remaining "biases" are still in
    flight.  Every element is covered exactly once."""
    assert n_weights_w6 == w6_rows * w6_cols
    plan = []
    step = ((w6_rows + panels - 1) // panels + align_rows - 1) // align_rows * align_rows
    for r0 in range(0, w6_rows, step):
        r1 = min(w6_rows, r0 + step)
        plan.append((r0 * w6_cols, (r1 - r0) * w6_cols, "fc6_panel"))
    w6p = n_weights_w6_padded(n_weights_w6)
    plan.append((w6p, n_weights - w6p, "small_weights"))
    if 0 < n_bias_fc6 < n_total - n_weights:
        plan.append((n_weights, n_bias_fc6, "biases_fc6"))
        plan.append((n_weights + n_bias_fc6, n_total - n_weights - n_bias_fc6, "biases"))
    else:
        plan.append((n_weights, n_total - n_weights, "biases"))
    return plan


def rank_slice(offset: int, length: int, world: int, rank: int):
    """Rank's contiguous share of a bucket (the reduce-scatter / all-gather chunk)."""
    if length % world:
        raise RuntimeError("bucket of %d elements is not divisible by world size %d" % (length, world))
    n = length // world
    return offset + rank * n, n


def slices_aligned(length: int, world: int) -> bool:
    """A bucket can be exchanged in per-rank slices only if it splits evenly into slices of a multiple of 8 elements:
    the update kernel moves 16-byte vectors of the fp32 buffers AND of the bf16 operand shadow, so every slice must
    start on a 16-byte boundary in both.  Other buckets take the whole-bucket all-reduce with a redundant update on
    every rank."""
    return length % world == 0 and (length // world) % 8 == 0


def bucket_is_sliced(length: int, tag: str, world: int) -> bool:
    """Whether a bucket's update is split over the ranks (reduce-scatter -> update of the owned slice -> all-gather of
    the GEMM operands) or REPLICATED (every rank reduces all W contributions and updates the whole bucket).

    The "biases" bucket is always replicated, whatever its alignment: the forward pass reads b6 / b7 / b8 from the
    fp32 MASTERS (heads.py: ``self.p["b*"]`` are views of flat_param), and the sliced schedules only send the
    GEMM-operand shadow back -- a rank would keep training on stale copies of every bias outside its own slice.  The
    bucket is 66 KB; replicating its update costs nothing and keeps masters, momenta and shadow current everywhere."""
    return not tag.startswith("biases") and slices_aligned(length, world)


class GradientExchange:
    """Bucket pipeline on a side stream (CUDA) or inline (CPU tensors / gloo).

    sharded:   reduce_scatter(grad bucket) -> update_fn(bucket, own slice) -> all_gather(out bucket)
    allreduce: all_reduce(grad bucket)                       (update happens once, after finish())
    """

    def __init__(self, flat_grad: torch.Tensor, flat_out, group=None, sharded: bool = True, update_fn=None):
        self.flat, self.out = flat_grad, flat_out
        self.group = group
        self.local = not dist.is_initialized()       # one GPU: no collective, the pipeline still hides the update
        self.world = 1 if self.local else dist.get_world_size(group)
        self.rank = 0 if self.local else dist.get_rank(group)
        self.sharded = sharded
        self.update_fn = update_fn
        self.cuda = flat_grad.is_cuda
        self.stream = None
        if self.cuda:
            self.stream = torch.cuda.Stream(device=flat_grad.device, priority=-1)     # high priority
        self.in_flight = False
        self.bytes_out = 0      # gradient bytes handed to the collective (per step accounting by the caller)
        self._done = {}         # tag -> events on the side stream behind that tag's buckets (partial joins, finish(tags=...))

    def _sliced(self, length, tag):
        return bucket_is_sliced(length, tag, self.world)

    def launch(self, offset: int, length: int, tag: str):
        """Call once the kernels producing flat[offset:offset+length] are enqueued on the current stream."""
        if length <= 0:
            return
        view = self.flat[offset: offset + length]
        self.bytes_out += view.numel() * view.element_size()
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            self.stream.wait_event(ev)
            ctx = torch.cuda.stream(self.stream)
        else:
            ctx = _Null()
        self.in_flight = True
        with ctx:
            if self.local:
                self.update_fn(offset, length, tag, offset, length)
                if self.cuda:
                    done = torch.cuda.Event()
                    done.record(self.stream)
                    self._done.setdefault(tag, []).append(done)
                return
            if not self.sharded:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
                return
            sliced = self._sliced(length, tag)
            if sliced:
                so, sn = rank_slice(offset, length, self.world, self.rank)
                dist.reduce_scatter_tensor(self.flat[so: so + sn], view, op=dist.ReduceOp.SUM, group=self.group)
            else:                      # replicated bucket (the biases; uneven or misaligned slices): whole-bucket
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)      # all-reduce, redundant update
                so, sn = offset, length
            self.update_fn(offset, length, tag, so, sn)
            if sliced:
                dist.all_gather_into_tensor(self.out[offset: offset + length], self.out[so: so + sn], group=self.group)

    def finish(self, tags=None):
        """Make the current (compute) stream wait for every outstanding bucket, or -- one GPU, ``tags`` given -- only for the
        buckets of those tags: the next step's fc6 needs the fc6 panels and b6, not the fc7 / fc8 update still running behind
        them.  (With collectives in flight ``tags`` is ignored: a full join, as before.)"""
        if self.cuda and self.in_flight:
            cur = torch.cuda.current_stream(self.flat.device)
            if tags is not None and self.local:
                for tag in tags:
                    for ev in self._done.pop(tag, []):
                        cur.wait_event(ev)
                return                                # still in flight: a later finish() joins the rest
            cur.wait_stream(self.stream)
        self._done.clear()
        self.in_flight = False


P2P_VERIFIED_WORLD = 2   # largest world size the peer-mapped exchange has run on (profiles/r1e_*); larger ones self-test

_OPENED = {}     # cudaIpcMemHandle bytes -> base address mapped in this process (a handle may be opened once)


def _share_with_peers(t: torch.Tensor, group):
    """Map every rank's copy of ``t`` into this process's device context (CUDA IPC; one node).
    Returns (device addresses indexed by rank, error or None).  Every rank takes part in the one
    collective whatever fails locally, so a failure never desynchronises the group; the caller
    agrees on the outcome with an all-reduce."""
    import ctypes
    from . import _lib
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    infos, ptrs, err = [None] * world, [], None
    with torch.cuda.device(t.device):
        mine = None
        try:
            hbuf, off = ctypes.create_string_buffer(64), ctypes.c_int64()
            _lib.call("nawsod_p2p_get_mem_handle", ctypes.c_void_p(t.data_ptr()), hbuf, 64, ctypes.byref(off))
            mine = (hbuf.raw, int(off.value))
        except RuntimeError as e:
            err = e
        dist.all_gather_object(infos, mine, group=group)
        for k, info in enumerate(infos):
            if k == rank:
                ptrs.append(t.data_ptr())
                continue
            try:
                if info is None:
                    raise RuntimeError("rank %d could not export its buffer" % k)
                h, off = info
                if h not in _OPENED:
                    base = ctypes.c_void_p()
                    _lib.call("nawsod_p2p_open_mem_handle", h, len(h), ctypes.byref(base))
                    _OPENED[h] = base.value
                ptrs.append(_OPENED[h] + off)
            except RuntimeError as e:
                err = err or e
                ptrs.append(0)
    return ptrs, err


class P2PExchange:
    """The sharded bucket pipeline over peer-mapped memory: copy-engine transfers + flag kernels + a fused
    reduce-and-update on the owner (CUDA only, one NVLink / NVSwitch box).

    Per bucket b (length L, slice n = L / W; rank k owns slice k):
      1. copy my part of slice k into rank k's staging area, slot [b][my rank]        (W-1 peer copies)
      2. publish seq into flag RS[b][my rank] on every rank                             (one signal kernel)
      3. wait until RS[b][*] == seq here, then update my slice from the W contributions (wait + fused SGD kernel)
      4. copy my updated GEMM-operand slice into every peer's operand buffer            (W-1 peer copies)
      5. publish seq into flag AG[b][my rank] on every rank
    finish(): wait until AG[*][*] == seq, i.e. all operands of the step have landed here; since an owner signals
    AG only after it consumed its staging area, that wait also licenses the next step's writes into it."""

    RS, AG = 0, 1

    def __init__(self, flat_grad, flat_out, plan, group=None, update_fn=None, timeout_ms=20000):
        if not flat_grad.is_cuda:
            raise RuntimeError("the p2p exchange needs CUDA buffers (use sync='sharded' with gloo on CPU)")
        self.flat, self.out, self.group = flat_grad, flat_out, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.plan = [(o, n, t) for o, n, t in plan if n > 0]
        for o, n, t in self.plan:
            if not bucket_is_sliced(n, t, self.world) and (not t.startswith("biases") or n % 4):
                raise RuntimeError("bucket of %d elements does not split into %d 32-byte aligned slices" % (n, self.world))
        self.index = {o: i for i, (o, _, _) in enumerate(self.plan)}
        self.update_fn = update_fn
        self.timeout_ms = int(os.environ.get("NAWSOD_P2P_TIMEOUT_MS", timeout_ms))     # watchdog of every wait kernel
        dev = flat_grad.device
        nb, W = len(self.plan), self.world
        # reduce-scatter leg: "pull" -- the owner's update kernel reads the ranks' gradient slices IN PLACE over NVLink (the
        # producers only publish "my bucket is complete"); "push" -- producers copy their slices into the owner's staging
        # area (engine-driven copies), the owner reduces from local memory.  Pull needs no staging memory (958 MB), no copy
        # kernels beside the GEMMs and saves the staging's HBM write + re-read (1.7 GB per step and GPU at 8 ranks).
        # Measured (profiles/r2i..r2m_bench_n*_*.json; ms per step, push/ce | pull/ce | pull/sm | push/sm): 2 GPUs 4.33 | 4.46 |
        # 4.55 | 4.81; 8 GPUs 5.50 | 4.93 | 4.99 | 5.01-5.15 -- the copy engines leave the GEMMs alone but sustain only ~360 GB/s
        # of egress when eight ranks push at once; pulling needs a seventh of the engine-driven bytes (the operand leg only).
        self.rs_mode = os.environ.get("NAWSOD_P2P_RS", "push" if self.world <= 2 else "pull")
        if self.rs_mode not in ("pull", "push"):
            raise RuntimeError("NAWSOD_P2P_RS must be 'pull' or 'push'")
        pull = self.rs_mode == "pull"
        self.stage = torch.empty(4 if pull else flat_grad.numel(), dtype=torch.float32, device=dev)
        # replicated buckets (the biases, see bucket_is_sliced): every rank receives every rank's WHOLE contribution
        # -> slot [rank] of a W x L area, one per replicated bucket, carved from one allocation
        self.rep_off, rep_total = {}, 0
        for o, n, t in self.plan:
            if not bucket_is_sliced(n, t, self.world):
                self.rep_off[o] = rep_total
                rep_total += n * W
        self.rep_stage = torch.empty(4 if pull else max(rep_total, 4), dtype=torch.float32, device=dev)
        self.flags = torch.zeros(2 * nb * W, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        self.peer_stage, e1 = _share_with_peers(self.stage, group)
        self.peer_flags, e2 = _share_with_peers(self.flags, group)
        self.peer_out, e3 = _share_with_peers(self.out, group)
        self.peer_rep, e4 = _share_with_peers(self.rep_stage, group)
        self.peer_grad, e5 = _share_with_peers(self.flat, group) if pull else (None, None)
        e3 = e3 or e4 or e5
        torch.cuda.synchronize(dev)
        ok = torch.tensor([0 if (e1 or e2 or e3) else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)          # also the setup barrier
        if int(ok.item()) == 0:
            raise RuntimeError("peer mapping failed on at least one rank: %s" % (e1 or e2 or e3 or "on a peer"))
        self.stream = torch.cuda.Stream(device=dev, priority=-1)
        self.send_stream = torch.cuda.Stream(device=dev, priority=-1)
        # b6's (tiny, replicated) update runs on a stream of its own: behind the last fc6 panel's update it would hold up
        # the next step's fc6, which reads b6 but meets the W6 panels through its in-kernel gate
        self.bias6_stream = torch.cuda.Stream(device=dev, priority=-1)
        # Consecutive buckets alternate between two update streams: a bucket's chain on its stream is wait-for-the-W-flags ->
        # reduce + update -> operand copies -> publish, and on ONE stream bucket k + 1's wait and update queued behind bucket
        # k's operand copies (8 GPUs, profiles/r2x_bench_n8_default_p2p_timeline.txt: update 0.30-0.42 ms + publish 0.27-0.37 ms
        # per fc6 panel, serialised: the last panel was published 1.5 ms after its gradient was ready).  The buckets touch
        # disjoint slices, each joins the compute stream through its own event / flags, so their order is free.
        self.ustreams = [self.stream]            # the second one is created LAST (below): the streams created so far keep the
                                                 # creation order -- and with it the hardware queues -- of the measured runs
        # "ce": copy-engine transfers (default: no SM is taken from the GEMMs); "sm": one co-resident scatter kernel per bucket
        # and leg
        self.engine = os.environ.get("NAWSOD_P2P_ENGINE", "ce")
        if self.engine not in ("sm", "ce"):
            raise RuntimeError("NAWSOD_P2P_ENGINE must be 'sm' or 'ce'")
        if len(self.plan) > 64:
            raise RuntimeError("at most 64 exchange buckets")
        nfan = int(os.environ.get("NAWSOD_P2P_COPY_STREAMS", "3"))
        self.rs_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(nfan)]
        self.ag_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(nfan)]
        self.seq = 0
        self.in_flight = False
        self.bytes_out = 0
        self.profile = None      # list of (label, bucket, event) while a caller instruments one step
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()     # lagged copy of the watchdog word
        self._status_ev = None
        self._done = [None] * len(self.plan)     # per bucket: event on the update stream behind its publish
        self._joined = set()                     # buckets of the step in flight the compute stream has already joined
        self.ustreams += [torch.cuda.Stream(device=dev, priority=-1)
                          for _ in range(max(0, int(os.environ.get("NAWSOD_P2P_USTREAMS", "2")) - 1))]

    def _flag_ptrs(self, kind, b):
        W, nb = self.world, len(self.plan)
        off = 4 * ((kind * nb + b) * W + self.rank)
        return [self.peer_flags[k] + off for k in range(W)]

    def begin_step(self):
        self.seq += 1

    def launch(self, offset: int, length: int, tag: str):
        from . import ops
        if length <= 0:
            return
        b = self.index[offset]
        W, rank = self.world, self.rank
        replicated = offset in self.rep_off       # every rank gets every contribution and updates the whole bucket
        n = length if replicated else length // W
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat.device))
        self.in_flight = True
        es_out = self.out.element_size()
        prof = self.profile
        # scatter side: never blocked by a peer, so the contributions leave as soon as their GEMM has finished
        if prof is not None:
            self.send_stream.wait_event(ev)
            with torch.cuda.stream(self.send_stream):
                prof.append(("ready", b, self._mark()))
        peers = [(rank + i) % W for i in range(1, W)]             # staggered targets
        if replicated:
            ro = self.rep_off[offset]
            rs_copies = [(self.peer_rep[k] + 4 * (ro + rank * n), self.flat.data_ptr() + 4 * offset, 4 * n) for k in peers]
        else:
            rs_copies = [(self.peer_stage[k] + 4 * (offset + rank * n), self.flat.data_ptr() + 4 * (offset + k * n), 4 * n) for k in peers]
        self.bytes_out += 4 * n * (W - 1)          # bytes of this rank's gradient that cross NVLink (pushed, or read by the owners)
        kernel_engine = self.engine == "sm"
        pull = self.rs_mode == "pull"
        if pull:
            # nothing to copy: tell every rank that this bucket of my gradient buffer is complete (it stays untouched until
            # all owners have published this step's operands, i.e. until they have read it: see finish())
            self.send_stream.wait_event(ev)
            with torch.cuda.stream(self.send_stream):
                ops.p2p_signal(self._flag_ptrs(self.RS, b), self.seq)
                if prof is not None:
                    prof.append(("sent", b, self._mark()))
        elif kernel_engine:
            self.send_stream.wait_event(ev)
            with torch.cuda.stream(self.send_stream):
                ops.p2p_scatter([c[1] for c in rs_copies], [c[0] for c in rs_copies], 4 * n, self._flag_ptrs(self.RS, b), self.seq, b)
                if prof is not None:
                    prof.append(("sent", b, self._mark()))
        else:
            self._fan_out(self.rs_streams, ev, self.send_stream, rs_copies)
            with torch.cuda.stream(self.send_stream):
                ops.p2p_signal(self._flag_ptrs(self.RS, b), self.seq)
                if prof is not None:
                    prof.append(("sent", b, self._mark()))
        # update side: wait for the W contributions, reduce + SGD on the owned slice, publish the operands
        ustream = self.bias6_stream if tag == "biases_fc6" else self.ustreams[b % len(self.ustreams)]
        ustream.wait_event(ev)
        so = offset if replicated else offset + rank * n
        with torch.cuda.stream(ustream):
            fb = (self.RS * len(self.plan) + b) * W
            ops.p2p_wait(self.flags[fb: fb + W], self.seq, self.timeout_ms, self.status)
            if prof is not None:
                prof.append(("arrived", b, self._mark()))
            if pull:             # every rank's copy of [so, so + n) of the gradient buffer, read in place (peers: over NVLink)
                grads = [self.flat[so: so + n] if r == rank else self.peer_grad[r] + 4 * so for r in range(W)]
            elif replicated:     # the same W addends in the same (rank) order on every rank: replicas stay bit-identical
                grads = [self.flat[so: so + n] if r == rank else self.rep_stage[ro + r * n: ro + (r + 1) * n] for r in range(W)]
            else:
                grads = [self.flat[so: so + n] if r == rank else self.stage[offset + r * n: offset + (r + 1) * n] for r in range(W)]
            self.update_fn(offset, length, tag, so, n, grads)
            if prof is not None:
                prof.append(("updated", b, self._mark()))
            # operand leg; a replicated bucket has nothing to send back, its AG flag only says "staging consumed"
            ag_copies = [] if replicated else [(self.peer_out[k] + es_out * so, self.out.data_ptr() + es_out * so, es_out * n) for k in peers]
            if kernel_engine:
                ops.p2p_scatter([c[1] for c in ag_copies], [c[0] for c in ag_copies], es_out * n, self._flag_ptrs(self.AG, b), self.seq, 64 + b)
            upd = torch.cuda.Event()
            upd.record()
        if not kernel_engine:
            self._fan_out(self.ag_streams, upd, ustream, ag_copies)
            with torch.cuda.stream(ustream):
                ops.p2p_signal(self._flag_ptrs(self.AG, b), self.seq)
        with torch.cuda.stream(ustream):
            if prof is not None:
                prof.append(("published", b, self._mark()))
            self._done[b] = torch.cuda.Event()
            self._done[b].record()

    def _fan_out(self, streams, after, join, copies):
        """Issue the peer copies round-robin over a few streams (their per-copy set-up latencies overlap, the
        link stays busy), each gated on event ``after``; stream ``join`` continues once all have completed."""
        from . import ops
        used = streams[:max(1, min(len(streams), len(copies)))]
        for st in used:
            st.wait_event(after)
        for i, (dst, src, nbytes) in enumerate(copies):
            with torch.cuda.stream(used[i % len(used)]):
                ops.p2p_copy(dst, src, nbytes)
        for st in used:
            e = torch.cuda.Event()
            e.record(st)
            join.wait_event(e)

    def _mark(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def finish(self, tags=None):
        """Make the current (compute) stream wait until the buckets with the given tags (default: all) are complete HERE:
        this rank's own update of them has run and every peer's operand slice of them has landed (their AG flags carry
        the step's sequence number).  An owner signals AG only after it consumed its staging area, so having joined ALL
        buckets also licenses the next step's writes into the peers' staging; step() joins the buckets in two groups --
        what fc6 reads before fc6, the rest before fc7 -- so the tail of the exchange hides behind the next step's RoI
        pooling and fc6 GEMM."""
        from . import ops
        if not self.in_flight:
            return
        nb, W = len(self.plan), self.world
        want = [i for i, (_, _, t) in enumerate(self.plan) if (tags is None or t in tags) and i not in self._joined]
        cur = torch.cuda.current_stream(self.flat.device)
        # contiguous runs of bucket indices -> one wait kernel each, on the COMPUTE stream (nothing else is held up)
        runs = []
        for i in want:
            if runs and runs[-1][1] == i:
                runs[-1][1] = i + 1
            else:
                runs.append([i, i + 1])
        for lo, hi in runs:
            for i in range(lo, hi):
                if self._done[i] is not None:
                    cur.wait_event(self._done[i])
            ops.p2p_wait(self.flags[(nb + lo) * W: (nb + hi) * W], self.seq, self.timeout_ms, self.status)
        self._joined.update(want)
        if self.profile is not None and want:
            self.profile.append(("joined %s" % ("all" if tags is None else "+".join(sorted(tags))), -1, self._mark()))
        if len(self._joined) == nb:
            for i in range(nb):
                if self._done[i] is not None:
                    cur.wait_event(self._done[i])
            cur.wait_stream(self.send_stream)
            self._status_host.copy_(self.status, non_blocking=True)
            self._status_ev = torch.cuda.Event()
            self._status_ev.record()
            self._joined = set()
            self.in_flight = False

    def fc6_gate(self):
        """The ops.FC ``gate`` for the NEXT step's stacked fc6 forward while this step's exchange is still in flight: the
        fc6 row panels are exchange buckets 0 .. P-1 in plan order, their "operands have landed" (AG) flags are consecutive
        words, and every panel has the same number of rows -- or None when that does not hold or nothing is in flight."""
        if not self.in_flight:
            return None
        panels = [(o, n) for o, n, t in self.plan if t == "fc6_panel"]
        P = len(panels)
        if P == 0 or any(t != "fc6_panel" for _, _, t in self.plan[:P]) or len({n for _, n in panels}) != 1:
            return None
        nb, W = len(self.plan), self.world
        return dict(flags=self.flags[nb * W: (nb + P) * W], nflags=W, seq=self.seq, timeout_ms=self.timeout_ms, status=self.status)

    def self_test(self, timeout_ms=3000):
        """One dry run of the whole bucket pipeline on recognisable data, so that a world size this box has not run
        before fails HERE (and the caller falls back to NCCL) instead of inside a training step.  Rank r contributes
        the constant r + 1 as every gradient; the owner checks the W contributions it received for its slice of every
        bucket and publishes r + 1 as the slice's operands; after finish() every rank checks the operand slices it
        received.  Overwrites flat_grad and the operand shadow (the caller restores the shadow from the masters).
        Returns (ok, message) -- the same on every rank (agreed with an all-reduce); never raises for a failed check."""
        W, rank, dev = self.world, self.rank, self.flat.device
        bad = torch.zeros(1, dtype=torch.int32, device=dev)        # number of failed checks on this rank
        saved_fn, saved_to, msg = self.update_fn, self.timeout_ms, ""
        sync = (lambda: torch.cuda.synchronize(dev)) if self.flat.is_cuda else (lambda: None)

        def expect(t, value):
            lo, hi = torch.aminmax(t.float())
            bad.add_(((lo != value) | (hi != value)).to(torch.int32))

        def check_and_publish(off, length, tag, so, n, grads):
            from . import ops
            for r, g in enumerate(grads):
                if isinstance(g, int):           # pull mode: a peer's gradient slice, mapped into this process
                    tmp = torch.empty(n, dtype=torch.float32, device=dev)
                    ops.p2p_copy(tmp.data_ptr(), g, 4 * n)
                    g = tmp
                expect(g, float(r + 1))
            self.out[so: so + n].fill_(float(rank + 1))

        try:
            self.update_fn, self.timeout_ms = check_and_publish, timeout_ms
            self.flat.fill_(float(rank + 1))
            sync()
            dist.barrier(group=self.group)
            self.begin_step()
            for off, length, tag in self.plan:
                self.launch(off, length, tag)
            self.finish()
            for off, length, tag in self.plan:
                if not bucket_is_sliced(length, tag, W):      # replicated: this rank updated (filled) all of it itself
                    expect(self.out[off: off + length], float(rank + 1))
                    continue
                n = length // W
                for r in range(W):
                    expect(self.out[off + r * n: off + (r + 1) * n], float(r + 1))
            sync()
            if int(self.status.item()) != 0:
                msg = "a wait kernel timed out after %d ms" % timeout_ms
            elif int(bad.item()) != 0:
                msg = "%d data checks failed on rank %d" % (int(bad.item()), rank)
        except Exception as e:               # noqa: BLE001 -- any local failure must still reach the agreement below
            msg = "%s: %s" % (type(e).__name__, e)
        finally:
            self.update_fn, self.timeout_ms = saved_fn, saved_to
        ok = torch.tensor([0 if msg else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        self.flat.zero_()
        if int(ok.item()) == 1:
            return True, "ok"
        return False, msg or "failed on a peer"

    def poll(self):
        """Non-blocking check, one finish() behind: raises once the copy of the watchdog word that the last COMPLETED
        finish() took is non-zero (the owner's update kernels have skipped their work since, see abort_flag)."""
        if self._status_ev is not None and self._status_ev.query() and int(self._status_host[0]) != 0:
            raise RuntimeError("p2p exchange: a peer did not deliver within %d ms; parameters were left un-updated" % self.timeout_ms)

    def check(self):
        """Host-side check of the watchdog word (synchronises)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("p2p exchange: a peer did not deliver within %d ms" % self.timeout_ms)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class DataParallelHead:
    """model: heads.WeblyHeadModel of this rank.  step() = fwd + bwd + gradient exchange + SGD."""

    def __init__(self, model, group=None, fc6_panels: int = 4, sync: str = "sharded", comm_sms: int | None = None):
        if sync not in ("auto", "sharded", "p2p", "allreduce"):
            raise RuntimeError("sync must be 'auto', 'sharded', 'p2p' or 'allreduce'")
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if getattr(model, "x3", False) and self.world > 1 and sync != "allreduce":
            # the sharded / peer schedules replicate only the operand (high-part) shadow of a slice to the other ranks; the
            # three-pass fp32 path re-splits the fp32 masters every step, so every rank must hold them: the reference's schedule
            raise RuntimeError("precision='fp32' across ranks needs sync='allreduce' (masters replicated on every rank)")
        self.sync = sync
        self.fc6_panels = fc6_panels
        off, n, shp = model._slices["W6"]
        assert off == 0
        nb6 = model._slices["b7"][0] - model._slices["b6"][0]          # padded length of b6, the head of the bias block
        self.plan = bucket_plan(n, shp[0], shp[1], model.n_weights, model.n_total, self.fc6_panels, n_bias_fc6=nb6)
        self.exchange = None
        self.master_sharded = False
        self.comm_sms = int(os.environ.get("NAWSOD_COMM_SMS", "0")) if comm_sms is None else comm_sms
        self._hyper = dict(momentum=0.9, weight_decay=5e-4)
        self.p2p_selftest = None                     # outcome of P2PExchange.self_test() when it ran ("ok" or the reason)
        # p2p: the next step's fc6 forward waits for its weight panels inside the kernel instead of on the stream
        self.gated_fc6 = os.environ.get("NAWSOD_P2P_GATED_FC6", "1") == "1"
        if self.world > 1 and sync in ("p2p", "auto") and model.flat_grad.is_cuda:
            # "auto": the peer-mapped path when every rank can set it up (one NVLink / NVSwitch box), else NCCL
            requested = sync
            try:                                     # raises on every rank or on none (see P2PExchange.__init__)
                self.exchange = P2PExchange(model.flat_grad, model.flat_lp, self.plan, group, update_fn=self._update_slice)
                sync = "p2p"
                self.comm_sms = 0                    # copy engines move the data: nothing to reserve
            except RuntimeError as e:
                if sync == "p2p":
                    raise
                self.exchange = None
                sync = "sharded"
                self.p2p_selftest = "not set up: %s" % e
            # "auto" on a world size the peer path was not verified on (verified on B200: 2) proves itself first
            want_test = os.environ.get("NAWSOD_P2P_SELFTEST", "auto")
            if self.exchange is not None and (want_test == "1" or (want_test == "auto" and requested == "auto" and
                                                                   self.world > P2P_VERIFIED_WORLD)):
                ok, why = self.exchange.self_test()
                self.p2p_selftest = why
                model.sync_shadow()                  # the dry run used the operand shadow as its payload
                if not ok:
                    if requested == "p2p":
                        raise RuntimeError("p2p exchange self-test failed: %s" % why)
                    self.exchange = None
                    sync = "sharded"
                    self.comm_sms = int(os.environ.get("NAWSOD_COMM_SMS", "0")) if comm_sms is None else comm_sms
        elif sync == "auto":
            sync = "sharded"
        self.sync = sync
        if self.world == 1:
            self.sync = sync = "local"
            if not model.flat_grad.is_cuda or os.environ.get("NAWSOD_LOCAL_PIPELINE", "1") == "0":
                self.fc6_panels = 1                  # plain schedule: backward, then two SGD launches
        if self.exchange is None and (self.world > 1 or self.fc6_panels > 1):
            self.exchange = GradientExchange(model.flat_grad, model.flat_lp, group, sharded=(sync != "allreduce"),
                                             update_fn=self._update_slice)
        # UpdateWorkspaceLr / RunTestNet / export_* / weights_file_blobs on the model join the update pipeline first
        model.pre_mutation_hooks.append(self.flush)

    # ------------------------------------------------------------------ parameters
    def broadcast_parameters(self):
        """detectron/utils/net_wsl.py:183-207: rank 0's parameters and momenta to every rank, once."""
        if self.world > 1:
            for t in (self.model.flat_param, self.model.flat_mom):
                dist.broadcast(t, src=0, group=self.group)
            self.model.sync_shadow()

    def gather_master_state(self):
        """Sharded mode keeps the fp32 master parameters and momenta current only on their owner rank;
        collect them everywhere (checkpoint / export contract, detectron/utils/net_wsl.py:140-181)."""
        self.flush()
        if self.world == 1 or not self.master_sharded:
            return
        for off, length, _ in self.plan:
            if length <= 0 or not bucket_is_sliced(length, _, self.world):
                continue          # updated redundantly on every rank: already complete everywhere
            so, sn = rank_slice(off, length, self.world, self.rank)
            for flat in (self.model.flat_param, self.model.flat_mom):
                dist.all_gather_into_tensor(flat[off: off + length], flat[so: so + sn], group=self.group)
        self.master_sharded = False

    # ------------------------------------------------------------------ the step
    def _update_slice(self, off, length, tag, so, sn, grads=None):
        """ACMWeightDecayMomentumSGDUpdate on [so, so+sn) (runs on the exchange stream, right behind the
        bucket's reduce-scatter; with ``grads`` the W contributions are summed inside the update kernel).
        Weights: wd, lr_mult 1; biases: no decay, lr_mult 2 (optimizer_wsl.py:106-123)."""
        from . import ops
        m = self.model
        bias = tag.startswith("biases")
        kw = dict(momentum=self._hyper["momentum"], gpu_num=self.world, lr_mult=2.0 if bias else 1.0,
                  weight_decay=0.0 if bias else self._hyper["weight_decay"], iter_count=m.iter_count,
                  p_shadow=m.flat_lp[so: so + sn])
        if grads is None:
            ops.ACMWeightDecayMomentumSGDUpdate(m.flat_grad[so: so + sn], m.flat_mom[so: so + sn], m.lr,
                                                m.flat_param[so: so + sn], None, **kw)
        else:        # a wait kernel whose watchdog fired sets the status word: the update is then skipped, step() raises
            ops.ACMWeightDecayMomentumSGDUpdateReduce(grads, m.flat_mom[so: so + sn], m.lr, m.flat_param[so: so + sn],
                                                      abort_flag=getattr(self.exchange, "status", None), **kw)

    def _limit_gemm_grid(self, on: bool):
        if self.model.flat_grad.is_cuda and self.comm_sms > 0:
            from . import _lib
            _lib.set_tuning("gemm_max_ctas", max(1, _sm_count(self.model.flat_grad.device) - self.comm_sms) if on else 0)

    def step(self, dropout_seed=None, dropout_masks=None, momentum=0.9, weight_decay=5e-4, dropout=True):
        """Dropout is on by default like the reference's training net (seed derived from the iteration count when none
        is given; ``dropout=False`` disables it); every rank draws its own keep pattern: seed * world + rank."""
        m = self.model
        if dropout and dropout_masks is None:
            base = m.iter_count + 1 if dropout_seed is None else int(dropout_seed)
            if base <= 0:
                raise RuntimeError("dropout_seed must be positive (pass dropout=False to train without Dropout)")
            dropout_seed = base * self.world + self.rank
        if self.exchange is None:
            bl = m.RunTrainStep(dropout_masks=dropout_masks, dropout_seed=dropout_seed, dropout=dropout)
            m.param_update(momentum=momentum, weight_decay=weight_decay, gpu_num=1)
            return bl
        from . import ops
        ex, cols = self.exchange, m._slices["W6"][2][1]
        self._hyper = dict(momentum=momentum, weight_decay=weight_decay)
        by_tag = {t: (o, n) for o, n, t in self.plan}
        small, biases, bias6 = by_tag["small_weights"], by_tag["biases"], by_tag.get("biases_fc6")

        p2p = self.sync == "p2p"
        rows_w6 = m._slices["W6"][2][0]
        panel_rows = ((rows_w6 + self.fc6_panels - 1) // self.fc6_panels + 255) // 256 * 256
        cols_rows_ok = rows_w6 % panel_rows == 0 and by_tag.get("biases_fc6") is not None

        def before_fc6():                            # RoI pooling needs no parameters: join the previous exchange after it
            m.fc6_gate = None
            if p2p:
                gate = ex.fc6_gate() if self.gated_fc6 else None
                if gate is not None and cols_rows_ok:
                    # fc6 meets its weight panels INSIDE the kernel (the TMA producer waits for a panel's flags when its
                    # tiles get there); the stream only joins b6, whose update runs on a stream of its own
                    gate["rows"] = panel_rows
                    m.fc6_gate = gate
                    ex.finish(tags=("biases_fc6",))
                else:
                    ex.finish(tags=("fc6_panel", "biases_fc6"))     # all that fc6 reads; the rest lands while fc6 runs
            elif self.sync == "local" and bias6 is not None:
                ex.finish(tags=("fc6_panel", "biases_fc6"))      # the fc7 / fc8 update may still run while fc6 does
            else:
                ex.finish()
            self._limit_gemm_grid(False)

        def before_fc7():
            if p2p:
                ex.finish()                          # fc7 / fc8 weights and biases of the previous step
                ex.poll()                            # a lost peer surfaces here, one step late, without a host sync
                ex.begin_step()
            elif self.sync == "local":
                ex.finish()

        rows_total = m._slices["W6"][2][0]

        def on_panel(r0, r1):
            self._limit_gemm_grid(True)              # GEMMs launched from here on share the GPU with NCCL
            ex.launch(r0 * cols, (r1 - r0) * cols, "fc6_panel")
            if r1 == rows_total and bias6 is not None:
                ex.launch(bias6[0], bias6[1], "biases_fc6")      # b6's gradient is complete with the last panel

        bl = m.RunTrainStep(dropout_masks=dropout_masks, dropout_seed=dropout_seed, dropout=dropout, fc6_panels=self.fc6_panels,
                            on_fc6_panel=on_panel, on_before_params=before_fc6, on_before_fc7=before_fc7,
                            on_small_grads=lambda: ex.launch(small[0], small[1], "small_weights"))
        ex.launch(biases[0], biases[1], "biases")
        if self.sync == "allreduce":
            ex.finish()
            self._limit_gemm_grid(False)
            m.param_update(momentum=momentum, weight_decay=weight_decay, gpu_num=self.world)
        else:
            self.master_sharded = self.world > 1
            m.iter_count += 1
        return bl

    def flush(self):
        """Join the exchange stream (end of a timed region, before reading parameters)."""
        if self.exchange is not None:
            self.exchange.finish()
            self._limit_gemm_grid(False)


def _sm_count(device):
    return torch.cuda.get_device_properties(device).multi_processor_count
