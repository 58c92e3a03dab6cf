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


def bucket_plan(n_weights_w6: int, w6_rows: int, w6_cols: int, n_weights: int, n_total: int, panels: int, align_rows: int = 256):
    """Ordered exchange buckets (offset, length, tag) over the flat gradient buffer laid out as
    [W6 | other weights | biases], in the order the backward pass completes them: fc6 row panels,
    then the fc7/fc8 weights, then the biases.  Every element is covered exactly once."""
    assert n_weights_w6 == w6_rows * w6_cols
    plan = []
    step = ((w6_rows + panels - 1) // panels + align_rows - 1) // align_rows * align_rows
    for r0 in range(0, w6_rows, step):
        r1 = min(w6_rows, r0 + step)
        plan.append((r0 * w6_cols, (r1 - r0) * w6_cols, "fc6_panel"))
    w6p = n_weights_w6_padded(n_weights_w6)
    plan.append((w6p, n_weights - w6p, "small_weights"))
    plan.append((n_weights, n_total - n_weights, "biases"))
    return plan


def rank_slice(offset: int, length: int, world: int, rank: int):
    """Rank's contiguous share of a bucket (the reduce-scatter / all-gather chunk)."""
    if length % world:
        raise RuntimeError("bucket of %d elements is not divisible by world size %d" % (length, world))
    n = length // world
    return offset + rank * n, n


class GradientExchange:
    """Bucket pipeline on a side stream (CUDA) or inline (CPU tensors / gloo).

    sharded:   reduce_scatter(grad bucket) -> update_fn(bucket, own slice) -> all_gather(out bucket)
    allreduce: all_reduce(grad bucket)                       (update happens once, after finish())
    """

    def __init__(self, flat_grad: torch.Tensor, flat_out, group=None, sharded: bool = True, update_fn=None):
        self.flat, self.out = flat_grad, flat_out
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.sharded = sharded
        self.update_fn = update_fn
        self.cuda = flat_grad.is_cuda
        self.stream = None
        if self.cuda:
            self.stream = torch.cuda.Stream(device=flat_grad.device, priority=-1)     # high priority
        self.in_flight = False
        self.bytes_out = 0      # gradient bytes handed to the collective (per step accounting by the caller)

    def _divisible(self, length):
        return length % self.world == 0

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
            if not self.sharded:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
                return
            if self._divisible(length):
                so, sn = rank_slice(offset, length, self.world, self.rank)
                dist.reduce_scatter_tensor(self.flat[so: so + sn], view, op=dist.ReduceOp.SUM, group=self.group)
            else:                      # odd world sizes: whole-bucket all-reduce, every rank updates its share redundantly
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
                so, sn = offset, length
            self.update_fn(offset, length, tag, so, sn)
            if self._divisible(length):
                dist.all_gather_into_tensor(self.out[offset: offset + length], self.out[so: so + sn], group=self.group)

    def finish(self):
        """Make the current (compute) stream wait for every outstanding bucket."""
        if self.cuda and self.in_flight:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        self.in_flight = False


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class DataParallelHead:
    """model: heads.WeblyHeadModel of this rank.  step() = fwd + bwd + gradient exchange + SGD."""

    def __init__(self, model, group=None, fc6_panels: int = 4, sync: str = "sharded", comm_sms: int | None = None):
        if sync not in ("sharded", "allreduce"):
            raise RuntimeError("sync must be 'sharded' or 'allreduce'")
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.sync = sync
        self.fc6_panels = fc6_panels if self.world > 1 else 1
        off, n, shp = model._slices["W6"]
        assert off == 0
        self.plan = bucket_plan(n, shp[0], shp[1], model.n_weights, model.n_total, self.fc6_panels)
        self.exchange = None
        self.master_sharded = False
        self.comm_sms = int(os.environ.get("NAWSOD_COMM_SMS", "0")) if comm_sms is None else comm_sms
        self._hyper = dict(momentum=0.9, weight_decay=5e-4)
        if self.world > 1:
            self.exchange = GradientExchange(model.flat_grad, model.flat_lp, group, sharded=(sync == "sharded"),
                                             update_fn=self._update_slice)

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
            if length <= 0 or length % self.world:
                continue
            so, sn = rank_slice(off, length, self.world, self.rank)
            for flat in (self.model.flat_param, self.model.flat_mom):
                dist.all_gather_into_tensor(flat[off: off + length], flat[so: so + sn], group=self.group)
        self.master_sharded = False

    # ------------------------------------------------------------------ the step
    def _update_slice(self, off, length, tag, so, sn):
        """ACMWeightDecayMomentumSGDUpdate on [so, so+sn) (runs on the exchange stream, right behind the
        bucket's reduce-scatter).  Weights: wd, lr_mult 1; biases: no decay, lr_mult 2 (optimizer_wsl.py:106-123)."""
        from . import ops
        m = self.model
        bias = tag == "biases"
        ops.ACMWeightDecayMomentumSGDUpdate(
            m.flat_grad[so: so + sn], m.flat_mom[so: so + sn], m.lr, m.flat_param[so: so + sn], None,
            momentum=self._hyper["momentum"], gpu_num=self.world, lr_mult=2.0 if bias else 1.0,
            weight_decay=0.0 if bias else self._hyper["weight_decay"], iter_count=m.iter_count,
            p_shadow=m.flat_lp[so: so + sn])

    def _limit_gemm_grid(self, on: bool):
        if self.model.flat_grad.is_cuda and self.comm_sms > 0:
            from . import _lib
            _lib.set_tuning("gemm_max_ctas", max(1, _sm_count(self.model.flat_grad.device) - self.comm_sms) if on else 0)

    def step(self, dropout_seed=0, dropout_masks=None, momentum=0.9, weight_decay=5e-4):
        m = self.model
        if self.world == 1:
            bl = m.RunTrainStep(dropout_masks=dropout_masks, dropout_seed=dropout_seed)
            m.param_update(momentum=momentum, weight_decay=weight_decay, gpu_num=1)
            return bl
        ex, cols = self.exchange, m._slices["W6"][2][1]
        self._hyper = dict(momentum=momentum, weight_decay=weight_decay)
        ex.finish()                                  # the previous step's updated operands must have landed
        self._limit_gemm_grid(False)
        small, biases = self.plan[-2], self.plan[-1]

        def on_panel(r0, r1):
            self._limit_gemm_grid(True)              # GEMMs launched from here on share the GPU with NCCL
            ex.launch(r0 * cols, (r1 - r0) * cols, "fc6_panel")

        bl = m.RunTrainStep(dropout_masks=dropout_masks, dropout_seed=dropout_seed, fc6_panels=self.fc6_panels,
                            on_fc6_panel=on_panel,
                            on_small_grads=lambda: ex.launch(small[0], small[1], "small_weights"))
        ex.launch(biases[0], biases[1], "biases")
        if self.sync == "allreduce":
            ex.finish()
            self._limit_gemm_grid(False)
            m.param_update(momentum=momentum, weight_decay=weight_decay, gpu_num=self.world)
        else:
            self.master_sharded = True
            m.iter_count += 1
        return bl

    def flush(self):
        """Join the exchange stream (end of a timed region, before reading parameters)."""
        if self.exchange is not None:
            self.exchange.finish()
            self._limit_gemm_grid(False)


def _sm_count(device):
    return torch.cuda.get_device_properties(device).multi_processor_count
