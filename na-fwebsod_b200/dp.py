"""Data-parallel training step of the head: one process per GPU, images sharded by rank, one
gradient all-reduce (sum) per step and an identical local SGD update on every rank.

Mirrors detectron/modeling/optimizer_wsl.py:18-137: ``_add_allreduce_graph`` (:52-72) issues one
``NCCLAllreduce`` (sum, in place) per parameter blob after the whole backward graph and the
``1/(iter_size*gpu_num)`` averaging lives in the SGD op (acm_weightdecay_momentum_sgd_op.h:79-84).
Here the parameters' gradients are one flat float32 buffer, so the exchange is a handful of large
bucketed ``all_reduce`` calls on a side stream, ordered by when each bucket becomes ready:

    1. fc7 / fc8 weight gradients (ready first, ~14 % of the bytes),
    2. the fc6 weight gradient in row panels, each launched as soon as its GEMM has been
       enqueued -> the transfer of panel p overlaps the tensor-core GEMM of panel p+1,
    3. the bias gradients (tiny tail).

torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is the
plumbing; the path has no other collective.  Inference shards by image with no
collective at all (replicas only).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_images(num_images_total: int, world_size: int, rank: int):
    """Rank g gets images {g*B .. g*B+B-1} with all their RoIs; B must divide evenly."""
    if num_images_total % world_size:
        raise RuntimeError("global image batch %d is not divisible by world size %d" % (num_images_total, world_size))
    b = num_images_total // world_size
    return list(range(rank * b, (rank + 1) * b))


def bucket_plan(n_weights_w6: int, w6_rows: int, w6_cols: int, n_weights: int, n_total: int, panels: int, align_rows: int = 256):
    """Ordered all-reduce buckets (offset, length, tag) over the flat gradient buffer laid out as
    [W6 | other weights | biases].  Every element is covered exactly once."""
    assert n_weights_w6 == w6_rows * w6_cols
    plan = [(n_weights_w6_padded(n_weights_w6), n_weights - n_weights_w6_padded(n_weights_w6), "small_weights")]
    step = ((w6_rows + panels - 1) // panels + align_rows - 1) // align_rows * align_rows
    for r0 in range(0, w6_rows, step):
        r1 = min(w6_rows, r0 + step)
        plan.append((r0 * w6_cols, (r1 - r0) * w6_cols, "fc6_panel"))
    plan.append((n_weights, n_total - n_weights, "biases"))
    return plan


def n_weights_w6_padded(n: int, align: int = 64) -> int:
    return (n + align - 1) // align * align


class GradientAllReducer:
    """Issues the bucketed all-reduces on a side stream (CUDA) or inline (CPU / gloo)."""

    def __init__(self, flat_grad: torch.Tensor, group=None):
        self.flat = flat_grad
        self.group = group
        self.cuda = flat_grad.is_cuda
        self.stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self.bytes = 0

    def reduce_bucket(self, offset: int, length: int):
        if length <= 0:
            return
        view = self.flat[offset: offset + length]
        self.bytes += view.numel() * view.element_size()
        if not self.cuda:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat.device))   # the producer kernels enqueued so far
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """Make the compute stream wait for every outstanding all-reduce."""
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)


class DataParallelHead:
    """model: heads.WeblyHeadModel of this rank.  step() = fwd + bwd + all-reduce + SGD."""

    def __init__(self, model, group=None, fc6_panels: int = 4):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.fc6_panels = fc6_panels if self.world > 1 else 1
        self.reducer = GradientAllReducer(model.flat_grad, group) if self.world > 1 else None
        off, n, shp = model._slices["W6"]
        assert off == 0
        self.plan = bucket_plan(n, shp[0], shp[1], model.n_weights, model.n_total, self.fc6_panels)

    def broadcast_parameters(self):
        """detectron/utils/net_wsl.py:183-207: rank 0's parameters and momenta to every rank, once."""
        if self.world > 1:
            for t in (self.model.flat_param, self.model.flat_mom):
                dist.broadcast(t, src=0, group=self.group)
            self.model.sync_shadow()

    def step(self, dropout_seed=0, dropout_masks=None, momentum=0.9, weight_decay=5e-4):
        m = self.model
        if self.world == 1:
            bl = m.RunTrainStep(dropout_masks=dropout_masks, dropout_seed=dropout_seed)
        else:
            red, cols = self.reducer, m._slices["W6"][2][1]
            small = self.plan[0]
            bl = m.RunTrainStep(dropout_masks=dropout_masks, dropout_seed=dropout_seed, fc6_panels=self.fc6_panels,
                                on_small_grads=lambda: red.reduce_bucket(small[0], small[1]),
                                on_fc6_panel=lambda r0, r1: red.reduce_bucket(r0 * cols, (r1 - r0) * cols))
            red.reduce_bucket(self.plan[-1][0], self.plan[-1][1])
            red.finish()
        m.param_update(momentum=momentum, weight_decay=weight_decay, gpu_num=self.world)
        return bl
