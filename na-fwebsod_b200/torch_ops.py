"""The reference's operators as PyTorch custom ops (``torch.ops.nawsod.*``) with autograd -- the harness-facing form of the
drop-in boundary (BASELINE.json north_star: "a thin C-ABI library called from Python, PyTorch custom ops as the harness";
SURVEY.md 8b: "identical argument order to the Caffe2 blob lists so a NetDef -> call translation is 1:1").

Each op takes the Caffe2 operator's input blobs in their order, then its named arguments, and returns its output blobs
in their order (registry strings and argument names: detectron/ops/*.cc ``OPERATOR_SCHEMA``; ``RoIPoolF``:
detectron/modeling/detector.py:321-329).  The gradient of an op is the reference's gradient OPERATOR -- the kernel
``GetGradientDefs`` of the op's ``REGISTER_GRADIENT`` names -- not a derivative re-derived by autograd:

    nawsod::RoIPoolF(X, rois, pooled_h, pooled_w, spatial_scale) -> (Y, argmax)      grad: RoIPoolFGradient([X, rois, argmax, dY])
    nawsod::RoIFeatureBoost(X, S) -> Y                                               grad: RoIFeatureBoostGradient([dY, S]) (none for S)
    nawsod::FC(X, W, b) -> Y                                                         grad: FCGradient([X, W, dY]) -> dW, db, dX
    nawsod::RoIIoU(rois) -> J                                                        (no gradient: roi_iou_op.cc has none)
    nawsod::CrossEntropyWithLogits(X, L, is_mean) -> Y                               grad: ...Gradient([X, L, dY])
    nawsod::WeightedCrossEntropyWithLogits(X, L, W, is_mean) -> Y                    grad: ...Gradient([X, L, W, dY])
    nawsod::MinEntropyLoss(X, L) -> Y                                                grad: MinEntropyLossGradient([X, L, dY])
    nawsod::ACMWeightDecayMomentumSGDUpdate(g, m, lr, p, acc?, momentum, iter_size, gpu_num, lr_mult, weight_decay,
                                            iter_count) -> ()                        in place on m, p, acc (the op's AllowInplace)

Every implementation is one call into ``ops`` (ctypes -> libnawsod.so): CUDA tensors only, a CPU tensor raises
``RuntimeError`` -- there is no fallback.  The shape functions (``register_fake``) let the ops be traced on meta / fake
tensors.  The fused hot path (heads.WeblyHeadModel) does not go through autograd; these ops are for callers that hold a
graph of reference operators and want each one replaced 1:1.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.library import custom_op

from . import ops

__all__ = ["RoIPoolF", "RoIFeatureBoost", "FC", "RoIIoU", "CrossEntropyWithLogits", "WeightedCrossEntropyWithLogits",
           "MinEntropyLoss", "ACMWeightDecayMomentumSGDUpdate"]


# ------------------------------------------------------------------------------------------------- RoIPoolF
@custom_op("nawsod::RoIPoolF", mutates_args=())
def RoIPoolF(X: torch.Tensor, rois: torch.Tensor, pooled_h: int, pooled_w: int, spatial_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    Y, argmax = ops.RoIPoolF(X, rois, pooled_h=pooled_h, pooled_w=pooled_w, spatial_scale=spatial_scale)
    return Y, argmax


@RoIPoolF.register_fake
def _(X, rois, pooled_h, pooled_w, spatial_scale):
    shape = (rois.shape[0], X.shape[1], pooled_h, pooled_w)
    return X.new_empty(shape), X.new_empty(shape, dtype=torch.int32)


def _roi_pool_setup(ctx, inputs, output):
    X, rois = inputs[0], inputs[1]
    ctx.save_for_backward(X, rois, output[1])


def _roi_pool_backward(ctx, dY, _dargmax):
    X, rois, argmax = ctx.saved_tensors
    dX = torch.ops.nawsod.RoIPoolFGradient(X, rois, argmax, dY.contiguous())
    return dX, None, None, None, None


@custom_op("nawsod::RoIPoolFGradient", mutates_args=())
def RoIPoolFGradient(X: torch.Tensor, rois: torch.Tensor, argmax: torch.Tensor, dY: torch.Tensor) -> torch.Tensor:
    return ops.RoIPoolFGradient(X, rois, argmax, dY, layout="NCHW")


@RoIPoolFGradient.register_fake
def _(X, rois, argmax, dY):
    return X.new_empty(X.shape, dtype=torch.float32)


RoIPoolF.register_autograd(_roi_pool_backward, setup_context=_roi_pool_setup)


# ------------------------------------------------------------------------------------------------- RoIFeatureBoost
@custom_op("nawsod::RoIFeatureBoost", mutates_args=())
def RoIFeatureBoost(X: torch.Tensor, S: torch.Tensor) -> torch.Tensor:
    return ops.RoIFeatureBoost(X, S)


@RoIFeatureBoost.register_fake
def _(X, S):
    return torch.empty_like(X)


def _boost_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])


def _boost_backward(ctx, dY):
    (S,) = ctx.saved_tensors
    return torch.ops.nawsod.RoIFeatureBoost(dY.contiguous(), S), None      # RoIFeatureBoostGradient is the same scaling


RoIFeatureBoost.register_autograd(_boost_backward, setup_context=_boost_setup)


# ------------------------------------------------------------------------------------------------- FC / FCGradient
@custom_op("nawsod::FC", mutates_args=())
def FC(X: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return ops.FC(X, W, b)


@FC.register_fake
def _(X, W, b):
    return X.new_empty((X.shape[0], W.shape[0]))


@custom_op("nawsod::FCGradient", mutates_args=())
def FCGradient(X: torch.Tensor, W: torch.Tensor, dY: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    dW, db = ops.FCGradientW(dY, X)
    dX = ops.FCGradientX(dY, W)
    return dW, db, dX


@FCGradient.register_fake
def _(X, W, dY):
    return (W.new_empty(W.shape, dtype=torch.float32), W.new_empty((W.shape[0],), dtype=torch.float32), torch.empty_like(X))


def _fc_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _fc_backward(ctx, dY):
    X, W = ctx.saved_tensors
    dW, db, dX = torch.ops.nawsod.FCGradient(X, W, dY.contiguous())
    return dX, dW.to(W.dtype), db


FC.register_autograd(_fc_backward, setup_context=_fc_setup)


# ------------------------------------------------------------------------------------------------- RoIIoU
@custom_op("nawsod::RoIIoU", mutates_args=())
def RoIIoU(rois: torch.Tensor) -> torch.Tensor:
    return ops.RoIIoU(rois)


@RoIIoU.register_fake
def _(rois):
    return rois.new_empty((rois.shape[0], rois.shape[0]))


# ------------------------------------------------------------------------------------------------- cross-entropy losses
@custom_op("nawsod::CrossEntropyWithLogits", mutates_args=())
def CrossEntropyWithLogits(X: torch.Tensor, L: torch.Tensor, is_mean: bool) -> torch.Tensor:
    return ops.CrossEntropyWithLogits(X, L, is_mean=is_mean)


@CrossEntropyWithLogits.register_fake
def _(X, L, is_mean):
    return X.new_empty(())


@custom_op("nawsod::CrossEntropyWithLogitsGradient", mutates_args=())
def CrossEntropyWithLogitsGradient(X: torch.Tensor, L: torch.Tensor, dY: torch.Tensor, is_mean: bool) -> torch.Tensor:
    return ops.CrossEntropyWithLogitsGradient(X, L, dY, is_mean=is_mean)


@CrossEntropyWithLogitsGradient.register_fake
def _(X, L, dY, is_mean):
    return torch.empty_like(X)


def _ce_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.is_mean = inputs[2]


def _ce_backward(ctx, dY):
    X, L = ctx.saved_tensors
    return torch.ops.nawsod.CrossEntropyWithLogitsGradient(X, L, dY.contiguous(), ctx.is_mean), None, None


CrossEntropyWithLogits.register_autograd(_ce_backward, setup_context=_ce_setup)


@custom_op("nawsod::WeightedCrossEntropyWithLogits", mutates_args=())
def WeightedCrossEntropyWithLogits(X: torch.Tensor, L: torch.Tensor, W: torch.Tensor, is_mean: bool) -> torch.Tensor:
    return ops.WeightedCrossEntropyWithLogits(X, L, W, is_mean=is_mean)


@WeightedCrossEntropyWithLogits.register_fake
def _(X, L, W, is_mean):
    return X.new_empty(())


@custom_op("nawsod::WeightedCrossEntropyWithLogitsGradient", mutates_args=())
def WeightedCrossEntropyWithLogitsGradient(X: torch.Tensor, L: torch.Tensor, W: torch.Tensor, dY: torch.Tensor, is_mean: bool) -> torch.Tensor:
    return ops.WeightedCrossEntropyWithLogitsGradient(X, L, W, dY, is_mean=is_mean)


@WeightedCrossEntropyWithLogitsGradient.register_fake
def _(X, L, W, dY, is_mean):
    return torch.empty_like(X)


def _wce_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], inputs[2])
    ctx.is_mean = inputs[3]


def _wce_backward(ctx, dY):
    X, L, W = ctx.saved_tensors
    # the class weights are StopGradient-ed in the reference (modeling/webly_heads.py:390-391): no gradient for W
    return torch.ops.nawsod.WeightedCrossEntropyWithLogitsGradient(X, L, W, dY.contiguous(), ctx.is_mean), None, None, None


WeightedCrossEntropyWithLogits.register_autograd(_wce_backward, setup_context=_wce_setup)


# ------------------------------------------------------------------------------------------------- MinEntropyLoss
@custom_op("nawsod::MinEntropyLoss", mutates_args=())
def MinEntropyLoss(X: torch.Tensor, L: torch.Tensor) -> torch.Tensor:
    return ops.MinEntropyLoss(X, L)


@MinEntropyLoss.register_fake
def _(X, L):
    return X.new_empty(())


@custom_op("nawsod::MinEntropyLossGradient", mutates_args=())
def MinEntropyLossGradient(X: torch.Tensor, L: torch.Tensor, dY: torch.Tensor) -> torch.Tensor:
    return ops.MinEntropyLossGradient(X, L, dY)


@MinEntropyLossGradient.register_fake
def _(X, L, dY):
    return torch.empty_like(X)


def _me_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _me_backward(ctx, dY):
    X, L = ctx.saved_tensors
    return torch.ops.nawsod.MinEntropyLossGradient(X, L, dY.contiguous()), None


MinEntropyLoss.register_autograd(_me_backward, setup_context=_me_setup)


# ------------------------------------------------------------------------------------------------- SGD update
@custom_op("nawsod::ACMWeightDecayMomentumSGDUpdate", mutates_args=("m", "p", "acc"))
def ACMWeightDecayMomentumSGDUpdate(g: torch.Tensor, m: torch.Tensor, lr: torch.Tensor, p: torch.Tensor, acc: Optional[torch.Tensor],
                                    momentum: float, iter_size: int, gpu_num: int, lr_mult: float, weight_decay: float,
                                    iter_count: int) -> None:
    ops.ACMWeightDecayMomentumSGDUpdate(g, m, lr, p, acc, momentum=momentum, iter_size=iter_size, gpu_num=gpu_num, lr_mult=lr_mult,
                                        weight_decay=weight_decay, iter_count=iter_count)


@ACMWeightDecayMomentumSGDUpdate.register_fake
def _(g, m, lr, p, acc, momentum, iter_size, gpu_num, lr_mult, weight_decay, iter_count):
    return None
