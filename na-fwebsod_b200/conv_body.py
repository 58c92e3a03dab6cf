"""Frozen VGG16 conv body feeding the head's RoIPoolF directly (second half).

The reference's ``add_VGG16_conv5_body_origin`` (detectron/modeling/VGG16.py:9-58, MODEL.CONV_BODY of the flickr
configs) is thirteen ``Conv(3x3, stride 1)`` + ``Relu`` pairs and the 2x2 ``MaxPool``s between the five groups, NCHW fp32
on cuDNN, all frozen (TRAIN.FREEZE_CONV_BODY): forward only.  ``WSL.DILATION == 2`` (the shipped flickr value, yaml:64)
keeps conv5 at 1/8 resolution: pool4 becomes kernel 2 / stride 1 and conv5_x get pad 2 / dilation 2 (VGG16.py:39-48).

Here the body keeps its activations channels-last in bf16 end to end -- the layout ``heads.WeblyHeadModel.FeedBlobs``
takes with ``x_layout='NHWC'`` -- so conv5_3 goes into RoIPoolF without a transpose or a cast.  Each convolution is an
implicit GEMM on the tcgen05 tensor cores (shifted 4-D TMA boxes of the map as the A operand, bias + ReLU in the epilogue;
csrc/conv_body.cu); conv1_1 (three input planes padded to eight) keeps the patch-matrix form (ops.Conv3x3Relu).  The layer list
below is the reference builder's own operator sequence (pinned by tests/golden/vgg16_body.npz, which records that builder
run on a tracing model).  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

# (conv blob, dim_in, dim_out) per group; a MaxPool follows groups 1-4 (VGG16.py:10-38)
_GROUPS = [
    [("conv1_1", 3, 64), ("conv1_2", 64, 64)],
    [("conv2_1", 64, 128), ("conv2_2", 128, 128)],
    [("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256)],
    [("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512)],
    [("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)],
]


def body_ops(dilation=2):
    """The operator sequence ``add_VGG16_conv5_body_origin`` emits, as (type, in, out, args) tuples in the reference's
    order and with its blob names (``Relu`` in place; ``StopGradient`` of pool2 under FREEZE_AT == 2 is a no-op for a
    forward-only body and is not listed)."""
    seq, blob = [], "data"
    for gi, group in enumerate(_GROUPS):
        last = gi == len(_GROUPS) - 1
        d = dilation if (last and dilation == 2) else 1
        for name, cin, cout in group:
            seq.append(("Conv", blob, name, dict(dim_in=cin, dim_out=cout, kernel=3, pad=d, stride=1, dilation=d)))
            seq.append(("Relu", name, name, {}))
            blob = name
        if not last:
            pool = "pool%d" % (gi + 1)
            stride = 1 if (gi == 3 and dilation == 2) else 2
            seq.append(("MaxPool", blob, pool, dict(kernel=2, pad=0, stride=stride)))
            blob = pool
    return seq


def spatial_scale(dilation=2):
    """What the builder returns beside the blob: (dim_out, spatial_scale) = (512, 1/8) with DILATION 2, else (512, 1/16)."""
    return 512, (1.0 / 8.0 if dilation == 2 else 1.0 / 16.0)


class VGG16ConvBody:
    """Weights under the reference's blob names (``conv1_1_w`` [64,3,3,3], ``conv1_1_b`` [64], ...), kept on the device as
    bf16 GEMM operands [Cout, (kh, kw, c)] (conv1_1's three input planes padded to eight) and float32 biases."""

    def __init__(self, dilation=2, device="cuda"):
        if dilation not in (1, 2):
            raise RuntimeError("WSL.DILATION must be 1 or 2 (VGG16.py:39-58)")
        self.dilation, self.device = dilation, torch.device(device)
        self.w, self.b = {}, {}
        self.blobs = {}
        self.implicit = None          # None: implicit GEMM wherever Cin % 64 == 0 (ops.Conv3x3Relu); False: patch matrix everywhere

    @staticmethod
    def _cin_padded(cin):
        return (cin + 7) // 8 * 8

    def load_reference_params(self, params):
        """params: ``<conv>_w`` [Cout, Cin, 3, 3] and ``<conv>_b`` [Cout] float32 arrays (NumPy or torch) for all thirteen
        convolutions -- the blobs ``initialize_gpu_from_weights_file`` loads from the ImageNet VGG16 pickle
        (detectron/utils/net_wsl.py:53-137)."""
        for group in _GROUPS:
            for name, cin, cout in group:
                w, b = params[name + "_w"], params[name + "_b"]
                w = torch.from_numpy(np.ascontiguousarray(w)) if isinstance(w, np.ndarray) else w
                b = torch.from_numpy(np.ascontiguousarray(b)) if isinstance(b, np.ndarray) else b
                if tuple(w.shape) != (cout, cin, 3, 3) or tuple(b.shape) != (cout,):
                    raise RuntimeError("blob %s_w / _b has shape %s / %s, expected %s / %s" % (
                        name, tuple(w.shape), tuple(b.shape), (cout, cin, 3, 3), (cout,)))
                cp = self._cin_padded(cin)
                wm = torch.zeros((cout, 3, 3, cp), dtype=torch.float32)
                wm[:, :, :, :cin] = w.float().permute(0, 2, 3, 1)
                self.w[name] = wm.reshape(cout, 9 * cp).to(self.device, torch.bfloat16).contiguous()
                self.b[name] = b.float().to(self.device).contiguous()

    def feed_image(self, data):
        """``data`` blob [N,3,H,W] float32 (NCHW, mean-subtracted BGR like the reference's, roi_data/minibatch_wsl.py) ->
        channels-last bf16 with the three planes padded to eight."""
        if not data.is_cuda or data.dim() != 4 or data.shape[1] != 3:
            raise RuntimeError("data must be a CUDA [N,3,H,W] tensor")
        N, _, H, W = data.shape
        x = torch.zeros((N, H, W, 8), dtype=torch.bfloat16, device=data.device)
        x[..., :3] = data.permute(0, 2, 3, 1).to(torch.bfloat16)
        self.blobs["data"] = x
        return x

    def run(self, keep=()):
        """Forward through conv5_3.  Returns (conv5_3 [N,h,w,512] bf16 channels-last, 512, spatial_scale) -- the builder's
        return values; blobs named in ``keep`` stay in ``self.blobs`` (debugging / per-layer parity)."""
        if not self.w:
            raise RuntimeError("VGG16ConvBody: load_reference_params first")
        cur = {"data": self.blobs["data"]}
        for kind, src, dst, args in body_ops(self.dilation):
            if kind == "Conv":
                imp = None if self.implicit is None else (self.implicit and args["dim_in"] % 64 == 0)
                cur = {dst: ops.Conv3x3Relu(cur[src], self.w[dst], self.b[dst], dilation=args["dilation"], relu=True, implicit=imp)}
            elif kind == "MaxPool":
                cur = {dst: ops.MaxPool2x2(cur[src], stride=args["stride"])}
            # Relu: fused into the convolution's GEMM epilogue
            if dst in keep:
                self.blobs[dst] = cur[dst]
        out = cur["conv5_3"]
        self.blobs["conv5_3"] = out
        dim, scale = spatial_scale(self.dilation)
        return out, dim, scale


def add_VGG16_conv5_body_origin(body: VGG16ConvBody):
    """Same name and return values as detectron/modeling/VGG16.py:9-58: (blob_out, dim_out, spatial_scale)."""
    return body.run()
