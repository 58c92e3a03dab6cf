"""Host-side mirror of the reference's training-input contract for the head.

The reference's loader threads build the ``rois`` / ``obn_scores`` / ``labels_oh`` blobs with NumPy
(detectron/roi_data/wsl.py:59-225, detectron/roi_data/minibatch_wsl.py:53-171) and, for webly data, mix two
images of one class into a single training example (detectron/roi_data/loader_wsl.py:130-170) before the
blobs are copied to the GPU.  The same-named functions here take roidb entries whose arrays already live in
HBM (``boxes`` [n,4] float32, ``obn_scores`` [n,1] float32, ``gt_classes`` [n] int32) and fill the minibatch
blobs with stream-ordered libnawsod calls; the only host values are the per-image scalars the reference also
computes on the host (scale, crop window, mixup coefficient).  The blobs feed ``WeblyHeadModel.FeedBlobs``
directly; ``roi_offsets`` (the per-image row ranges, contiguous by construction of ``add_wsl_blobs``) is the
one blob the reference does not have: its head asserts one image per GPU (modeling/wsl_heads.py:214).

PyTorch is plumbing here (allocation, slicing); no arithmetic happens in this file.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def prep_im_scale(im_h, im_w, target_size, max_size):
    """The scale ``prep_im_for_blob`` returns (detectron/utils/blob.py:117-122)."""
    im_size_min, im_size_max = min(im_h, im_w), max(im_h, im_w)
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    return im_scale


def random_crop(im_h, im_w, crop, r0, r1):
    """WSL.USE_CROP window (detectron/roi_data/minibatch_wsl.py:142-153) from the two uniform draws:
    (row1, col1, row2, col2) int32, truncated like ``np.array(..., dtype=np.int32)``."""
    shape = np.array([im_h, im_w])
    crop_dims = shape * crop
    s = shape - crop_dims
    s[0] *= r0
    s[1] *= r1
    return np.array([s[0], s[1], s[0] + crop_dims[0] - 1, s[1] + crop_dims[1] - 1], dtype=np.int32)


def crops_to_xyxy(im_crops):
    """(row, col, row, col) -> (x1, y1, x2, y2) int32 (detectron/roi_data/minibatch_wsl.py:63-64)."""
    return np.array(im_crops, dtype=np.int32).reshape(-1, 4)[:, (1, 0, 3, 2)]


def _as_device_i32(a, device):
    if isinstance(a, torch.Tensor):
        return a
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)


def _sample_rois(roidb, im_scale, im_crop, batch_idx, *, rois_per_image, num_classes, out=None):
    """``_sample_rois`` (detectron/roi_data/wsl.py:87-181): the first ``rois_per_image`` boxes of the entry,
    projected into the cropped + rescaled training image, ``obn_scores + 1``, and the image-level labels.
    ``gt_classes`` given as a NumPy array is checked on the host like the reference's assert (:148-152); a
    device tensor is not read back (labels_int32 is -1 if the entry has no ground-truth row).
    ``out``: dict of row slices of the minibatch blobs to fill in place (used by add_wsl_blobs)."""
    boxes, obn = roidb["boxes"], roidb["obn_scores"]
    n = int(min(int(rois_per_image), boxes.shape[0]))
    gt = roidb["gt_classes"]
    if not isinstance(gt, torch.Tensor) and not np.any(np.asarray(gt) > 0):
        raise RuntimeError("Empty ground truth empty for image is not allowed. Please check.")
    out = out or {}
    rois, scores = ops.sample_rois(boxes[:n], im_scale, im_crop, batch_idx, obn_scores=obn[:n],
                                   out_rois=out.get("rois"), out_obn=out.get("obn_scores"))
    oh, li = ops.image_labels(_as_device_i32(gt, boxes.device), num_classes, out_oh=out.get("labels_oh"), out_int=out.get("labels_int32"))
    return dict(labels_int32=li, labels_oh=oh, rois=rois, obn_scores=scores)


def add_wsl_blobs(blobs, im_scales, im_crops, roidb, *, rois_per_image, num_classes):
    """``add_wsl_blobs`` (detectron/roi_data/wsl.py:59-85): fills ``blobs`` with the concatenation over the
    images of the minibatch -- rois [sum R,5] (image i's rows carry batch index i), obn_scores [sum R,1],
    labels_int32 [B], labels_oh [B, num_classes-1] -- written in place by each image's kernels (no concat
    copies), plus roi_offsets [B+1] int32.  ``im_crops``: (x1,y1,x2,y2) per image.  Returns True (``valid``)."""
    B = len(roidb)
    if B == 0 or len(im_scales) != B or len(im_crops) != B:
        raise RuntimeError("add_wsl_blobs: need one scale and one crop per roidb entry")
    dev = roidb[0]["boxes"].device
    counts = [int(min(int(rois_per_image), e["boxes"].shape[0])) for e in roidb]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    total = int(offs[-1])
    blobs["rois"] = torch.empty((total, 5), dtype=torch.float32, device=dev)
    blobs["obn_scores"] = torch.empty((total, 1), dtype=torch.float32, device=dev)
    blobs["labels_int32"] = torch.empty(B, dtype=torch.int32, device=dev)
    blobs["labels_oh"] = torch.empty((B, num_classes - 1), dtype=torch.float32, device=dev)
    for i, e in enumerate(roidb):
        a, b = int(offs[i]), int(offs[i + 1])
        _sample_rois(e, im_scales[i], im_crops[i], i, rois_per_image=rois_per_image, num_classes=num_classes,
                     out=dict(rois=blobs["rois"][a:b], obn_scores=blobs["obn_scores"][a:b],
                              labels_oh=blobs["labels_oh"][i:i + 1], labels_int32=blobs["labels_int32"][i:i + 1]))
    blobs["roi_offsets"] = torch.from_numpy(offs).to(dev)
    return True


def bagging_mixup(blobs, lam):
    """The bagging-mixup block of ``RoIDataLoader.get_next_minibatch`` (detectron/roi_data/loader_wsl.py:149-168)
    on a two-image minibatch: ``data`` (if present) and ``labels_oh`` become the (lam, 1-lam) combination, every
    RoI of both images belongs to image 0, the per-image blobs keep image 0's entry.  In place on ``blobs``.
    (``data`` is the network-input image blob, upstream of the conv body; it is mixed only if the caller put it
    there.  A caller that feeds conv5 maps runs the body on the mixed image, not the mix on two conv5 maps.)"""
    if blobs["labels_oh"].shape[0] != 2:
        raise RuntimeError("bagging_mixup needs a two-image minibatch")
    for k in ("data", "labels_oh"):
        if k in blobs:
            blobs[k] = ops.bagging_mixup(blobs[k], lam)
    ops.set_column(blobs["rois"], 0, 0.0)
    for k in ("data_ids", "labels_int32"):
        if k in blobs:
            blobs[k] = blobs[k][0:1]
    total = blobs["rois"].shape[0]
    blobs["roi_offsets"] = torch.tensor([0, total], dtype=torch.int32).to(blobs["rois"].device)
    return blobs


def convert_mcg_boxes(bboxes):
    """tools/convert_mcg.py:45-49 on the GPU: float64 .mat boxes (1-indexed y1,x1,y2,x2) -> (x1,y1,x2,y2) stored as
    uint16 bit patterns in an int16 tensor (see ops.convert_mcg_boxes)."""
    return ops.convert_mcg_boxes(bboxes)
