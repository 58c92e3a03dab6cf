"""CPU oracle for the test-time wrapper around the head.
TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and nothing under ``na-fwebsod_b200/``).

Restates, function by function (citations relative to /root/reference/detectron):
  * ``_project_im_rois`` / ``_get_rois_blob`` / ``flip_boxes``  core/test_wsl.py:998-1027, utils/boxes.py:246-251
  * the dedup block of ``im_detect_bbox``                      core/test_wsl.py:125-133, 173-176
  * the 'AVG' score heuristic of ``im_detect_bbox_aug``        core/test_wsl.py:262-263
  * ``box_results_with_nms_and_limit``                         core/test_wsl.py:803-863
  * greedy NMS                                                 utils/cython_nms.pyx:38-93 (C restatement oracle/post_ref.c)
  * ``MinEntropyLoss`` / gradient                              ops/min_entropy_loss_op.cu:34-66, 70-152

Parity status: the NMS restatement is PINNED against the reference's own Cython NMS compiled
unmodified from /root/reference (``oracle/build_ref_nms.sh`` -> ``oracle/_ref/cython_nms*.so``;
vectors in ``tests/golden/nms_ref.npz``).  The dedup / mean / limit steps are NumPy calls in the
reference and are the same NumPy calls here.  ``core/test_wsl.py`` itself imports caffe2 and cannot
be imported, so ``box_results_with_nms_and_limit`` is a restatement around that pinned NMS.
MinEntropyLoss is GPU-only in the reference (atomics) -> restated, unpinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liboracle_post.so")
_lib = None
F32 = np.float32


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["bash", os.path.join(_HERE, "build_oracle.sh")])
        _lib = ctypes.CDLL(_PATH)
    return _lib


# ----------------------------------------------------------------------------- N1
def flip_boxes(boxes, im_width):
    """utils/boxes.py:246-251."""
    f = boxes.copy()
    f[:, 0::4] = im_width - boxes[:, 2::4] - 1
    f[:, 2::4] = im_width - boxes[:, 0::4] - 1
    return f


def get_rois_blob(im_rois, im_scale, batch_idx=0):
    """core/test_wsl.py:998-1027: float64 product, level column, cast to float32."""
    rois = im_rois.astype(np.float64, copy=False) * im_scale
    levels = np.full((im_rois.shape[0], 1), batch_idx, dtype=np.int64)
    return np.hstack((levels, rois)).astype(np.float32, copy=False)


def dedup_rois(rois, dedup_boxes=1.0 / 16):
    """core/test_wsl.py:125-133.  Returns (index, inv_index)."""
    v = np.array([1, 1e3, 1e6, 1e9, 1e12])
    hashes = np.round(rois * dedup_boxes).dot(v)
    _, index, inv_index = np.unique(hashes, return_index=True, return_inverse=True)
    return index, inv_index.reshape(-1)


def test_cls_prob(rois_pred):
    """modeling/wsl_heads.py:57-67: cls_prob = concat(rois_pred[:, :1], rois_pred)."""
    return np.concatenate([rois_pred[:, :1], rois_pred], axis=1)


def tta_average(scores_ts):
    """core/test_wsl.py:262-263 (SCORE_HEUR 'AVG')."""
    return np.mean(scores_ts, axis=0)


# ----------------------------------------------------------------------------- N2
def nms(dets, thresh):
    """utils/boxes.py:314-318 + utils/cython_nms.pyx:38-93.  dets [n,5] float32."""
    if dets.shape[0] == 0:
        return np.zeros((0,), np.int64)
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    order = np.ascontiguousarray(dets[:, 4].argsort()[::-1], dtype=np.int64)        # cython_nms.pyx:45
    keep = np.empty(dets.shape[0], np.uint8)
    _load().nawsod_oracle_nms(dets.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), dets.shape[0],
                              order.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_float(thresh),
                              keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return np.where(keep != 0)[0]


def nms_invariants_hold(dets, keep, thresh):
    """Greedy-NMS invariants that hold whatever the order of tied scores: survivors do not overlap
    each other >= thresh, and every suppressed box overlaps a survivor of >= score by >= thresh."""
    def iou(a, b):
        w = max(np.float32(0), min(a[2], b[2]) - max(a[0], b[0]) + 1)
        h = max(np.float32(0), min(a[3], b[3]) - max(a[1], b[1]) + 1)
        inter = np.float32(w * h)
        return inter / ((a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - inter)
    keep = list(keep)
    ks = set(keep)
    for i in keep:
        for j in keep:
            if i < j and iou(dets[i], dets[j]) >= thresh:
                return False
    for j in range(dets.shape[0]):
        if j not in ks and not any(dets[i, 4] >= dets[j, 4] and iou(dets[i], dets[j]) >= thresh for i in keep):
            return False
    return True


def reference_nms():
    """The reference's own Cython NMS (oracle/_ref, built by oracle/build_ref_nms.sh) or None."""
    ref = os.path.join(_HERE, "_ref")
    if not any(f.startswith("cython_nms") and f.endswith(".so") for f in (os.listdir(ref) if os.path.isdir(ref) else [])):
        return None
    if not hasattr(np, "int"):
        np.int = int            # numpy-1.x alias the reference's .pyx still calls at run time
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import cython_nms
    return cython_nms.nms


def box_results_with_nms_and_limit(scores, boxes, num_classes, score_thresh=0.05, nms_thresh=0.3,
                                   detections_per_im=100, nms_fn=None):
    """core/test_wsl.py:803-863 with soft-NMS and box voting off (their defaults; the flickr
    configs do not enable them).  boxes [R, 4*num_classes] or [R, 4] (COORD_HEUR 'ID': the same
    box for every class).  Returns (scores, boxes, cls_boxes, keep_mask [num_classes, R])."""
    nms_fn = nms_fn or nms
    R = scores.shape[0]
    cls_boxes = [np.zeros((0, 5), np.float32) for _ in range(num_classes)]
    kept_rows = [np.zeros((0,), np.int64) for _ in range(num_classes)]
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] > score_thresh)[0]                             # :824
        scores_j = scores[inds, j]
        boxes_j = boxes[inds, j * 4:(j + 1) * 4] if boxes.shape[1] > 4 else boxes[inds, :]
        dets_j = np.hstack((boxes_j, scores_j[:, np.newaxis])).astype(np.float32, copy=False)
        keep = nms_fn(dets_j, F32(nms_thresh)) if dets_j.shape[0] else []          # :838
        keep = np.asarray(keep, dtype=np.int64)
        cls_boxes[j] = dets_j[keep, :]
        kept_rows[j] = inds[keep]
    if detections_per_im > 0:                                                      # :852-860
        image_scores = np.hstack([cls_boxes[j][:, -1] for j in range(1, num_classes)])
        if len(image_scores) > detections_per_im:
            image_thresh = np.sort(image_scores)[-detections_per_im]
            for j in range(1, num_classes):
                keep = np.where(cls_boxes[j][:, -1] >= image_thresh)[0]
                cls_boxes[j] = cls_boxes[j][keep, :]
                kept_rows[j] = kept_rows[j][keep]
    im_results = np.vstack([cls_boxes[j] for j in range(1, num_classes)])
    mask = np.zeros((num_classes, R), np.uint8)
    for j in range(1, num_classes):
        mask[j, kept_rows[j]] = 1
    return im_results[:, -1], im_results[:, :-1], cls_boxes, mask


# ----------------------------------------------------------------------------- N4
def min_entropy_loss(X, L):
    """ops/min_entropy_loss_op.cu:34-49,70-104: Y = sum_{L[0,c]>=0.5} -p log p / (1 + count)."""
    X = np.asarray(X, F32)
    sel = ~(np.asarray(L, F32)[0] < F32(0.5))
    prob = np.maximum(X[:, sel], F32(1e-20))
    loss = (-prob * np.log(prob)).astype(F32)
    norm = F32(1.0) + F32(prob.size)
    return F32(loss.astype(np.float64).sum()) / norm, norm


def min_entropy_loss_grad(X, L, dY):
    """ops/min_entropy_loss_op.cu:52-66,106-152."""
    X = np.asarray(X, F32)
    sel = ~(np.asarray(L, F32)[0] < F32(0.5))
    norm = F32(1.0) + F32(X.shape[0] * int(sel.sum()))
    scale = F32(dY) / norm
    prob = np.maximum(X, F32(1e-20))
    d = np.minimum(scale * (F32(-1.0) + F32(-1.0) * np.log(prob)), F32(1e4)).astype(F32)
    d[:, ~sel] = 0
    return d
