"""Host-side mirror of the reference's test-time wrapper around the head (SURVEY.md 8f, N1 + N2).

The reference (detectron/core/test_wsl.py) prepares the blobs on the host with NumPy, feeds them,
runs the net, fetches ``cls_prob`` back and post-processes on the CPU (NumPy + Cython NMS) -- ten
host round trips per image with the flickr test-time augmentation (5 scales x {orig, hflip}).  Here
the same-named functions keep every step on the GPU: projection / flip / de-duplication, the head
forward, the inverse scatter and the TTA running sum are stream-ordered libnawsod calls; only the
final keep mask of the NMS is read back to assemble the reference's ``cls_boxes`` lists.

The conv body is out of scope: where the reference takes the image ``im``, these
functions take the conv5 map(s) of that image at the requested scale, channels-last in the model's
dtype (or NCHW float32 with ``x_layout='NCHW'``).
"""
from __future__ import annotations

import torch

from . import ops


def im_detect_bbox(model, conv5, im_scale, boxes, obn_scores, *, dedup_boxes=1.0 / 16, flip_width=None,
                   x_layout="NHWC", sync=True, out=None, accumulate=False):
    """``im_detect_bbox`` (core/test_wsl.py:100-178) for one image and one scale.

    boxes [R,4] float32 proposals in ORIGINAL image coordinates, obn_scores [R] or [R,1] (raw; the +1
    of core/test_wsl.py:1058 happens here), ``flip_width``: the original image width when ``conv5`` is
    the map of the horizontally flipped image (``im_detect_bbox_hflip``, core/test_wsl.py:284-307).
    Returns scores [R, num_classes] (column 0 duplicates column 1, modeling/wsl_heads.py:57-67) for the
    ORIGINAL boxes; with ``accumulate`` the scores are added into ``out``.

    ``sync=True`` reads the unique-RoI count back (4 bytes) and runs the head on exactly that many rows,
    like the reference; ``sync=False`` skips the round trip: the head runs on all R rows (the tail
    repeats the first unique RoI) with the device-side row range {0, num_unique}, so the RoI-axis softmax
    still sees each unique RoI once.
    """
    R = boxes.shape[0]
    rois, obn1 = ops.project_rois(boxes, im_scale, flip_width=flip_width, obn_scores=obn_scores.reshape(-1))
    if dedup_boxes and dedup_boxes > 0:
        index, inv_index, num_unique, offsets = ops.dedup_rois(rois, dedup_boxes)
        n = int(num_unique.item()) if sync else R
        rois_u = ops.gather_rows(rois, index, n)
        obn_u = ops.gather_rows(obn1.view(R, 1), index, n).view(n)
        if sync:
            offsets = None                      # exactly n rows: the default one-image range
    else:
        inv_index, offsets, rois_u, obn_u = None, None, rois, obn1
    model.FeedBlobs(conv5, rois_u, obn_u, roi_offsets=offsets, x_layout=x_layout)
    model.RunTestNet(want_cls_prob=False)
    return ops.scatter_scores(model.blobs["rois_pred"], inv_index, R=R, out=out, accumulate=accumulate)


def im_detect_bbox_aug(model, passes, boxes, obn_scores, *, dedup_boxes=1.0 / 16, x_layout="NHWC", sync=True):
    """``im_detect_bbox_aug`` with SCORE_HEUR 'AVG' / COORD_HEUR 'ID' (core/test_wsl.py:181-281; flickr
    configs: TEST.BBOX_AUG H_FLIP, SCALES (480, 576, 864, 1200), SCALE_H_FLIP).

    ``passes``: list of (conv5, im_scale, flip_width_or_None) in the reference's order -- the flipped
    image at the test scale, then every extra scale (each followed by its flip), the identity
    transform LAST (core/test_wsl.py:211-256).  Returns the averaged scores [R, num_classes]; the sum
    runs over the passes in that order in float32 and is divided by their number at the end, which is
    what ``np.mean(scores_ts, axis=0)`` computes."""
    if not passes:
        raise RuntimeError("im_detect_bbox_aug: no passes")
    acc = None
    for conv5, im_scale, flip_width in passes:
        acc = im_detect_bbox(model, conv5, im_scale, boxes, obn_scores, dedup_boxes=dedup_boxes, flip_width=flip_width,
                             x_layout=x_layout, sync=sync, out=acc, accumulate=acc is not None)
    return ops.scores_finalize(acc, len(passes))


def box_results_with_nms_and_limit(scores, boxes, *, score_thresh=0.05, nms_thresh=0.3, detections_per_im=100):
    """``box_results_with_nms_and_limit`` (core/test_wsl.py:803-863; soft-NMS and box voting are off in
    the flickr configs).  scores [R, num_classes] and boxes [R,4] are CUDA tensors; the threshold, the
    per-class greedy NMS and the detections-per-image limit run on the GPU, and the reference's return
    values ``(scores, boxes, cls_boxes)`` are assembled from the keep mask (cls_boxes[j]: [n_j, 5]
    rows of x1, y1, x2, y2, score; cls_boxes[0] is empty).  Row order within a class: ASCENDING proposal row, which is
    the reference's -- its Cython NMS returns ``np.where(suppressed == 0)[0]`` (utils/cython_nms.pyx:93), i.e. the kept
    indices sorted ascending, not the descending-score visiting order, ``dets_j[keep]`` keeps that order
    (core/test_wsl.py:838-846) and the detections-per-image filter preserves it (:852-860); pinned bit for bit by
    tests/test_test_wsl_golden.py against the reference's own driver."""
    keep, num_keep, _ = ops.nms_and_limit(scores, boxes, score_thresh=score_thresh, nms_thresh=nms_thresh,
                                          detections_per_im=detections_per_im)
    K1 = scores.shape[1]
    cls_idx, row_idx = torch.nonzero(keep, as_tuple=True)          # ascending (class, row): np.where(suppressed == 0) per class, vstack over classes
    dets = torch.cat([boxes[row_idx], scores[row_idx, cls_idx].unsqueeze(1)], dim=1)
    counts = num_keep.cpu().tolist()
    cls_boxes, start = [], 0
    for j in range(K1):
        cls_boxes.append(dets[start:start + counts[j]])
        start += counts[j]
    return dets[:, 4], dets[:, :4], cls_boxes
