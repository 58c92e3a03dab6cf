"""TEST INFRASTRUCTURE -- CPU restatement of the reference's training-input contract for the head
(SURVEY.md 8f row N3): MCG proposals -> ``rois[R,5]`` / ``obn_scores`` / ``labels_oh`` blobs, crop
offsets, and the bagging-mixup of two webly images.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this module; the product (na-fwebsod_b200/) never does.

Pinned by the reference's OWN Python: tests/golden/make_golden_roi_data.py imports
/root/reference/detectron/roi_data/wsl.py (with stand-ins for the absent caffe2 / future / Cython
modules, none of which the functions below touch) and stores what ``_project_im_rois``, ``_sample_rois``
and ``add_wsl_blobs`` return in tests/golden/roi_data.npz.  The two pieces that are not importable as
functions (the mixup block inside ``RoIDataLoader.get_next_minibatch`` and the ``__main__`` body of
tools/convert_mcg.py) are restated line by line and marked so.

Citations are relative to /root/reference.
"""
import numpy as np


def convert_mcg_boxes(bboxes):
    """tools/convert_mcg.py:37-49 (restated; the script has no importable function).
    ``bboxes``: the .mat array, 1-indexed (y1, x1, y2, x2), any real dtype.
    ``astype(np.uint16) - 1`` then the column order (1, 0, 3, 2) -> 0-indexed (x1, y1, x2, y2) uint16;
    a 0 coordinate wraps to 65535 exactly like the uint16 subtraction of the script."""
    b = np.asarray(bboxes).astype(np.uint16) - np.uint16(1)
    return b[:, (1, 0, 3, 2)].astype(np.uint16)


def im_scale_for(im_h, im_w, target_size, max_size):
    """detectron/utils/blob.py:117-122 (prep_im_for_blob): the scale that maps the short side to
    ``target_size`` unless the long side would exceed ``max_size``."""
    im_size_min, im_size_max = min(im_h, im_w), max(im_h, im_w)
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    return im_scale


def random_crop(im_h, im_w, crop, r0, r1):
    """detectron/roi_data/minibatch_wsl.py:142-153 (WSL.USE_CROP) with the two uniform draws passed in.
    Returns im_crop as the reference builds it, (row1, col1, row2, col2) int32 (truncation)."""
    shape = np.array([im_h, im_w])
    crop_dims = shape * crop
    s = shape - crop_dims
    s[0] *= r0
    s[1] *= r1
    return np.array([s[0], s[1], s[0] + crop_dims[0] - 1, s[1] + crop_dims[1] - 1], dtype=np.int32)


def crops_to_xyxy(im_crops):
    """detectron/roi_data/minibatch_wsl.py:63-64: (row, col, row, col) -> (x1, y1, x2, y2) int32."""
    return np.array(im_crops, dtype=np.int32).reshape(-1, 4)[:, (1, 0, 3, 2)]


def project_im_rois(im_rois, im_scale_factor, im_crop):
    """detectron/roi_data/wsl.py:212-225.  Clip the boxes to the crop window (x1,y1 clipped from below
    first, x2,y2 from above first), shift by the crop origin, scale.  float32 boxes minus an int32 tile is
    float64 in NumPy, times a Python float: the result is float64 and is cast to float32 by the caller
    (wsl.py:160)."""
    r = np.array(im_rois, dtype=np.float32, copy=True)
    c = [float(v) for v in im_crop]
    r[:, 0] = np.minimum(np.maximum(r[:, 0], c[0]), c[2])
    r[:, 1] = np.minimum(np.maximum(r[:, 1], c[1]), c[3])
    r[:, 2] = np.maximum(np.minimum(r[:, 2], c[2]), c[0])
    r[:, 3] = np.maximum(np.minimum(r[:, 3], c[3]), c[1])
    off = np.array([c[0], c[1], c[0], c[1]], dtype=np.float64)
    return (r.astype(np.float64) - off) * float(im_scale_factor)


def sample_rois(boxes, obn_scores, gt_classes, im_scale, im_crop, batch_idx, rois_per_image, num_classes):
    """detectron/roi_data/wsl.py:87-181 (_sample_rois; the live ``else`` branch :101-104, the two
    ``np.delete`` calls :115-116 discard their result, so ground-truth rows stay).
    boxes [n,4] float32, obn_scores [n,1] float32, gt_classes [n] int (0 = proposal),
    im_crop (x1,y1,x2,y2).  Returns the blob dict of :157-162."""
    n = int(min(int(rois_per_image), boxes.shape[0]))
    sampled_scores = np.add(np.asarray(obn_scores, dtype=np.float32)[:n], 1.0).astype(np.float32)
    rois = project_im_rois(boxes[:n], im_scale, im_crop)
    rois = np.hstack((batch_idx * np.ones((n, 1), dtype=np.float32), rois)).astype(np.float32)
    labels_oh = np.zeros((1, num_classes - 1), dtype=np.float32)
    labels = np.zeros((1,), dtype=np.float32)
    gt = np.asarray(gt_classes)
    gt = gt[gt > 0]
    if gt.size == 0:
        raise AssertionError("Empty ground truth empty for image is not allowed. Please check.")
    for g in gt:
        labels_oh[0, int(g) - 1] = 1
        labels[0] = int(g) - 1                  # the last ground-truth class wins (:153-155)
    return dict(labels_int32=labels.astype(np.int32), labels_oh=labels_oh, rois=rois, obn_scores=sampled_scores)


def add_wsl_blobs(roidb, im_scales, im_crops, rois_per_image, num_classes):
    """detectron/roi_data/wsl.py:59-85: per-image blobs concatenated along axis 0; image i's rows carry
    batch index i.  ``roidb``: list of dicts with boxes / obn_scores / gt_classes."""
    out = {}
    for i, e in enumerate(roidb):
        b = sample_rois(e["boxes"], e["obn_scores"], e["gt_classes"], im_scales[i], im_crops[i], i, rois_per_image, num_classes)
        for k, v in b.items():
            out.setdefault(k, []).append(v)
    return {k: np.concatenate(v) for k, v in out.items()}


def bagging_mixup(blobs, lam):
    """detectron/roi_data/loader_wsl.py:149-168 (restated: the block lives inside a thread-owning method).
    Two images of one class become ONE training image: data and labels_oh are the convex combination
    (lam, 1 - lam), all RoIs of both images get batch index 0, the per-image blobs keep image 0's entry.
    The products are float32 (float64 scalar x float32 array under the NumPy 1.x casting the reference
    ran on: the scalar is cast to the array's type) and accumulate into a float32 zero array."""
    lams = [np.float32(lam), np.float32(1.0 - lam)]
    out = dict(blobs)
    for k in ("data", "labels_oh"):
        if k not in blobs:
            continue
        src = np.asarray(blobs[k], dtype=np.float32)
        acc = np.zeros((1,) + src.shape[1:], dtype=np.float32)
        for i in range(2):
            acc += lams[i] * src[i:i + 1]
        out[k] = acc
    rois = np.array(blobs["rois"], dtype=np.float32, copy=True)
    rois[:, 0] = 0
    out["rois"] = rois
    for k in ("data_ids", "labels_int32"):
        if k in blobs:
            out[k] = blobs[k][0:1]
    return out
