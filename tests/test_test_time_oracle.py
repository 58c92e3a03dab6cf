"""CPU suite for the test-time oracle: the NMS restatement against the golden
vectors produced by the reference's own Cython NMS (and against that code live when oracle/_ref holds
it), the NumPy steps of the reference's dedup / TTA mean, and MinEntropyLoss against finite differences."""
import os

import numpy as np
import pytest

from oracle import test_time_oracle as T


_valid_nms = T.nms_invariants_hold


def test_nms_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_ref.npz"))
    for k in range(int(g["n_cases"])):
        dets, th, want = g["dets%d" % k], g["thresh%d" % k], g["keep%d" % k]
        got = T.nms(dets, th)
        if int(g["ties%d" % k]) and dets.shape[0] > 16:
            # NumPy's introsort leaves the order of equal scores to the CPU's sort kernel
            assert _valid_nms(dets, got, th) and _valid_nms(dets, want, th)
        else:
            assert np.array_equal(got, want), k


def test_nms_matches_reference_live():
    ref = T.reference_nms()
    if ref is None:
        pytest.skip("oracle/_ref/cython_nms is not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = int(rng.integers(1, 500))
        x1 = rng.integers(0, 600, n).astype(np.float32) + (rng.random(n).astype(np.float32) if trial % 2 else 0)
        y1 = rng.integers(0, 400, n).astype(np.float32)
        dets = np.stack([x1, y1, x1 + rng.integers(1, 300, n), y1 + rng.integers(1, 300, n),
                         rng.random(n)], axis=1).astype(np.float32)
        th = np.float32([0.3, 0.5, 0.7][trial % 3])
        assert np.array_equal(T.nms(dets, th), ref(dets, th))


def test_dedup_and_scatter_roundtrip():
    rng = np.random.default_rng(0)
    boxes = np.floor(rng.random((600, 4)) * 300).astype(np.float32)
    boxes[:, 2:] += boxes[:, :2]
    boxes[100:200] = boxes[:100] + rng.integers(0, 3, (100, 4))           # near-duplicates: same feature RoI
    rois = T.get_rois_blob(boxes, 1.376)
    index, inv = T.dedup_rois(rois)
    assert len(index) < 600 and np.array_equal(np.round(rois[index] / 16)[inv], np.round(rois / 16))
    assert np.all(np.diff(index[np.argsort(index)]) > 0)
    f = T.flip_boxes(boxes, 500)
    assert np.array_equal(T.flip_boxes(f, 500), boxes)                   # flipping twice is the identity
    s = [rng.random((50, 21)).astype(np.float32) for _ in range(10)]
    acc = s[0].copy()
    for x in s[1:]:
        acc = acc + x
    assert np.array_equal(acc / np.float32(10), T.tta_average(s))        # np.mean(axis=0) = ordered float32 sum / T


def test_limit_keeps_top_scores():
    rng = np.random.default_rng(3)
    R, K1 = 300, 6
    scores = ((rng.permutation(R * K1).reshape(R, K1) + 1) / (R * K1 + 1)).astype(np.float32)
    x1 = rng.integers(0, 500, R).astype(np.float32)
    y1 = rng.integers(0, 300, R).astype(np.float32)
    boxes = np.stack([x1, y1, x1 + rng.integers(5, 200, R), y1 + rng.integers(5, 200, R)], 1).astype(np.float32)
    s, b, cls_boxes, mask = T.box_results_with_nms_and_limit(scores, boxes, K1, 0.05, 0.5, 40)
    assert 40 <= len(s) and mask.sum() == len(s) and mask[0].sum() == 0
    s_all, _, _, mask_all = T.box_results_with_nms_and_limit(scores, boxes, K1, 0.05, 0.5, 0)
    assert len(s) == 40 and np.array_equal(np.sort(s_all)[-40:], np.sort(s))
    assert np.all(mask <= mask_all)


def test_min_entropy_loss_gradient_check():
    """Numeric gradient of the forward restatement (thresholds of the reference's own gradient tests,
    detectron/tests/test_smooth_l1_loss_op.py:51-55: stepsize 0.005, threshold 0.005)."""
    rng = np.random.default_rng(2)
    X = (rng.random((12, 6)) * 0.8 + 0.1).astype(np.float32)
    L = np.array([[1, 0, 1, 0, 0, 1]], np.float32)
    y, norm = T.min_entropy_loss(X, L)
    assert norm == 1 + 12 * 3
    d = T.min_entropy_loss_grad(X, L, np.float32(1.0))
    assert np.all(d[:, [1, 3, 4]] == 0)
    for (n, c) in [(0, 0), (5, 2), (11, 5)]:
        Xp, Xm = X.copy(), X.copy()
        Xp[n, c] += 0.005
        Xm[n, c] -= 0.005
        num = (float(T.min_entropy_loss(Xp, L)[0]) - float(T.min_entropy_loss(Xm, L)[0])) / 0.01
        assert abs(num - d[n, c]) <= 0.005 * max(1.0, abs(num))
