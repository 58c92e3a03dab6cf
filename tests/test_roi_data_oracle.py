"""CPU: the oracle's restatement of the training-input contract against what the
reference's own Python returned (tests/golden/roi_data.npz, made by tests/golden/make_golden_roi_data.py)."""
import os

import numpy as np
import pytest

from oracle import roi_data_oracle as RD


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "roi_data.npz"))


def test_project_im_rois_matches_reference(gold):
    boxes = gold["project_boxes"]
    for i in range(int(gold["project_cases"])):
        got = RD.project_im_rois(boxes, float(gold["project_scale_%d" % i]), gold["project_crop_%d" % i])
        assert got.dtype == np.float64
        assert np.array_equal(got, gold["project_out_%d" % i])                 # bit-exact, float64
        assert np.array_equal(got.astype(np.float32), gold["project_out32_%d" % i])
        assert np.array_equal(boxes, gold["project_boxes"])                    # the oracle does not clobber its input


def test_add_wsl_blobs_matches_reference(gold):
    roidb = [dict(boxes=gold["mb_boxes_%d" % i], obn_scores=gold["mb_obn_scores_%d" % i], gt_classes=gold["mb_gt_classes_%d" % i])
             for i in range(2)]
    blobs = RD.add_wsl_blobs(roidb, gold["mb_im_scales"], gold["mb_im_crops"], int(gold["mb_rois_per_image"]), int(gold["mb_num_classes"]))
    for k in ("rois", "obn_scores", "labels_int32", "labels_oh"):
        ref = gold["mb_out_" + k]
        assert blobs[k].dtype == ref.dtype and blobs[k].shape == ref.shape, k
        assert np.array_equal(blobs[k], ref), k
    # image 1 has fewer boxes than BATCH_SIZE_PER_IM: all of them are kept; rows are contiguous per image
    assert np.array_equal(np.bincount(blobs["rois"][:, 0].astype(int)), [300, 180])
    assert blobs["labels_oh"][1].sum() in (1.0, 2.0)


def test_scale_rule_matches_reference(gold):
    for h, w, target, max_size, s in gold["scale_cases"]:
        assert RD.im_scale_for(int(h), int(w), int(target), int(max_size)) == s


def test_sample_rois_requires_ground_truth():
    with pytest.raises(AssertionError):
        RD.sample_rois(np.zeros((3, 4), np.float32), np.zeros((3, 1), np.float32), np.zeros(3, np.int32), 1.0, [0, 0, 9, 9], 0, 10, 21)


def test_random_crop_and_reorder():
    c = RD.random_crop(375, 500, 0.9, 0.25, 0.75)
    assert c.dtype == np.int32
    # (row1, col1, row2, col2): 37.5*0.25 = 9.375 -> 9; 50*0.75 = 37.5 -> 37; + 337.5 - 1, + 450 - 1 (truncated)
    assert c.tolist() == [9, 37, 345, 486]
    assert RD.crops_to_xyxy([c]).tolist() == [[37, 9, 486, 345]]


def test_bagging_mixup_semantics():
    rng = np.random.default_rng(5)
    data = rng.standard_normal((2, 3, 8, 9)).astype(np.float32)
    L = np.zeros((2, 20), np.float32)
    L[0, 3] = 1
    L[1, 3] = 1
    L[1, 7] = 1
    rois = np.concatenate([np.c_[np.zeros(5), rng.random((5, 4))], np.c_[np.ones(7), rng.random((7, 4))]]).astype(np.float32)
    lam = 0.37123456789
    out = RD.bagging_mixup(dict(data=data, labels_oh=L, rois=rois, labels_int32=np.array([3, 7], np.int32)), lam)
    assert out["data"].shape == (1, 3, 8, 9) and out["data"].dtype == np.float32
    l0, l1 = np.float32(lam), np.float32(1 - lam)
    assert np.array_equal(out["data"][0], l0 * data[0] + l1 * data[1])
    assert np.array_equal(out["labels_oh"][0, [3, 7]], [l0 + l1, l1])
    assert np.all(out["rois"][:, 0] == 0) and np.array_equal(out["rois"][:, 1:], rois[:, 1:])
    assert out["labels_int32"].tolist() == [3]


def test_convert_mcg_boxes():
    mat = np.array([[1, 1, 375, 500], [12.0, 30.0, 40.0, 77.0], [3.9, 2.2, 8.7, 9.1]], np.float64)     # 1-indexed (y1,x1,y2,x2)
    out = RD.convert_mcg_boxes(mat)
    assert out.dtype == np.uint16
    assert out.tolist() == [[0, 0, 499, 374], [29, 11, 76, 39], [1, 2, 8, 7]]
    assert RD.convert_mcg_boxes(np.array([[0, 5, 6, 7]], np.float64))[0, 1] == 65535            # uint16 wrap, like the script
