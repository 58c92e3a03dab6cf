"""The test-time oracle against the reference's OWN test driver (tests/golden/test_wsl.npz, made by
tests/golden/make_golden_test_wsl.py): detectron/core/test_wsl.py's im_detect_bbox / im_detect_bbox_aug /
box_results_with_nms_and_limit were imported unmodified and run with the shipped flickr_voc config against a stand-in
workspace whose net is a deterministic pseudo head.  Pins rows N1 / N2 of SURVEY.md section 8f -- projection to the input
scale, the float64 hash dedup (DEDUP_BOXES 0.125) and its inverse map, flipping, the ten passes of the flickr test-time
augmentation in the reference's order, float32 score averaging, thresholding + NMS + the detections-per-image limit --
to reference code.  The GPU counterpart is tests/test_gpu_zzz_reference_vectors.py::test_tta_and_nms_vs_reference_driver."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import roi_data_oracle as RD
from oracle import test_time_oracle as T

HERE = os.path.dirname(os.path.abspath(__file__))


def _pseudo():
    spec = importlib.util.spec_from_file_location("make_golden_test_wsl", os.path.join(HERE, "golden", "make_golden_test_wsl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                 # imports NumPy only; /root/reference is touched by its main() alone
    return mod.pseudo_cls_prob


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "test_wsl.npz"))


def reference_passes(gold, i):
    """(target scale, max size, flip) in the order of core/test_wsl.py:211-256: the flipped image at the test scale, every
    augmentation scale followed by its flip, the identity transform last."""
    passes = [(int(gold["test_scale"]), int(gold["test_max_size"]), True)]
    for s in gold["aug_scales"]:
        passes += [(int(s), int(gold["aug_max_size"]), False), (int(s), int(gold["aug_max_size"]), True)]
    return passes + [(int(gold["test_scale"]), int(gold["test_max_size"]), False)]


def one_pass(gold, i, target, max_size, flip, pseudo):
    h, w = (int(v) for v in gold["case%d_im_shape" % i][:2])
    boxes, obn = gold["case%d_boxes" % i], gold["case%d_obn" % i]
    im_scale = RD.im_scale_for(h, w, target, max_size)
    b = T.flip_boxes(boxes, w) if flip else boxes
    rois = T.get_rois_blob(b, im_scale)
    index, inv = T.dedup_rois(rois, float(gold["dedup_boxes"]))
    obn1 = np.add(obn, 1.0)                                              # core/test_wsl.py:1058
    scores = pseudo(rois[index], obn1[index], int(gold["num_classes"]))[inv]
    return scores, rois[index], obn1[index], im_scale


@pytest.mark.parametrize("i", [0, 1, 2])
def test_single_pass(gold, i):
    pseudo = _pseudo()
    scores, fed_rois, fed_obn, im_scale = one_pass(gold, i, int(gold["test_scale"]), int(gold["test_max_size"]), False, pseudo)
    assert im_scale == float(gold["case%d_single_im_scale" % i])
    assert np.array_equal(fed_rois, gold["case%d_single_fed_rois" % i]) and fed_rois.dtype == np.float32
    assert np.array_equal(fed_obn, gold["case%d_single_fed_obn" % i])
    assert np.array_equal(scores, gold["case%d_single_scores" % i])
    boxes = gold["case%d_boxes" % i]
    assert fed_rois.shape[0] < boxes.shape[0] or boxes.shape[0] == 1     # the cases do contain colliding boxes
    # BBOX_REG off: the predicted boxes are the ORIGINAL proposals tiled per class (core/test_wsl.py:169-171)
    assert np.array_equal(gold["case%d_single_boxes" % i], np.tile(boxes, (1, int(gold["num_classes"]))))


@pytest.mark.parametrize("i", [0, 1, 2])
def test_augmented_passes_and_average(gold, i):
    pseudo = _pseudo()
    passes = reference_passes(gold, i)
    assert len(passes) == int(gold["case%d_aug_passes" % i]) == 10
    outs = [one_pass(gold, i, t, m, f, pseudo) for t, m, f in passes]
    assert [o[1].shape[0] for o in outs] == list(gold["case%d_aug_fed_counts" % i])
    avg = T.tta_average([o[0] for o in outs])
    assert avg.dtype == np.float32 and np.array_equal(avg, gold["case%d_aug_scores" % i])
    # the data blob of each pass has the size the scale rule gives (the resize itself is outside the path)
    h, w = (int(v) for v in gold["case%d_im_shape" % i][:2])
    for (t, m, f), shp in zip(passes, gold["case%d_aug_data_shapes" % i]):
        s = RD.im_scale_for(h, w, t, m)
        assert abs(shp[2] - h * s) <= 1 and abs(shp[3] - w * s) <= 1


@pytest.mark.parametrize("i", [0, 1, 2])
def test_nms_and_limit(gold, i):
    sc, bx, cls_boxes, mask = T.box_results_with_nms_and_limit(
        gold["case%d_aug_scores" % i], gold["case%d_aug_boxes" % i], int(gold["num_classes"]),
        score_thresh=float(gold["score_thresh"]), nms_thresh=float(gold["nms"]), detections_per_im=int(gold["detections_per_im"]))
    assert np.array_equal(sc, gold["case%d_det_scores" % i]) and np.array_equal(bx, gold["case%d_det_boxes" % i])
    assert [len(c) for c in cls_boxes] == list(gold["case%d_det_counts" % i])
    assert sc.shape[0] <= int(gold["detections_per_im"]) or len(set(sc)) < sc.shape[0]     # ties at the threshold may exceed the limit
