"""The oracle against the reference's OWN CUDA kernels (RoIIoU, the in-tree RoI max-pooling clone RoILoopPool,
MinEntropyLoss).  Those kernels are cut unmodified out of /root/reference/detectron/ops/*.cu and executed on the
host (oracle/build_ref_kernels.py); their outputs are committed in tests/golden/ref_kernels.npz
(tests/golden/make_golden_ref_kernels.py).  This pins rows a1/a2 (the arithmetic RoIPoolF shares with the clone),
a7 (RoIIoU) and N4 of SURVEY.md section 8 to reference code instead of to a re-reading of it."""
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import nawsod_oracle as O
from oracle import ref_kernels as RK
from oracle import test_time_oracle as TT


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_kernels.npz"))


def test_roi_iou_restatement_is_the_reference_kernel(gold):
    J = O.roi_iou(gold["iou_rois"])
    assert np.array_equal(J, gold["iou_J"], equal_nan=True)
    assert np.array_equal(np.diag(J), np.ones(J.shape[0], np.float32))


@pytest.mark.parametrize("tag", ["p16", "p8"])
def test_roi_pool_restatement_is_the_reference_kernel(gold, tag):
    """RoILoopPool with its inner rectangle disabled, on a strictly positive map, IS RoIPoolF (the
    clone differs by the rois stride, the inner-rectangle skip and maxval starting at 0): values and argmax of the C
    and the NumPy restatements equal the reference kernel bit for bit, on MCG-like, tiny, full-image, overhanging,
    outside, inverted and half-integer RoIs, at 1/16 and 1/8."""
    X, rois, scale = gold[tag + "_X"], gold[tag + "_rois"], float(gold[tag + "_scale"])
    Y, A = CO.roi_pool_f(X, rois, scale)
    assert np.array_equal(Y, gold[tag + "_Y"]) and np.array_equal(A, gold[tag + "_A"])
    Yn, An = O.roi_pool_f(X, rois[:40], scale)
    assert np.array_equal(Yn, gold[tag + "_Y"][:40]) and np.array_equal(An, gold[tag + "_A"][:40])
    # RoIPoolFGradient == the clone's ROIPoolBackward (same accumulation order here: both run sequentially)
    dX = CO.roi_pool_f_grad(X.shape, rois, A, gold[tag + "_dY"])
    np.testing.assert_allclose(dX, gold[tag + "_dX"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(dX == 0, gold[tag + "_dX"] == 0)


def test_roi_pool_departs_from_the_clone_only_where_survey_says(gold):
    """Post-ReLU-like map (half zeros): values agree; the clone starts maxval at 0 (roi_loop_pool_op.cu:72-74), so a bin
    whose maximum is 0 keeps argmax -1 there, while RoIPoolF (-FLT_MAX start) reports its first cell."""
    X, rois = gold["z_X"], gold["z_rois"]
    Y, A = CO.roi_pool_f(X, rois, 1 / 16)
    assert np.array_equal(Y, gold["z_Y"])
    assert np.array_equal(np.where(Y > 0, A, -1), gold["z_A"])
    assert (A[Y == 0] >= 0).any()


def test_min_entropy_restatement_matches_the_reference_kernels(gold):
    X, L = gold["me_X"], gold["me_L"]
    y, norm = TT.min_entropy_loss(X, L)
    assert float(norm) == 1.0 + float(gold["me_count"])
    # the kernel sums with float atomics, the restatement in double: float accumulation error only
    assert abs(float(y) * float(norm) - float(gold["me_sum"])) <= 1e-5 * abs(float(gold["me_sum"]))
    d = TT.min_entropy_loss_grad(X, L, gold["me_dY"])
    np.testing.assert_allclose(d, gold["me_dX"], rtol=2e-6, atol=0)       # logf: glibc vs NumPy, last ulp
    assert np.array_equal(d == 0, gold["me_dX"] == 0)


@pytest.mark.skipif(not RK.available(), reason="oracle/_ref/libnawsod_ref_kernels.so not built (needs /root/reference)")
def test_golden_vectors_are_what_the_reference_kernels_return(gold):
    """Freshness of the fixtures: re-run the reference kernels wherever the library is present."""
    assert np.array_equal(RK.roi_iou(gold["iou_rois"]), gold["iou_J"], equal_nan=True)
    for tag in ("p16", "p8"):
        Y, A = RK.roi_loop_pool(gold[tag + "_X"], RK.rois9(gold[tag + "_rois"]), float(gold[tag + "_scale"]))
        assert np.array_equal(Y, gold[tag + "_Y"]) and np.array_equal(A, gold[tag + "_A"])
    s, cnt = RK.min_entropy_forward_kernel(gold["me_X"], gold["me_L"])
    assert s == gold["me_sum"] and cnt == gold["me_count"]
