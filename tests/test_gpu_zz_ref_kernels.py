"""GPU kernels against the outputs of the reference's OWN CUDA kernels (tests/golden/ref_kernels.npz: RoIIoU, the
in-tree RoI max-pooling clone with its inner rectangle disabled, MinEntropyLoss -- see tests/test_ref_kernels.py for
how the vectors are made).  Through the C ABI, like every other GPU test."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from nafwebsod_b200 import ops
    return ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_kernels.npz"))


def test_roi_iou_equals_reference_kernel(gold):
    J = _ops().RoIIoU(dev(gold["iou_rois"])).cpu().numpy()
    assert np.array_equal(J, gold["iou_J"], equal_nan=True)


@pytest.mark.parametrize("tag", ["p16", "p8"])
def test_roi_pool_equals_reference_kernel(gold, tag):
    ops = _ops()
    X, rois, scale = gold[tag + "_X"], gold[tag + "_rois"], float(gold[tag + "_scale"])
    Y, A = ops.RoIPoolF(dev(X), dev(rois), spatial_scale=scale)              # reference layout in and out
    assert np.array_equal(Y.cpu().numpy(), gold[tag + "_Y"])
    assert np.array_equal(A.cpu().numpy(), gold[tag + "_A"])
    dX = ops.RoIPoolFGradient(dev(X), dev(rois), dev(gold[tag + "_A"]), dev(gold[tag + "_dY"]), layout="NCHW").cpu().numpy()
    # atomics: summation order differs from the host run of the reference kernel -> fp32 rounding only
    np.testing.assert_allclose(dX, gold[tag + "_dX"], rtol=1e-4, atol=1e-4)
    assert np.array_equal(dX == 0, gold[tag + "_dX"] == 0)


def test_min_entropy_loss_equals_reference_kernels(gold):
    ops = _ops()
    X, L = gold["me_X"], gold["me_L"]
    want = float(gold["me_sum"]) / (1.0 + float(gold["me_count"]))          # min_entropy_loss_op.cu:70-104
    y = ops.MinEntropyLoss(dev(X), dev(L)).item()
    assert abs(y - want) <= 1e-5 * abs(want)
    d = ops.MinEntropyLossGradient(dev(X), dev(L), torch.tensor([float(gold["me_dY"])], device="cuda")).cpu().numpy()
    np.testing.assert_allclose(d, gold["me_dX"], rtol=1e-5, atol=1e-9)
    assert np.array_equal(d == 0, gold["me_dX"] == 0)
