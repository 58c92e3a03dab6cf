"""GPU parity of the training-input contract, through the C ABI: bit-exact against the golden
vectors the reference's own Python produced (tests/golden/roi_data.npz) and against oracle/roi_data_oracle.py
on larger seeded inputs.  Everything here is float32 / integer work with a stated rounding order, so the bar is
bit equality."""
import os

import numpy as np
import pytest
import torch

from oracle import roi_data_oracle as RD

pytestmark = pytest.mark.gpu


def _mods():
    from nafwebsod_b200 import ops, roi_data
    return ops, roi_data


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "roi_data.npz"))


def test_project_matches_reference_golden(gold):
    ops, _ = _mods()
    boxes = dev(gold["project_boxes"])
    for i in range(int(gold["project_cases"])):
        rois = ops.sample_rois(boxes, float(gold["project_scale_%d" % i]), gold["project_crop_%d" % i], batch_idx=3)
        got = rois.cpu().numpy()
        assert np.array_equal(got[:, 1:], gold["project_out32_%d" % i]), i
        assert np.all(got[:, 0] == 3)
    assert np.array_equal(boxes.cpu().numpy(), gold["project_boxes"])            # inputs untouched


def test_add_wsl_blobs_matches_reference_golden(gold):
    _, roi_data = _mods()
    roidb = [dict(boxes=dev(gold["mb_boxes_%d" % i]), obn_scores=dev(gold["mb_obn_scores_%d" % i]),
                  gt_classes=gold["mb_gt_classes_%d" % i]) for i in range(2)]
    blobs = {}
    assert roi_data.add_wsl_blobs(blobs, gold["mb_im_scales"].tolist(), gold["mb_im_crops"], roidb,
                                  rois_per_image=int(gold["mb_rois_per_image"]), num_classes=int(gold["mb_num_classes"]))
    for k in ("rois", "obn_scores", "labels_int32", "labels_oh"):
        ref = gold["mb_out_" + k]
        got = blobs[k].cpu().numpy()
        assert got.dtype == ref.dtype and got.shape == ref.shape, k
        assert np.array_equal(got, ref), k
    assert blobs["roi_offsets"].cpu().tolist() == [0, 300, 480]
    # gt_classes as a device tensor: same blobs, no host check
    roidb_d = [dict(e, gt_classes=dev(np.asarray(e["gt_classes"], np.int32))) for e in roidb]
    blobs_d = {}
    roi_data.add_wsl_blobs(blobs_d, gold["mb_im_scales"].tolist(), gold["mb_im_crops"], roidb_d,
                           rois_per_image=int(gold["mb_rois_per_image"]), num_classes=int(gold["mb_num_classes"]))
    for k in ("rois", "obn_scores", "labels_int32", "labels_oh"):
        assert torch.equal(blobs[k], blobs_d[k]), k


@pytest.mark.parametrize("R", [1, 255, 2000, 8000])
def test_sample_rois_vs_oracle(R):
    ops, _ = _mods()
    rng = np.random.default_rng(R)
    W, H = 1333, 999
    x1, y1 = rng.integers(0, W - 2, R), rng.integers(0, H - 2, R)
    boxes = np.stack([x1, y1, np.minimum(x1 + rng.integers(1, W, R), W - 1), np.minimum(y1 + rng.integers(1, H, R), H - 1)], 1).astype(np.float32)
    obn = rng.random((R, 1)).astype(np.float32)
    crop = RD.crops_to_xyxy([RD.random_crop(H, W, 0.9, rng.random(), rng.random())])[0]
    scale = RD.im_scale_for(int(crop[3] - crop[1] + 1), int(crop[2] - crop[0] + 1), 688, 2000)
    rois, s1 = ops.sample_rois(dev(boxes), scale, crop, batch_idx=1, obn_scores=dev(obn))
    ref = RD.sample_rois(boxes, obn, np.array([5] + [0] * (R - 1)), scale, crop, 1, R, 21)
    assert np.array_equal(rois.cpu().numpy(), ref["rois"])
    assert np.array_equal(s1.cpu().numpy(), ref["obn_scores"])
    # projected boxes stay inside the rescaled crop and keep x1 <= x2, y1 <= y2
    r = rois.cpu().numpy()
    assert r[:, 1:].min() >= 0 and np.all(r[:, 3] >= r[:, 1]) and np.all(r[:, 4] >= r[:, 2])
    assert r[:, 3].max() <= (crop[2] - crop[0]) * scale + 1e-3 and r[:, 4].max() <= (crop[3] - crop[1]) * scale + 1e-3


def test_image_labels_and_errors():
    ops, roi_data = _mods()
    gt = np.zeros(100, np.int32)
    gt[[0, 7, 50]] = [4, 20, 4]
    oh, li = ops.image_labels(dev(gt), 21)
    want = np.zeros((1, 20), np.float32)
    want[0, [3, 19]] = 1
    assert np.array_equal(oh.cpu().numpy(), want) and li.cpu().tolist() == [3]      # the LAST ground-truth row wins
    oh, li = ops.image_labels(dev(np.zeros(5, np.int32)), 21)
    assert oh.sum().item() == 0 and li.cpu().tolist() == [-1]
    e = dict(boxes=dev(np.zeros((3, 4), np.float32)), obn_scores=dev(np.zeros((3, 1), np.float32)), gt_classes=np.zeros(3, np.int32))
    with pytest.raises(RuntimeError):              # the reference asserts on an entry without ground truth
        roi_data._sample_rois(e, 1.0, [0, 0, 9, 9], 0, rois_per_image=10, num_classes=21)
    with pytest.raises(RuntimeError):              # empty crop window
        ops.sample_rois(e["boxes"], 1.0, [5, 5, 4, 9])
    with pytest.raises(RuntimeError):
        ops.sample_rois(torch.zeros(3, 4), 1.0, [0, 0, 9, 9])     # CPU tensor: no fallback


@pytest.mark.parametrize("shape", [(2, 3, 37, 53), (2, 3, 608, 800), (2, 20)])
def test_bagging_mixup_bit_exact(shape):
    ops, _ = _mods()
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape).astype(np.float32)
    x.reshape(2, -1)[0, :3] = [-0.0, 0.0, -1e-40]                   # signed zeros / a denormal
    x.reshape(2, -1)[1, :3] = [0.0, -0.0, 0.0]
    for lam in (0.5, 0.37123456789, 1e-3, 0.9999):
        got = ops.bagging_mixup(dev(x), lam).cpu().numpy()
        ref = RD.bagging_mixup(dict(data=x, rois=np.zeros((1, 5), np.float32)), lam)["data"]
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), lam    # bit patterns, signed zeros included
    # unaligned views take the scalar path
    flat = dev(np.concatenate([[0], x.reshape(2, -1)[0], x.reshape(2, -1)[1]]).astype(np.float32))
    n = x[0].size
    if n % 4:
        two = torch.stack([flat[1:1 + n], flat[1 + n:1 + 2 * n]])
        assert np.array_equal(ops.bagging_mixup(two, 0.25).cpu().numpy().ravel(),
                              RD.bagging_mixup(dict(data=x.reshape(2, -1), rois=np.zeros((1, 5), np.float32)), 0.25)["data"].ravel())


def test_bagging_mixup_minibatch_feeds_one_image():
    ops, roi_data = _mods()
    rng = np.random.default_rng(3)
    def entry(n, cls):
        b = rng.integers(0, 300, (n, 4)).astype(np.float32)
        b[:, 2:] += b[:, :2]
        gt = np.zeros(n, np.int32)
        gt[0] = cls
        return dict(boxes=b, obn_scores=rng.random((n, 1)).astype(np.float32), gt_classes=gt)
    host = [entry(40, 6), entry(55, 6)]
    roidb = [dict(boxes=dev(e["boxes"]), obn_scores=dev(e["obn_scores"]), gt_classes=e["gt_classes"]) for e in host]
    crops = [[0, 0, 599, 599], [10, 20, 500, 550]]
    scales = [1.2, 0.8]
    blobs = {}
    roi_data.add_wsl_blobs(blobs, scales, crops, roidb, rois_per_image=50, num_classes=21)
    data = rng.standard_normal((2, 3, 16, 24)).astype(np.float32)
    blobs["data"] = dev(data)
    lam = 0.3141592653589793
    roi_data.bagging_mixup(blobs, lam)
    ref = RD.bagging_mixup(dict(RD.add_wsl_blobs(host, scales, crops, 50, 21), data=data), lam)
    for k in ("rois", "obn_scores", "labels_oh", "labels_int32", "data"):
        assert np.array_equal(blobs[k].cpu().numpy(), ref[k]), k
    assert blobs["roi_offsets"].cpu().tolist() == [0, 90]
    assert blobs["labels_oh"].shape == (1, 20) and abs(blobs["labels_oh"].sum().item() - 1.0) < 1e-6


def test_convert_mcg_boxes_bit_exact():
    ops, _ = _mods()
    rng = np.random.default_rng(9)
    R = 3000
    y1, x1 = rng.integers(1, 400, R), rng.integers(1, 600, R)
    mat = np.stack([y1, x1, y1 + rng.integers(0, 200, R), x1 + rng.integers(0, 300, R)], 1).astype(np.float64)
    mat[5] += 0.75                                # fractional coordinates truncate
    mat[6, 0] = 0                                 # 0 wraps to 65535 like the script's uint16 arithmetic
    got = ops.convert_mcg_boxes(dev(mat)).cpu().numpy().view(np.uint16)
    assert np.array_equal(got, RD.convert_mcg_boxes(mat))
