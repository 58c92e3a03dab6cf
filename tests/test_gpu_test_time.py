"""GPU parity of the test-time wrapper around the head, through the C ABI,
against oracle/test_time_oracle.py (its NMS is pinned to the reference's own Cython NMS) and the golden
vectors of tests/golden/nms_ref.npz.  Integer / index / mask results are bit-exact; the float32 score
arithmetic (projection, TTA sum, mean) is bit-exact; only paths through the head GEMMs carry the
north-star tolerance (rel <= 1e-3 fp32/TF32)."""
import os

import numpy as np
import pytest
import torch

from oracle import nawsod_oracle as O
from oracle import test_time_oracle as T

pytestmark = pytest.mark.gpu


def _ops():
    from nafwebsod_b200 import ops
    return ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _mcg_boxes(R, img_h, img_w, seed, dup_frac=0.25):
    """Integer MCG-like proposals in original image coordinates, with near-duplicates (boxes that differ
    by a few pixels and collapse to one feature RoI after the 1/16 rounding of the dedup hash)."""
    rng = np.random.default_rng(seed)
    b = O.synth_rois(R, img_h, img_w, 0, seed=seed)[:, 1:].copy()
    nd = int(R * dup_frac)
    src = rng.integers(0, R - nd, nd)
    b[R - nd:] = np.clip(b[src] + rng.integers(-2, 3, (nd, 4)), 0, [img_w - 1, img_h - 1, img_w - 1, img_h - 1])
    b[:, 2] = np.maximum(b[:, 2], b[:, 0])
    b[:, 3] = np.maximum(b[:, 3], b[:, 1])
    return b.astype(np.float32)


# ---------------------------------------------------------------------------------------------- N1
# 9999 = TEST.PROPOSAL_LIMIT of the shipped configs (configs/flickr_voc/na_wsddn_V-16-C5_1x.yaml:39); 16384 = the kernel's capacity
@pytest.mark.parametrize("R", [1, 7, 500, 2000, 4000, 8000, 9999, 16384])
def test_project_and_dedup_match_numpy(R):
    ops = _ops()
    boxes = _mcg_boxes(R, 375, 500, seed=R) if R > 8 else _mcg_boxes(16, 375, 500, seed=R)[:R]
    obn = np.random.default_rng(1).random(R).astype(np.float32)
    for im_scale, flip in ((1.376, None), (688.0 / 375.0, 500), (2.4, None)):
        want = T.get_rois_blob(T.flip_boxes(boxes, flip) if flip else boxes, im_scale)
        rois, obn1 = ops.project_rois(dev(boxes), im_scale, flip_width=flip, obn_scores=dev(obn))
        assert np.array_equal(rois.cpu().numpy(), want)
        assert np.array_equal(obn1.cpu().numpy(), obn + np.float32(1.0))
        index, inv, nu, offs = ops.dedup_rois(rois, 1.0 / 16)
        widx, winv = T.dedup_rois(want, 1.0 / 16)
        n = int(nu.item())
        assert n == len(widx) and offs.cpu().tolist() == [0, n]
        assert np.array_equal(index.cpu().numpy()[:n], widx)
        assert np.array_equal(inv.cpu().numpy(), winv)
        assert np.all(index.cpu().numpy()[n:] == widx[0])
        if R >= 500:
            assert n < R                                           # the near-duplicates did collapse
        got = ops.gather_rows(rois, index, n).cpu().numpy()
        assert np.array_equal(got, want[widx])


def test_dedup_errors():
    ops = _ops()
    with pytest.raises(RuntimeError, match="R=16385"):
        ops.dedup_rois(torch.zeros((16385, 5), device="cuda"))
    with pytest.raises(RuntimeError):
        ops.dedup_rois(torch.zeros((10, 4), device="cuda"))


def test_scatter_accumulate_finalize_bit_exact():
    ops = _ops()
    rng = np.random.default_rng(4)
    R, C = 900, 20
    passes, acc = [], None
    for t in range(10):
        nu = int(rng.integers(300, R))
        inv = rng.integers(0, nu, R).astype(np.int32)
        rp = rng.random((nu, C)).astype(np.float32) * np.float32(1e-2)
        passes.append(T.test_cls_prob(rp)[inv])
        # a column slice of a wider matrix exercises the row pitch
        wide = torch.zeros((nu, C + 12), device="cuda")
        wide[:, 4:4 + C] = dev(rp)
        acc = ops.scatter_scores(wide[:, 4:4 + C], dev(inv), out=acc, accumulate=acc is not None)
        if t == 0:
            assert np.array_equal(acc.cpu().numpy(), passes[0])
    got = ops.scores_finalize(acc, len(passes)).cpu().numpy()
    assert np.array_equal(got, T.tta_average(passes))
    rp = rng.random((50, C)).astype(np.float32)
    assert np.array_equal(ops.scatter_scores(dev(rp)).cpu().numpy(), T.test_cls_prob(rp))    # no dedup: identity map


# ---------------------------------------------------------------------------------------------- N2
def _unique_scores(rng, R, K1):
    """Distinct scores inside every class (no ties): a permutation of (k+1)/(R+1) per column."""
    s = np.stack([(rng.permutation(R) + 1) / np.float32(R + 1) for _ in range(K1)], axis=1).astype(np.float32)
    return (s * np.float32(0.2)).astype(np.float32)


def _check_mask(keep, num_keep, want_mask):
    k = keep.cpu().numpy()
    assert np.array_equal(k, want_mask)
    assert np.array_equal(num_keep.cpu().numpy(), want_mask.sum(axis=1))


def test_nms_reference_golden(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "nms_ref.npz"))
    for k in range(int(g["n_cases"])):
        dets, th, want = g["dets%d" % k], float(g["thresh%d" % k]), g["keep%d" % k]
        n = dets.shape[0]
        scores = np.zeros((n, 2), np.float32)
        scores[:, 1] = dets[:, 4]
        keep, num_keep, _ = ops.nms_and_limit(dev(scores), dev(dets[:, :4]), score_thresh=-1.0, nms_thresh=th, detections_per_im=0)
        got = np.where(keep.cpu().numpy()[1] != 0)[0]
        if int(g["ties%d" % k]) and n > 16:
            # equal scores: NumPy's order is unspecified, ours is "higher row first" -> check the NMS invariants
            assert T.nms_invariants_hold(dets, got, np.float32(th))
        else:
            assert np.array_equal(got, want), k
        assert int(num_keep.cpu()[1]) == len(got) and int(num_keep.cpu()[0]) == 0


@pytest.mark.parametrize("cfg", [(2000, 21, 1e-9, 0.5, 100), (4000, 81, 1e-9, 0.5, 100), (2000, 21, 0.05, 0.3, 0),
                                 (8000, 21, 0.01, 0.4, 100), (300, 6, 0.05, 0.5, 10000), (37, 3, 0.0, 0.5, 5)])
def test_nms_and_limit_vs_oracle(cfg):
    ops = _ops()
    from nafwebsod_b200 import test_time
    R, K1, st, nt, dpi = cfg
    rng = np.random.default_rng(R + K1)
    scores = _unique_scores(rng, R, K1)
    boxes = _mcg_boxes(R, 375, 500, seed=R + 1) if R > 8 else _mcg_boxes(16, 375, 500, seed=3)[:R]
    boxes = (boxes * np.float32(1.0)).astype(np.float32)
    ws, wb, wcls, wmask = T.box_results_with_nms_and_limit(scores, boxes, K1, st, nt, dpi)
    keep, num_keep, thr = ops.nms_and_limit(dev(scores), dev(boxes), score_thresh=st, nms_thresh=nt, detections_per_im=dpi)
    _check_mask(keep, num_keep, wmask)
    _, _, _, mask0 = T.box_results_with_nms_and_limit(scores, boxes, K1, st, nt, 0)
    if dpi > 0 and mask0.sum() > dpi:                                 # the limit bites: image_thresh = dpi-th largest score
        assert float(thr.item()) == float(ws.min()) and wmask.sum() >= dpi
    else:
        assert float(thr.item()) == -float(np.finfo(np.float32).max) and np.array_equal(wmask, mask0)
    s, b, cls_boxes = test_time.box_results_with_nms_and_limit(dev(scores), dev(boxes), score_thresh=st, nms_thresh=nt,
                                                               detections_per_im=dpi)
    assert np.array_equal(s.cpu().numpy(), ws) and np.array_equal(b.cpu().numpy(), wb)
    assert len(cls_boxes) == K1 and cls_boxes[0].shape[0] == 0
    for j in range(1, K1):
        assert np.array_equal(cls_boxes[j].cpu().numpy(), wcls[j]), j


def test_nms_tied_scores_are_deterministic_and_valid():
    """Scores of de-duplicated proposals tie exactly; the visiting order of ties is 'higher row first'."""
    ops = _ops()
    _valid_nms = T.nms_invariants_hold
    rng = np.random.default_rng(8)
    R = 400
    boxes = _mcg_boxes(R, 375, 500, seed=5)
    scores = np.zeros((R, 2), np.float32)
    scores[:, 1] = (np.round(rng.random(R) * 16) / 16 + 0.01).astype(np.float32)
    k1, _, _ = ops.nms_and_limit(dev(scores), dev(boxes), score_thresh=0.0, nms_thresh=0.5, detections_per_im=0)
    k2, _, _ = ops.nms_and_limit(dev(scores), dev(boxes), score_thresh=0.0, nms_thresh=0.5, detections_per_im=0)
    assert torch.equal(k1, k2)
    dets = np.hstack([boxes, scores[:, 1:2]]).astype(np.float32)
    assert _valid_nms(dets, np.where(k1.cpu().numpy()[1] != 0)[0], np.float32(0.5))
    # the same order handed to the C oracle reproduces the mask bit for bit
    order = np.lexsort((-np.arange(R), -dets[:, 4])).astype(np.int64)       # score desc, then row desc
    import ctypes
    keep = np.empty(R, np.uint8)
    T._load().nawsod_oracle_nms(dets.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), R,
                                order.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_float(0.5),
                                keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    assert np.array_equal(k1.cpu().numpy()[1], keep)


# ---------------------------------------------------------------------------------------------- N4
@pytest.mark.parametrize("shape", [(12, 6), (2000, 20), (4000, 80)])
def test_min_entropy_loss(shape):
    ops = _ops()
    N, C = shape
    rng = np.random.default_rng(N)
    X = rng.random((N, C)).astype(np.float32) ** 4
    X[0, 0] = 0.0                                                  # exercises the 1e-20 floor and the 1e4 clamp
    L = (rng.random((1, C)) < 0.3).astype(np.float32)
    L[0, 0] = 1
    y = ops.MinEntropyLoss(dev(X), dev(L)).item()
    want, _ = T.min_entropy_loss(X, L)
    assert abs(y - float(want)) <= 1e-5 * abs(float(want))
    d = ops.MinEntropyLossGradient(dev(X), dev(L), torch.tensor([0.1], device="cuda")).cpu().numpy()
    wd = T.min_entropy_loss_grad(X, L, np.float32(0.1))
    scale = np.float32(0.1) / T.min_entropy_loss(X, L)[1]
    np.testing.assert_allclose(d / scale, wd / scale, rtol=1e-5, atol=1e-5)
    assert np.array_equal(d == 0, wd == 0)
    with pytest.raises(RuntimeError, match="one row"):
        ops.MinEntropyLoss(dev(X), dev(np.repeat(L, 2, axis=0)))


# ---------------------------------------------------------------------------------------------- N1 end to end
def _small_model(ncls=6, Cc=16, Hd=64, seed=3):
    from nafwebsod_b200.heads import WeblyHeadModel
    params = O.synth_params(ncls - 1, Cc * 49, Hd, noise=True, seed=seed)
    for k in params:
        if k.endswith("fc6_w"):
            params[k] = (params[k] * np.float32(np.sqrt(25088.0 / (Cc * 49)))).astype(np.float32)
        if k.endswith("fc7_w"):
            params[k] = (params[k] * np.float32(np.sqrt(4096.0 / Hd))).astype(np.float32)
    m = WeblyHeadModel(ncls, Cc, 7, Hd, dtype=torch.float32, train=False)
    m.load_reference_params(params)
    return m, params


def _oracle_detect(X, boxes, obn, im_scale, flip, params, ncls):
    b = T.flip_boxes(boxes, flip) if flip else boxes
    rois = T.get_rois_blob(b, im_scale)
    index, inv = T.dedup_rois(rois)
    ref = O.head_forward_backward(X, rois[index], (obn + np.float32(1.0))[index].reshape(-1, 1), np.zeros((1, ncls - 1), np.float32),
                                  params, noise=False, entropy=False, backward=False)
    return T.test_cls_prob(ref["rois_pred"])[inv]


def test_im_detect_bbox_and_aug_vs_oracle():
    from nafwebsod_b200 import test_time
    ncls, Cc = 6, 16
    m, params = _small_model(ncls, Cc)
    R, img_h, img_w = 300, 192, 256
    boxes = _mcg_boxes(R, img_h, img_w, seed=11)
    obn = np.random.default_rng(12).random(R).astype(np.float32)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
    # one pass, both host-sync modes
    X = O.synth_conv5(1, Cc, 17, 22, seed=13)                    # map of the image at scale 1.376 (stride 16)
    want = _oracle_detect(X, boxes, obn, 1.376, None, params, ncls)
    got = test_time.im_detect_bbox(m, dev(X), 1.376, dev(boxes), dev(obn), x_layout="NCHW").cpu().numpy()
    assert got.shape == (R, ncls) and rel(got, want) <= 1e-3
    assert np.array_equal(got[:, 0], got[:, 1])
    nosync = test_time.im_detect_bbox(m, dev(X), 1.376, dev(boxes), dev(obn), x_layout="NCHW", sync=False).cpu().numpy()
    assert rel(nosync, got) <= 1e-5                             # same unique set, same softmax support
    # duplicates share their score exactly (they ARE the same feature RoI)
    _, inv = T.dedup_rois(T.get_rois_blob(boxes, 1.376))
    for u in np.unique(inv)[:50]:
        rows = np.where(inv == u)[0]
        assert np.all(got[rows] == got[rows[0]])
    # TTA: flipped pass, a second scale and its flip, identity last; AVG over the four
    passes, wants = [], []
    for k, (scale, hw, flip) in enumerate([(1.376, (17, 22), img_w), (2.0, (24, 32), None), (2.0, (24, 32), img_w), (1.376, (17, 22), None)]):
        Xk = O.synth_conv5(1, Cc, hw[0], hw[1], seed=20 + k)
        passes.append((dev(Xk), scale, flip))
        wants.append(_oracle_detect(Xk, boxes, obn, scale, flip, params, ncls))
    avg = test_time.im_detect_bbox_aug(m, passes, dev(boxes), dev(obn), x_layout="NCHW").cpu().numpy()
    assert rel(avg, T.tta_average(wants)) <= 1e-3
