"""GPU tests added after the round's last GPU call: verified kernels on NEW vectors (the reference-pinned golden files
head_graph.npz and test_wsl.npz, the owner-side reduce + update).  They live in one file that sorts last so that, run with
`-x`, everything that has already passed on hardware is counted before the first of them executes."""
import os

import numpy as np
import pytest
import torch

from oracle import test_time_oracle as T

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-3, torch.bfloat16: 1e-2}


def _ops():
    from nafwebsod_b200 import ops
    return ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


t = dev


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("n", [4096, 4099, 1 << 20])
@pytest.mark.parametrize("shadow", [torch.bfloat16, torch.float32])
def test_sgd_update_reduce_equals_sum_then_update(n, shadow):
    """The data-parallel owner's kernel: contributions summed in list (rank) order inside the update == the
    reference's NCCLAllreduce + ACMWeightDecayMomentumSGDUpdate on that sum (modeling/optimizer_wsl.py:52-72,
    96-137), bit for bit, operand shadow included; n = 4099 exercises the scalar tail."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(n)
    pad = lambda k: torch.randn(k + 4, device="cuda", generator=g)[:k]            # 16-byte aligned views of any length
    grads = [pad(n) for _ in range(3)]
    lr = torch.tensor([1e-3], device="cuda")
    kw = dict(momentum=0.9, gpu_num=3, lr_mult=1.0, weight_decay=5e-4)
    p0, m0 = pad(n).clone(), pad(n).clone() * 0.01
    for it in (0, 2):
        pa, ma, sa = p0.clone(), m0.clone(), torch.zeros(n, device="cuda", dtype=shadow)
        pb, mb, sb = p0.clone(), m0.clone(), torch.zeros(n, device="cuda", dtype=shadow)
        total = (grads[0] + grads[1]) + grads[2]
        ops.ACMWeightDecayMomentumSGDUpdate(total, ma, lr, pa, None, iter_count=it, p_shadow=sa, **kw)
        ops.ACMWeightDecayMomentumSGDUpdateReduce(grads, mb, lr, pb, iter_count=it, p_shadow=sb, **kw)
        assert torch.equal(pa, pb) and torch.equal(ma, mb)
        assert torch.equal(sa.float(), sb.float())


def test_head_forward_vs_reference_graph_builders(golden_dir):
    """fp32/TF32 path against tests/golden/head_graph.npz -- the blobs the reference's OWN graph builders produce when they
    are executed operator by operator (tests/golden/make_golden_head_graph.py; CPU counterpart tests/test_head_graph.py).
    Forward quantities, tolerance of the fp32/TF32 path (rel <= 1e-3)."""
    import importlib.util
    from nafwebsod_b200.heads import WeblyHeadModel
    g = np.load(os.path.join(golden_dir, "head_graph.npz"))
    spec = importlib.util.spec_from_file_location("make_golden_head_graph", os.path.join(golden_dir, "make_golden_head_graph.py"))
    maker = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(maker)
    X, rois, obn, L, params, masks, cfg = maker.load_case(g, 0)
    masks = {k: v.astype(np.uint8) for k, v in masks.items()}
    ncls, hidden = cfg["ncls"], cfg["hidden"]
    m = WeblyHeadModel(ncls, X.shape[1], 7, hidden, noise=True, entropy=True, mean_loss=True, dtype=torch.float32)
    m.load_reference_params(params)
    m.FeedBlobs(t(X), t(rois), t(obn), t(L), x_layout="NCHW")
    bl = m.RunTrainStep(dropout_masks={k: t(v) for k, v in masks.items()})
    torch.cuda.synchronize()
    tol = TOL[torch.float32]
    want = lambda k: g["case0_out_" + k]
    # the head keeps roi_feat as the fc6 GEMM operand, i.e. rounded to the nearest TF32 (2^-11 relative) and pooled-NHWC
    assert rel_l2(bl["roi_feat"].cpu().numpy().reshape(rois.shape[0], 7, 7, -1).transpose(0, 3, 1, 2), want("roi_feat")) <= tol
    assert rel_l2(bl["rois_pred"].cpu().numpy(), want("rois_pred")) <= tol
    assert rel_l2(bl["rois_pred_noise"].cpu().numpy(), want("rois_pred_noise")) <= tol
    assert rel_l2(bl["cls_prob"][0].cpu().numpy(), want("cls_prob")[0]) <= tol
    assert rel_l2(bl["cls_prob_noise"][0].cpu().numpy(), want("cls_prob_noise")[0]) <= tol
    assert rel_l2(bl["class_weight_noise"][0].cpu().numpy(), want("rois_class_weight_noise")[0]) <= 5 * tol
    assert rel_l2(bl["class_weight"][0].cpu().numpy(), want("rois_class_weight")[0]) <= 5 * tol
    assert abs(bl["loss_cls"][0].item() - float(want("loss_cls"))) <= tol * abs(float(want("loss_cls")))
    assert abs(bl["loss_cls_noise"][0].item() - float(want("loss_cls_noise"))) <= tol * abs(float(want("loss_cls_noise")))


# ---------------------------------------------------------------------------------------------- N1 + N2 vs the reference's driver
@pytest.mark.parametrize("i", [0, 1, 2])
def test_tta_and_nms_vs_reference_driver(golden_dir, i):
    """tests/golden/test_wsl.npz: the reference's own im_detect_bbox_aug + box_results_with_nms_and_limit, run unmodified
    with the shipped flickr_voc config around a deterministic pseudo head (tests/golden/make_golden_test_wsl.py; CPU
    counterpart tests/test_test_wsl_golden.py).  Here the product's device-side wrapper runs around the SAME pseudo head:
    projection, flip, dedup, gather, inverse scatter, the float32 TTA sum and mean must give the reference's averaged
    scores bit for bit, and the device NMS + limit the reference's detections (as sets per class: the reference lists
    a class's detections by descending score, the product in proposal order)."""
    import importlib.util
    from nafwebsod_b200 import test_time
    from oracle import roi_data_oracle as RD
    g = np.load(os.path.join(golden_dir, "test_wsl.npz"))
    spec = importlib.util.spec_from_file_location("make_golden_test_wsl", os.path.join(golden_dir, "make_golden_test_wsl.py"))
    maker = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(maker)
    K = int(g["num_classes"])

    class PseudoHead:
        """Stands for WeblyHeadModel: scores depend on the fed RoI rows and obn scores only."""
        def __init__(self):
            self.blobs = {}

        def FeedBlobs(self, conv5, rois, obn, labels_oh=None, roi_offsets=None, x_layout="NHWC"):
            assert roi_offsets is None                         # sync=True: exactly the unique rows are fed
            self.blobs.update(rois=rois, obn_scores=obn)

        def RunTestNet(self, want_cls_prob=True):
            s = maker.pseudo_cls_prob(self.blobs["rois"].cpu().numpy(), self.blobs["obn_scores"].cpu().numpy(), K)
            self.blobs["rois_pred"] = dev(s[:, 1:])
            return self.blobs["rois_pred"]

    h, w = (int(v) for v in g["case%d_im_shape" % i][:2])
    boxes, obn = dev(g["case%d_boxes" % i]), dev(g["case%d_obn" % i])
    model, dedup = PseudoHead(), float(g["dedup_boxes"])
    s1 = test_time.im_detect_bbox(model, None, RD.im_scale_for(h, w, int(g["test_scale"]), int(g["test_max_size"])), boxes, obn,
                                  dedup_boxes=dedup)
    assert model.blobs["rois"].shape[0] == g["case%d_single_fed_rois" % i].shape[0]
    assert np.array_equal(model.blobs["rois"].cpu().numpy(), g["case%d_single_fed_rois" % i])
    assert np.array_equal(model.blobs["obn_scores"].cpu().numpy(), g["case%d_single_fed_obn" % i].reshape(-1))
    assert np.array_equal(s1.cpu().numpy(), g["case%d_single_scores" % i])
    # the ten passes in the reference's order (core/test_wsl.py:211-256)
    passes = [(None, RD.im_scale_for(h, w, int(g["test_scale"]), int(g["test_max_size"])), w)]
    for s in g["aug_scales"]:
        sc = RD.im_scale_for(h, w, int(s), int(g["aug_max_size"]))
        passes += [(None, sc, None), (None, sc, w)]
    passes.append((None, RD.im_scale_for(h, w, int(g["test_scale"]), int(g["test_max_size"])), None))
    avg = test_time.im_detect_bbox_aug(model, passes, boxes, obn, dedup_boxes=dedup)
    assert np.array_equal(avg.cpu().numpy(), g["case%d_aug_scores" % i])
    _, _, cls_boxes = test_time.box_results_with_nms_and_limit(avg, boxes, score_thresh=float(g["score_thresh"]),
                                                               nms_thresh=float(g["nms"]), detections_per_im=int(g["detections_per_im"]))
    # Detections.  Proposals that collapse to one feature RoI get IDENTICAL scores, and the order in which the reference
    # visits tied scores is whatever NumPy's introsort leaves (cython_nms.pyx:45 `scores.argsort()[::-1]`); the device NMS
    # visits ties higher-row-first.  A tied pair of near-duplicate boxes may therefore keep the other member: per class the
    # kept SCORES must equal the reference's, and the kept rows must equal the oracle run with the device's tie order.
    counts = list(g["case%d_det_counts" % i])
    assert [int(c.shape[0]) for c in cls_boxes] == counts
    want = np.concatenate([g["case%d_det_boxes" % i], g["case%d_det_scores" % i][:, None]], axis=1)

    def nms_row_desc(dets, thresh):
        import ctypes
        if dets.shape[0] == 0:
            return np.zeros((0,), np.int64)
        dets = np.ascontiguousarray(dets, dtype=np.float32)
        n = dets.shape[0]
        order = np.lexsort((-np.arange(n), -dets[:, 4])).astype(np.int64)
        keep = np.empty(n, np.uint8)
        T._load().nawsod_oracle_nms(dets.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n,
                                    order.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_float(thresh),
                                    keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return np.where(keep != 0)[0]
    _, _, ocls, _ = T.box_results_with_nms_and_limit(g["case%d_aug_scores" % i], g["case%d_boxes" % i], K,
                                                     score_thresh=float(g["score_thresh"]), nms_thresh=float(g["nms"]),
                                                     detections_per_im=int(g["detections_per_im"]), nms_fn=nms_row_desc)
    start = 0
    canon = lambda a: a[np.lexsort(a.T[::-1])]
    for j in range(K):
        mine = cls_boxes[j].cpu().numpy()
        assert np.array_equal(np.sort(mine[:, 4]), np.sort(want[start:start + counts[j], 4])), "class %d scores" % j
        assert np.array_equal(canon(mine), canon(ocls[j])), "class %d rows" % j
        start += counts[j]


def test_update_workspace_lr_scales_the_momenta_on_the_device(golden_dir):
    """WeblyHeadModel.UpdateWorkspaceLr through the reference's learning-rate sequence (tests/golden/head_graph.npz,
    `lrseq_*`: DetectionModelHelper._SetNewLr / _CorrectMomentum run on a dictionary workspace): the `lr` blob and the
    momentum buffer after every call, bit for bit (Scale is one float32 multiply per element)."""
    from nafwebsod_b200.heads import WeblyHeadModel
    g = np.load(os.path.join(golden_dir, "head_graph.npz"))
    m = WeblyHeadModel(7, 16, 7, 64, noise=True, dtype=torch.bfloat16)
    gen = torch.Generator(device="cuda").manual_seed(1)
    m.flat_mom.normal_(0.0, 1e-3, generator=gen)
    want = m.flat_mom.clone()
    for new, blob, factor in zip(g["lrseq_new"], g["lrseq_blob"], g["lrseq_momentum_factor"]):
        got = m.UpdateWorkspaceLr(new)
        assert np.float32(got) == factor
        want = want * float(factor) if factor != 1.0 else want
        assert m.lr.item() == float(blob)
        assert torch.equal(m.flat_mom, want)
