"""Generate tests/golden/test_wsl.npz: what the reference's OWN test driver returns around the head's forward pass
(SURVEY.md section 8f rows N1 / N2).  Run in the BUILD container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden_test_wsl.py

detectron/core/test_wsl.py is imported unmodified (its import chain needs caffe2, pycocotools, `future` and two Cython
extensions: inert stand-ins for the import, except cython_nms, which is the reference's own .pyx compiled by
oracle/build_ref_nms.sh).  `im_detect_bbox`, `im_detect_bbox_hflip`, `im_detect_bbox_scale`, `im_detect_bbox_aug` and
`box_results_with_nms_and_limit` then run as they are, with the shipped flickr_voc config plus TEST.BBOX_AUG.ENABLED, against
a stand-in Caffe2 workspace: FeedBlob stores the blobs, RunNet evaluates a deterministic pseudo head
(`pseudo_cls_prob`: class scores that depend only on the fed RoI row and its obn score, so that duplicate feature RoIs get
identical rows -- the property the reference's dedup relies on), FetchBlob returns `cls_prob`.  Everything AROUND the net --
RoI projection to the input scale, the float64 hash dedup and its inverse map, flipping, the per-scale passes, score
averaging, thresholding, NMS, the detections-per-image limit -- is therefore the reference's code, not a re-reading of it.
Two environment shims only: `np.float` / `np.int` (removed from NumPy >= 1.24, used by test_wsl.py:1027-1028) and the PyYAML
`Loader` argument.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def pseudo_cls_prob(rois, obn, num_classes):
    """[R, num_classes] float32 scores in (0, 1): a smooth deterministic function of one fed RoI row (float32 [5]) and its
    obn score.  Column 0 duplicates column 1 like the head's test-mode cls_prob (wsl_heads.py:57-67)."""
    r = np.asarray(rois, np.float64)
    o = np.asarray(obn, np.float64).reshape(-1, 1)
    c = np.arange(1, num_classes, dtype=np.float64)[None, :]
    phase = (r[:, 1:2] * 0.0131 + r[:, 2:3] * 0.0173 + r[:, 3:4] * 0.0071 + r[:, 4:5] * 0.0113) * (1.0 + 0.37 * c) + o * 2.1 + c
    s = (0.5 + 0.5 * np.sin(phase)) ** 6                       # mostly small, a few confident boxes per class
    return np.concatenate([s[:, :1], s], axis=1).astype(np.float32)


class FakeWorkspace:
    def __init__(self, num_classes):
        self.blobs, self.num_classes, self.fed = {}, num_classes, []

    def FeedBlob(self, name, value):
        self.blobs[str(name)] = np.array(value)

    def RunNet(self, name):
        rois, obn = self.blobs["rois"], self.blobs["obn_scores"]
        assert rois.dtype == np.float32 and rois.shape[1] == 5
        self.fed.append((rois.copy(), np.asarray(obn).copy(), self.blobs["data"].shape))
        self.blobs["cls_prob"] = pseudo_cls_prob(rois, obn, self.num_classes)

    def FetchBlob(self, name):
        return self.blobs[str(name)]


def import_reference():
    spec = importlib.util.spec_from_file_location("make_golden_roi_data", os.path.join(HERE, "make_golden_roi_data.py"))
    maker = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(maker)
    absent = maker._Absent()
    absent.EXACT = ("detectron.utils.cython_bbox",)                                     # cython_nms is the real one
    sys.meta_path.insert(0, absent)
    import future.utils
    future.utils.iteritems = lambda d: iter(d.items())
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import cython_nms                                                                     # oracle/build_ref_nms.sh
    sys.modules["detectron.utils.cython_nms"] = cython_nms
    if not hasattr(np, "float"):
        np.float, np.int = float, int                                                     # NumPy < 1.24 aliases
    import yaml
    import detectron.utils.env as envu
    envu.yaml_load = lambda f: yaml.load(f, Loader=yaml.SafeLoader)
    from detectron.core.config import cfg, merge_cfg_from_file
    import detectron.core.test_wsl as test_wsl
    merge_cfg_from_file("/root/reference/configs/flickr_voc/na_wsddn_V-16-C5_1x.yaml")
    return test_wsl, cfg


class _Net:
    class _P:
        name = "wsl_test_net"

    def Proto(self):
        return self._P()


class _Model:
    net = _Net()


def synth_image_and_boxes(rng, h, w, n):
    im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    x1 = rng.integers(0, w - 17, n)
    y1 = rng.integers(0, h - 17, n)
    x2 = np.minimum(x1 + rng.integers(8, w // 2, n), w - 1)
    y2 = np.minimum(y1 + rng.integers(8, h // 2, n), h - 1)
    boxes = np.stack([x1, y1, x2, y2], 1).astype(np.uint16).astype(np.float32)            # MCG boxes are uint16 (convert_mcg.py:46-49)
    k = n // 8
    boxes[n // 2: n // 2 + k] = boxes[:k]                                                 # exact duplicates
    boxes[n - k:] = boxes[k: 2 * k] + np.float32([1, 1, 0, 0])                            # near-duplicates: collide only after /8 rounding
    obn = rng.random((n, 1)).astype(np.float32)
    obn[n // 2: n // 2 + k] = obn[:k]
    return im, boxes, obn


def main():
    test_wsl, cfg = import_reference()
    num_classes = cfg.MODEL.NUM_CLASSES
    ws = FakeWorkspace(num_classes)
    test_wsl.workspace.FeedBlob, test_wsl.workspace.RunNet, test_wsl.workspace.FetchBlob = ws.FeedBlob, ws.RunNet, ws.FetchBlob
    test_wsl.core.ScopedName = lambda n: n
    rng = np.random.default_rng(4242)
    out = {"num_classes": np.int32(num_classes), "dedup_boxes": np.float64(cfg.DEDUP_BOXES), "test_scale": np.int32(cfg.TEST.SCALE),
           "test_max_size": np.int32(cfg.TEST.MAX_SIZE), "aug_scales": np.array(cfg.TEST.BBOX_AUG.SCALES, np.int32),
           "aug_max_size": np.int32(cfg.TEST.BBOX_AUG.MAX_SIZE), "score_thresh": np.float64(cfg.TEST.SCORE_THRESH),
           "nms": np.float64(cfg.TEST.NMS), "detections_per_im": np.int32(cfg.TEST.DETECTIONS_PER_IM)}
    assert cfg.TEST.BBOX_AUG.SCORE_HEUR == "AVG" and cfg.TEST.BBOX_AUG.COORD_HEUR == "ID" and cfg.TEST.BBOX_AUG.H_FLIP
    assert cfg.TEST.BBOX_AUG.SCALE_H_FLIP and not cfg.TEST.BBOX_REG and not cfg.MODEL.FASTER_RCNN
    cases = [(375, 500, 600), (333, 500, 257), (480, 360, 1)]
    for i, (h, w, n) in enumerate(cases):
        im, boxes, obn = synth_image_and_boxes(rng, h, w, max(n, 8))
        boxes, obn = boxes[:n], obn[:n]
        pre = "case%d_" % i
        out[pre + "im_shape"], out[pre + "boxes"], out[pre + "obn"] = np.array(im.shape, np.int32), boxes, obn
        # (1) one pass at TEST.SCALE
        ws.fed.clear()
        scores, pred_boxes, im_scale = test_wsl.im_detect_bbox(_Model(), im, cfg.TEST.SCALE, cfg.TEST.MAX_SIZE, boxes=boxes, obn_scores=obn)
        out[pre + "single_scores"], out[pre + "single_boxes"], out[pre + "single_im_scale"] = scores, pred_boxes, np.float64(im_scale)
        out[pre + "single_fed_rois"], out[pre + "single_fed_obn"] = ws.fed[0][0], ws.fed[0][1]
        # (2) test-time augmentation: hflip at TEST.SCALE, every BBOX_AUG scale with and without flip, then the identity pass
        ws.fed.clear()
        s_c, b_c, im_scale_i = test_wsl.im_detect_bbox_aug(_Model(), im, box_proposals=boxes, obn_scores=obn)
        out[pre + "aug_scores"], out[pre + "aug_boxes"] = s_c, b_c
        out[pre + "aug_passes"] = np.int32(len(ws.fed))
        out[pre + "aug_fed_counts"] = np.array([f[0].shape[0] for f in ws.fed], np.int32)
        out[pre + "aug_data_shapes"] = np.array([f[2] for f in ws.fed], np.int32)
        # (3) thresholding + NMS + detections-per-image limit on the averaged scores
        sc, bx, cls_boxes = test_wsl.box_results_with_nms_and_limit(s_c, b_c)
        out[pre + "det_scores"], out[pre + "det_boxes"] = sc, bx
        out[pre + "det_counts"] = np.array([len(c) for c in cls_boxes], np.int32)
        print("case", i, (h, w, n), "single-pass unique rois", ws.fed and out[pre + "single_fed_rois"].shape[0], "aug passes",
              int(out[pre + "aug_passes"]), "detections", sc.shape[0])
    out["cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "test_wsl.npz"), **out)
    print("wrote test_wsl.npz")


if __name__ == "__main__":
    main()
