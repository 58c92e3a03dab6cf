"""Generate tests/golden/roi_data.npz: what the reference's OWN Python returns for the training-input
contract of the head.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden_roi_data.py

detectron/roi_data/wsl.py is imported unmodified.  Its import chain pulls in caffe2, `future` and two
Cython extensions that do not exist here; none of them is touched by `_project_im_rois`, `_sample_rois`,
`add_wsl_blobs` or `prep_im_for_blob`, so they are replaced by inert stand-in modules for the import only.
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _Absent(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("caffe2", "future", "past", "pycocotools")
    EXACT = ("detectron.utils.cython_bbox", "detectron.utils.cython_nms")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS or name in self.EXACT:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__name__, m.__spec__, m.__loader__ = [], spec.name, spec, self
        return m

    def exec_module(self, module):
        pass


def import_reference():
    sys.meta_path.insert(0, _Absent())
    import future.utils
    future.utils.iteritems = lambda d: iter(d.items())
    sys.path.insert(0, "/root/reference")
    import detectron.roi_data.wsl as wsl
    import detectron.utils.blob as blob_utils
    from detectron.core.config import cfg
    return wsl, blob_utils, cfg


def synth_entry(rng, n, width, height, num_classes, n_gt=1):
    """A roidb entry shaped like json_dataset_wsl's: ground-truth rows first (gt_classes > 0), then the
    MCG proposals (uint16 boxes cast to float32, datasets/json_dataset_wsl.py:653-676)."""
    x1 = rng.integers(0, width - 17, n)
    y1 = rng.integers(0, height - 17, n)
    x2 = np.minimum(x1 + rng.integers(8, width // 2, n), width - 1)
    y2 = np.minimum(y1 + rng.integers(8, height // 2, n), height - 1)
    boxes = np.stack([x1, y1, x2, y2], 1).astype(np.uint16).astype(np.float32)
    gt = np.zeros(n, np.int32)
    gt[:n_gt] = rng.integers(1, num_classes, n_gt)
    boxes[:n_gt] = [0, 0, width - 1, height - 1]          # webly image-level "boxes" cover the image
    obn = rng.random((n, 1)).astype(np.float32)
    obn[:n_gt] = 0
    return dict(boxes=boxes, obn_scores=obn, gt_classes=gt, width=width, height=height)


def main():
    wsl, blob_utils, cfg = import_reference()
    rng = np.random.default_rng(77)
    out = {}
    num_classes = 21
    cfg.MODEL.NUM_CLASSES = num_classes
    cfg.TRAIN.BATCH_SIZE_PER_IM = 300

    # (1) _project_im_rois: boxes inside / straddling / outside the crop window, awkward scales
    entry = synth_entry(rng, 500, 500, 375, num_classes)
    cases = [(1.0, [0, 0, 499, 374]), (1.6, [0, 0, 499, 374]), (480.0 / 375.0, [37, 21, 486, 357]),
             (2000.0 / 500.0, [50, 0, 499, 336]), (0.9600000000000001, [12, 30, 460, 366]), (1200.0 / 337.0, [49, 37, 498, 373])]
    out["project_boxes"] = entry["boxes"]
    for i, (scale, crop) in enumerate(cases):
        r = wsl._project_im_rois(entry["boxes"].copy(), scale, np.array(crop, dtype=np.int32))
        out["project_scale_%d" % i] = np.float64(scale)
        out["project_crop_%d" % i] = np.array(crop, dtype=np.int32)
        out["project_out_%d" % i] = np.asarray(r)                     # float64, as returned
        out["project_out32_%d" % i] = np.asarray(r).astype(np.float32)
    out["project_cases"] = np.int32(len(cases))

    # (2) _sample_rois / add_wsl_blobs on a two-image minibatch (image 1 has fewer boxes than BATCH_SIZE_PER_IM
    #     and two ground-truth rows)
    roidb = [synth_entry(rng, 450, 500, 375, num_classes), synth_entry(rng, 180, 333, 500, num_classes, n_gt=2)]
    im_scales = [576.0 / 375.0, 688.0 / 333.0]
    im_crops = np.array([[18, 25, 355, 474], [0, 0, 499, 332]], dtype=np.int32)[:, (1, 0, 3, 2)]     # minibatch_wsl.py:63-64
    for i, e in enumerate(roidb):
        for k in ("boxes", "obn_scores", "gt_classes"):
            out["mb_%s_%d" % (k, i)] = e[k]
    out["mb_im_scales"] = np.array(im_scales, np.float64)
    out["mb_im_crops"] = im_crops
    out["mb_rois_per_image"] = np.int32(cfg.TRAIN.BATCH_SIZE_PER_IM)
    out["mb_num_classes"] = np.int32(num_classes)
    blobs = {k: [] for k in wsl.get_wsl_blob_names(is_training=True)}
    wsl.add_wsl_blobs(blobs, im_scales, im_crops, [dict(e) for e in roidb])
    for k, v in blobs.items():
        out["mb_out_" + k] = np.asarray(v)

    # (3) prep_im_for_blob's scale rule on a few image sizes (the resize itself is outside the path)
    sizes = [(375, 500), (500, 333), (300, 1000), (1200, 1600), (97, 1204)]
    scales = []
    for (h, w) in sizes:
        for target, max_size in ((480, 2000), (688, 2000), (1200, 2000), (600, 1000)):
            im = np.zeros((h, w, 3), np.uint8)
            _, s = blob_utils.prep_im_for_blob(im, np.zeros((1, 1, 3)), target, max_size)
            scales.append([h, w, target, max_size, s])
    out["scale_cases"] = np.array(scales, np.float64)

    np.savez_compressed(os.path.join(HERE, "roi_data.npz"), **out)
    print("wrote roi_data.npz:", {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith("mb_out")})


if __name__ == "__main__":
    main()
