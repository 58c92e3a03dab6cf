"""Generate tests/golden/nms_ref.npz from the REFERENCE's own Cython NMS
(/root/reference/detectron/utils/cython_nms.pyx compiled unmodified by oracle/build_ref_nms.sh).
Run in the build container only (the GPU box has no /root/reference); the vectors are committed.

    bash oracle/build_ref_nms.sh && python tests/golden/make_golden_nms.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import test_time_oracle as T  # noqa: E402


def make_dets(rng, n, frac, ties):
    x1 = rng.integers(0, 600, n).astype(np.float32)
    y1 = rng.integers(0, 400, n).astype(np.float32)
    w = rng.integers(1, 300, n).astype(np.float32)
    h = rng.integers(1, 300, n).astype(np.float32)
    if frac:                                   # projected / scaled boxes are not integers
        x1 += rng.random(n).astype(np.float32)
        y1 += rng.random(n).astype(np.float32)
        w *= np.float32(1.375)
    s = rng.random(n).astype(np.float32)
    if ties:
        s = (np.round(s * 16) / 16).astype(np.float32)
    return np.stack([x1, y1, x1 + w, y1 + h, s], axis=1).astype(np.float32)


def main():
    ref = T.reference_nms()
    assert ref is not None, "build oracle/_ref first (oracle/build_ref_nms.sh)"
    rng = np.random.default_rng(1234)
    out = {}
    cases = [(1, 0, 0, 0.3), (2, 0, 0, 0.5), (17, 0, 0, 0.3), (64, 1, 0, 0.5), (300, 0, 0, 0.3), (300, 1, 0, 0.5),
             (300, 0, 0, 0.7), (1000, 1, 0, 0.5), (2000, 0, 0, 0.4), (12, 0, 1, 0.5), (400, 1, 1, 0.3)]
    for k, (n, frac, ties, th) in enumerate(cases):
        d = make_dets(rng, n, frac, ties)
        keep = np.asarray(ref(d, np.float32(th)), dtype=np.int64)
        out["dets%d" % k], out["thresh%d" % k], out["keep%d" % k], out["ties%d" % k] = d, np.float32(th), keep, np.int32(ties)
    out["n_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nms_ref.npz"), **out)
    print("wrote nms_ref.npz:", len(cases), "cases")


if __name__ == "__main__":
    main()
