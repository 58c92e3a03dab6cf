"""Generate tests/golden/ref_kernels.npz: outputs of the reference's OWN CUDA kernels (RoIIoU, the in-tree RoI
max-pooling clone RoILoopPool, MinEntropyLoss), executed on the host by oracle/_ref/libnawsod_ref_kernels.so
(oracle/build_ref_kernels.py cuts the kernels out of /root/reference/detectron/ops/*.cu unmodified).
Run in the BUILD container (needs /root/reference):

    python oracle/build_ref_kernels.py && python tests/golden/make_golden_ref_kernels.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import nawsod_oracle as O          # noqa: E402  (input generators only)
from oracle import ref_kernels as RK           # noqa: E402
from test_gpu_ops import _mixed_rois           # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    out = {}
    # RoIIoU: MCG-like boxes scaled to fractional network-input coordinates (the op truncates them to int)
    rois = O.synth_rois(128, 608, 800, seed=3)
    rois[:, 1:] *= np.float32(1.37)
    rois[5, 1:] = rois[6, 1:]                                      # identical boxes: IoU 1 off the diagonal
    rois[7, 1:] = (10.2, 10.9, 9.1, 9.5)                           # degenerate box
    out["iou_rois"], out["iou_J"] = rois, RK.roi_iou(rois)
    # RoILoopPool with the inner rectangle disabled == the shared arithmetic of RoIPoolF
    for tag, (N, C, H, W, scale, stride) in (("p16", (2, 4, 38, 50, 1 / 16, 16)), ("p8", (1, 4, 60, 80, 1 / 8, 8))):
        X = (rng.random((N, C, H, W), dtype=np.float32) * np.float32(0.9) + np.float32(0.1)).astype(np.float32)   # > 0
        r = np.concatenate([_mixed_rois(128 // N, H * stride, W * stride, b, seed=40 + b) for b in range(N)])
        Y, A = RK.roi_loop_pool(X, RK.rois9(r), scale)
        dY = rng.standard_normal(Y.shape).astype(np.float32)
        dX = RK.roi_loop_pool_grad(X.shape, RK.rois9(r), A, dY, scale)
        out.update({tag + "_X": X, tag + "_rois": r, tag + "_scale": np.float32(scale), tag + "_Y": Y, tag + "_A": A,
                    tag + "_dY": dY, tag + "_dX": dX})
    # the clone's third delta on a post-ReLU-like map (zeros): maxval starts at 0, so an all-zero bin keeps argmax -1
    Xz = O.synth_conv5(1, 4, 38, 50, seed=5)
    rz = O.synth_rois(64, 608, 800, seed=6)
    Yz, Az = RK.roi_loop_pool(Xz, RK.rois9(rz), 1 / 16)
    out.update(z_X=Xz, z_rois=rz, z_Y=Yz, z_A=Az)
    # MinEntropyLoss kernels
    P = rng.random((300, 20)).astype(np.float32)
    P /= P.sum(1, keepdims=True)
    P[0, 3] = 0                                                     # hits the 1e-20 clamp
    L = np.zeros((1, 20), np.float32)
    L[0, [3, 7, 11]] = 1
    s, cnt = RK.min_entropy_forward_kernel(P, L)
    d = RK.min_entropy_backward_kernel(P, L, np.float32(0.37) / (np.float32(1) + cnt))
    out.update(me_X=P, me_L=L, me_sum=np.float32(s), me_count=np.float32(cnt), me_dX=d, me_dY=np.float32(0.37))
    np.savez_compressed(os.path.join(HERE, "ref_kernels.npz"), **out)
    print("wrote ref_kernels.npz:", {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    main()
