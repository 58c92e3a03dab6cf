"""Generate the committed golden fixtures for the hot path.  Run in the BUILD container
(needs /root/reference for oracle/_ref and torchvision for the RoIPool cross-check):

    python tests/golden/make_golden.py

Sources of truth (none of them is the oracle under test):
  roi_pool.npz    torch.ops.torchvision.roi_pool (CPU) -- same Caffe2 lineage as RoIPoolF
  ref_ops.npz     the reference's own CPU operators compiled unmodified from
                  /root/reference/detectron/ops/*.cc (oracle/build_ref.sh -> oracle/_ref)
  head_small.npz  torch autograd in float64 over an independent re-expression of the
                  MIL/loss graph (modeling/wsl_heads.py:23-56, webly_heads.py:32-74,123-197)
"""
import os
import subprocess
import sys

import numpy as np
import torch
import torchvision  # noqa: F401  (registers torch.ops.torchvision.roi_pool)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def gen_roi_pool():
    rng = np.random.default_rng(100)
    N, C, H, W = 2, 6, 13, 17
    X = (rng.random((N, C, H, W), dtype=np.float32) * (rng.random((N, C, H, W)) < 0.6)).astype(np.float32)
    X[1, 2] = 0.0                                    # an all-zero plane: ties -> first cell
    X[0, 3] = -X[0, 3] - 1.0                         # a strictly negative plane
    rois = []
    for _ in range(40):
        b = rng.integers(N)
        x1, y1 = rng.integers(0, W * 16 - 20), rng.integers(0, H * 16 - 20)
        w, h = rng.integers(1, W * 8), rng.integers(1, H * 8)
        rois.append([b, x1, y1, min(x1 + w, W * 16 - 1), min(y1 + h, H * 16 - 1)])
    rois += [
        [0, 0, 0, W * 16 - 1, H * 16 - 1],           # whole image
        [1, 8, 8, 8, 8],                             # single point -> 1x1 roi, 6/7 bins share a cell
        [0, 24, 40, 24, 40],                         # x.5 after scaling: roundf half away from zero
        [1, 300, 250, 500, 400],                     # partly outside the map -> clipped / empty bins
        [0, 1000, 1000, 1200, 1100],                 # fully outside -> all bins empty (0, -1)
        [1, 100, 100, 50, 60],                       # malformed (x2<x1): forced to 1x1
        [0, -40, -30, 60, 50],                       # negative coordinates
        [0, 7.3, 9.9, 130.2, 88.8],                  # fractional coords (im_scale != 1)
    ]
    rois = np.asarray(rois, dtype=np.float32)
    scale = 1.0 / 16.0
    Y, A = torch.ops.torchvision.roi_pool(torch.from_numpy(X), torch.from_numpy(rois), scale, 7, 7)
    # second scale (1/8: shipped WSL.DILATION == 2) on the same inputs
    Y8, A8 = torch.ops.torchvision.roi_pool(torch.from_numpy(X), torch.from_numpy(rois * 0.5), 1.0 / 8.0, 7, 7)
    np.savez_compressed(os.path.join(HERE, "roi_pool.npz"), X=X, rois=rois, scale=np.float32(scale),
                        Y=Y.numpy(), argmax=A.numpy().astype(np.int32),
                        rois8=(rois * 0.5).astype(np.float32), Y8=Y8.numpy(), argmax8=A8.numpy().astype(np.int32))


def gen_ref_ops():
    from oracle import ref_ops as RO
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    rng = np.random.default_rng(200)
    out = {}
    X = rng.random((9, 4, 7, 7)).astype(np.float32)
    S = (rng.random((9, 1)) + 1).astype(np.float32)
    out["boost_X"], out["boost_S"] = X, S
    out["boost_Y"] = RO.RefOp("RoIFeatureBoost").run([X, S], 1)[0]
    out["boost_dX"] = RO.RefOp("RoIFeatureBoostGradient").run([X, S], 1)[0]
    k = 0
    for C in (20, 80):
        for is_mean in (0, 1):
            for variant in range(3):
                x = (rng.random((1, C)) ** 3).astype(np.float32)
                lab = np.zeros((1, C), np.float32)
                lab[0, rng.integers(C)] = 1
                if variant == 1:                       # saturations hit the 1e-20 / 1e4 clamps
                    x[0, 0], x[0, 1], x[0, 2] = 0.0, 1.0, 1e-30
                if variant == 2:                       # mixup soft labels (loader_wsl.py:149-168)
                    lam = np.float32(rng.beta(1.5, 1.5))
                    lab = lab * lam
                    lab[0, rng.integers(C)] += np.float32(1) - lam
                w = rng.random((1, C)).astype(np.float32)
                one = np.ones(1, np.float32)
                pre = "ce%d_" % k
                out[pre + "x"], out[pre + "l"], out[pre + "w"] = x, lab, w
                out[pre + "is_mean"] = np.int32(is_mean)
                out[pre + "loss_w"] = RO.RefOp("WeightedCrossEntropyWithLogits", is_mean=is_mean).run([x, lab, w], 1)[0]
                out[pre + "loss_u"] = RO.RefOp("CrossEntropyWithLogits", is_mean=is_mean).run([x, lab], 1)[0]
                out[pre + "grad_w"] = RO.RefOp("WeightedCrossEntropyWithLogitsGradient", is_mean=is_mean).run([x, lab, w, one], 1)[0]
                out[pre + "grad_u"] = RO.RefOp("CrossEntropyWithLogitsGradient", is_mean=is_mean).run([x, lab, one], 1)[0]
                k += 1
    out["ce_count"] = np.int32(k)
    cfgs = [(1, 1, 5e-4, 1.0), (2, 4, 5e-4, 1.0), (1, 8, 0.0, 2.0), (3, 2, 5e-4, 10.0)]
    out["sgd_cfgs"] = np.asarray(cfgs, dtype=np.float64)
    for ci, (isz, gn, wd, lm) in enumerate(cfgs):
        op = RO.RefOp("ACMWeightDecayMomentumSGDUpdate", momentum=0.9, iter_size=isz, gpu_num=gn, lr_mult=lm,
                      weight_decay=wd)
        n = 257
        p = rng.standard_normal(n).astype(np.float32)
        m = rng.standard_normal(n).astype(np.float32)      # garbage: the op must zero it on call 0
        acc = rng.standard_normal(n).astype(np.float32)
        steps = 7
        G = rng.standard_normal((steps, n)).astype(np.float32)
        out["sgd%d_p0" % ci], out["sgd%d_m0" % ci], out["sgd%d_acc0" % ci], out["sgd%d_G" % ci] = p, m, acc, G
        P, M, A = [], [], []
        for s in range(steps):
            lr = np.array([1e-3 if s < 4 else 1e-4], np.float32)
            _, m, p, acc = op.run([G[s], m, lr, p, acc], 4, out_alias=[0, 1, 3, 4])
            P.append(p.copy()); M.append(m.copy()); A.append(acc.copy())
        out["sgd%d_P" % ci], out["sgd%d_M" % ci], out["sgd%d_A" % ci] = np.stack(P), np.stack(M), np.stack(A)
    np.savez_compressed(os.path.join(HERE, "ref_ops.npz"), **out)


def _mil_torch64(fc8c, fc8d, nfc8c, nfc8d, L, w_clean, w_noise, is_mean):
    """Independent float64 autograd statement of a5/a6/a8/a9 (weights are constants:
    StopGradient, webly_heads.py:390-391)."""
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    c, d, nc, nd = t(fc8c), t(fc8d), t(nfc8c), t(nfc8d)
    Lt = torch.tensor(np.asarray(L, np.float64))

    def stream(lc, ld):
        P = torch.softmax(lc, dim=1) * torch.softmax(ld, dim=0)
        return P, P.sum(0, keepdim=True)

    def bce(y, w):
        C = y.shape[1]
        val = -(w * (Lt * torch.log(y.clamp_min(1e-20)) + (1 - Lt) * torch.log((1 - y).clamp_min(1e-20)))).sum()
        return val / (C if is_mean else 1)

    P, y = stream(c, d)
    Pn, yn = stream(c + nc, d + nd)
    loss = bce(y, torch.tensor(np.asarray(w_clean, np.float64)))
    loss_n = bce(yn, torch.tensor(np.asarray(w_noise, np.float64)))
    (loss + loss_n).backward()
    g = lambda v: v.grad.numpy()
    return dict(P=P.detach().numpy(), y=y.detach().numpy(), Pn=Pn.detach().numpy(), yn=yn.detach().numpy(),
                loss=loss.item(), loss_n=loss_n.item(), d_fc8c=g(c), d_fc8d=g(d), d_nfc8c=g(nc), d_nfc8d=g(nd))


def gen_head_small():
    from oracle import nawsod_oracle as O
    rng = np.random.default_rng(300)
    out = {}
    for k, (R, C, soft) in enumerate([(37, 5, False), (64, 20, True), (150, 80, False)]):
        rois = O.synth_rois(R, 320, 480, seed=300 + k)
        fc8c, fc8d, nfc8c, nfc8d = [(rng.standard_normal((R, C)) * 2).astype(np.float32) for _ in range(4)]
        L = np.zeros((1, C), np.float32)
        L[0, rng.integers(C)] = 1
        if soft:
            lam = np.float32(rng.beta(1.5, 1.5))
            L = L * lam
            L[0, rng.integers(C)] += np.float32(1) - lam
        o = O.mil_head_forward_backward(fc8c, fc8d, rois, L, nfc8c, nfc8d, entropy=True, is_mean=True)
        # the float64 re-expression consumes the oracle's (forward-only) weights as constants
        ref = _mil_torch64(fc8c, fc8d, nfc8c, nfc8d, L, o["class_weight"], o["class_weight_noise"], True)
        # and an independent float64 statement of the noise weights themselves
        P64 = ref["P"]
        J = O.roi_iou(rois).astype(np.float64)
        E = np.where(P64 > 0, -P64 * np.log(np.where(P64 > 0, P64, 1.0)), 0.0)
        D = J @ E
        hat = (E * E / D).sum(0, keepdims=True)
        y64 = ref["y"]
        norm = np.clip(hat / (y64 * (np.log(R) - np.log(y64))), 0, 1)
        wn = norm * (1 - L.astype(np.float64))
        pre = "h%d_" % k
        out.update({pre + "rois": rois, pre + "fc8c": fc8c, pre + "fc8d": fc8d, pre + "nfc8c": nfc8c,
                    pre + "nfc8d": nfc8d, pre + "L": L, pre + "w_noise64": wn, pre + "w_clean64": 1 - wn})
        out.update({pre + "ref_" + kk: np.asarray(v) for kk, v in ref.items()})
    out["count"] = np.int32(3)
    np.savez_compressed(os.path.join(HERE, "head_small.npz"), **out)


if __name__ == "__main__":
    gen_roi_pool()
    gen_ref_ops()
    gen_head_small()
    print("golden fixtures written to", HERE)
