"""CPU suite: the oracle against the committed golden vectors (and, when the build
container has it, against the reference's own compiled operators)."""
import os

import numpy as np
import pytest

from oracle import nawsod_oracle as O
from oracle import c_oracle as CO
from oracle import ref_ops as RO


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ----------------------------------------------------------------------------- RoIPoolF
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_roi_pool_matches_torchvision_golden(golden_dir, impl):
    g = _g(golden_dir, "roi_pool.npz")
    f = O.roi_pool_f if impl == "numpy" else CO.roi_pool_f
    for rk, yk, ak, scale in (("rois", "Y", "argmax", 1 / 16), ("rois8", "Y8", "argmax8", 1 / 8)):
        Y, A = f(g["X"], g[rk], scale)
        assert np.array_equal(Y, g[yk])            # bit-exact values
        assert np.array_equal(A, g[ak])            # bit-exact int32 argmax


def test_roi_pool_edge_cases(golden_dir):
    g = _g(golden_dir, "roi_pool.npz")
    Y, A = CO.roi_pool_f(g["X"], g["rois"], 1 / 16)
    rois = g["rois"]
    outside = np.where((rois[:, 1] == 1000))[0][0]
    assert (Y[outside] == 0).all() and (A[outside] == -1).all()       # empty bins -> 0 / -1
    # empty rois / zero-size problem
    Y0, A0 = CO.roi_pool_f(g["X"], np.zeros((0, 5), np.float32), 1 / 16)
    assert Y0.shape == (0, g["X"].shape[1], 7, 7) and A0.shape == Y0.shape
    # argmax always indexes the value it reports
    X = g["X"]
    for r in range(rois.shape[0]):
        b = int(rois[r, 0])
        plane = X[b].reshape(X.shape[1], -1)
        m = A[r].reshape(X.shape[1], -1)
        for c in range(X.shape[1]):
            sel = m[c] >= 0
            assert np.array_equal(plane[c][m[c][sel]], Y[r, c].reshape(-1)[sel])


def test_roi_pool_grad_c_vs_numpy():
    X = O.synth_conv5(2, 16, 20, 25, seed=5)
    rois = np.concatenate([O.synth_rois(30, 320, 400, 0, seed=6), O.synth_rois(30, 320, 400, 1, seed=7)])
    Y, A = CO.roi_pool_f(X, rois, 1 / 16)
    dY = np.random.default_rng(8).standard_normal(Y.shape).astype(np.float32)
    a = O.roi_pool_f_grad(X.shape, rois, A, dY)
    b = CO.roi_pool_f_grad(X.shape, rois, A, dY)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5)
    # adjoint identity: <dY, pool(X)> gradient wrt X picks exactly the argmax cells
    assert np.isclose((b * X).sum(), (dY * Y).sum(), rtol=1e-4)


# ------------------------------------------------------------- ops pinned by the reference
def test_boost_matches_reference_golden(golden_dir):
    g = _g(golden_dir, "ref_ops.npz")
    assert np.array_equal(O.roi_feature_boost(g["boost_X"], g["boost_S"]), g["boost_Y"])
    assert np.array_equal(O.roi_feature_boost_grad(g["boost_X"], g["boost_S"]), g["boost_dX"])


def test_cross_entropy_matches_reference_golden(golden_dir):
    g = _g(golden_dir, "ref_ops.npz")
    for k in range(int(g["ce_count"])):
        pre = "ce%d_" % k
        x, l, w, im = g[pre + "x"], g[pre + "l"], g[pre + "w"], bool(g[pre + "is_mean"])
        assert O.cross_entropy_with_logits(x, l, w, im) == g[pre + "loss_w"]
        assert O.cross_entropy_with_logits(x, l, None, im) == g[pre + "loss_u"]
        assert np.array_equal(O.cross_entropy_with_logits_grad(x, l, 1.0, w, im), g[pre + "grad_w"])
        assert np.array_equal(O.cross_entropy_with_logits_grad(x, l, 1.0, None, im), g[pre + "grad_u"])


def test_sgd_matches_reference_golden(golden_dir):
    g = _g(golden_dir, "ref_ops.npz")
    for ci, (isz, gn, wd, lm) in enumerate(g["sgd_cfgs"]):
        p, m, acc, it = g["sgd%d_p0" % ci], g["sgd%d_m0" % ci], g["sgd%d_acc0" % ci], 0
        G = g["sgd%d_G" % ci]
        for s in range(G.shape[0]):
            lr = np.float32(1e-3 if s < 4 else 1e-4)
            m, p, acc, it = O.acm_sgd_update(G[s], m, lr, p, acc, momentum=0.9, weight_decay=wd, lr_mult=lm,
                                             iter_size=int(isz), gpu_num=int(gn), iter_count=it)
            assert np.array_equal(p, g["sgd%d_P" % ci][s])
            assert np.array_equal(m, g["sgd%d_M" % ci][s])
            assert np.array_equal(acc, g["sgd%d_A" % ci][s])


@pytest.mark.skipif(not RO.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_library_live():
    """When oracle/_ref exists, run the reference's real operators again on fresh inputs."""
    rng = np.random.default_rng(1)
    x = (rng.random((1, 20)) ** 2).astype(np.float32)
    l = np.zeros((1, 20), np.float32); l[0, 4] = 1
    w = rng.random((1, 20)).astype(np.float32)
    assert RO.RefOp("WeightedCrossEntropyWithLogits", is_mean=1).run([x, l, w], 1)[0] == \
        O.cross_entropy_with_logits(x, l, w, True)
    with pytest.raises(RuntimeError):                       # CAFFE_ENFORCE_EQ(X.sizes(), L.sizes())
        RO.RefOp("WeightedCrossEntropyWithLogits", is_mean=1).run([x, l[:, :10], w], 1)


# ------------------------------------------------------------------ MIL head vs fp64 autograd
def test_mil_head_matches_float64_autograd(golden_dir):
    g = _g(golden_dir, "head_small.npz")
    for k in range(int(g["count"])):
        pre = "h%d_" % k
        o = O.mil_head_forward_backward(g[pre + "fc8c"], g[pre + "fc8d"], g[pre + "rois"], g[pre + "L"],
                                        g[pre + "nfc8c"], g[pre + "nfc8d"], entropy=True, is_mean=True)
        np.testing.assert_allclose(o["rois_pred"], g[pre + "ref_P"], rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(o["cls_prob"], g[pre + "ref_y"], rtol=2e-5)
        np.testing.assert_allclose(o["cls_prob_noise"], g[pre + "ref_yn"], rtol=2e-5)
        np.testing.assert_allclose(o["class_weight_noise"], g[pre + "w_noise64"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["class_weight"], g[pre + "w_clean64"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["loss_cls"], g[pre + "ref_loss"], rtol=1e-5)
        np.testing.assert_allclose(o["loss_cls_noise"], g[pre + "ref_loss_n"], rtol=1e-5)
        for name in ("d_fc8c", "d_fc8d", "d_nfc8c", "d_nfc8d"):
            ref = g[pre + "ref_" + name]
            np.testing.assert_allclose(o[name], ref, rtol=1e-3, atol=1e-6 * np.abs(ref).max() + 1e-12)


def test_roi_iou_properties():
    rois = O.synth_rois(50, 300, 400, seed=3)
    rois[:, 1:] += np.float32(0.7)                      # fractional coords are truncated, not rounded
    J = O.roi_iou(rois)
    assert np.array_equal(J, J.T) and (np.diag(J) == 1).all()
    assert (J >= 0).all() and (J <= 1).all()
    i, j = 3, 9
    a, b = rois[i, 1:].astype(np.int32), rois[j, 1:].astype(np.int32)
    iw = max(min(a[2], b[2]) - max(a[0], b[0]) + 1, 0); ih = max(min(a[3], b[3]) - max(a[1], b[1]) + 1, 0)
    ua = (a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - iw * ih
    assert J[i, j] == np.float32(np.float32(iw * ih) / np.float32(ua))


def test_full_head_small_runs_and_grad_check():
    """Whole head on a tiny problem; finite-difference check of one fc7 weight through
    everything (the reference's gradient-check thresholds: stepsize/threshold 0.005)."""
    rng = np.random.default_rng(11)
    C, D, Hd, R = 4, 8 * 49, 16, 12
    X = O.synth_conv5(1, 8, 10, 12, seed=12)
    rois = O.synth_rois(R, 160, 192, seed=13)
    obn = (rng.random((R, 1)) + 1).astype(np.float32)
    L = np.zeros((1, C), np.float32); L[0, 1] = 1
    p = O.synth_params(C, D, Hd, noise=True, seed=14)
    for k in p:
        if k.endswith("_w"):
            p[k] = (p[k] * 10).astype(np.float32)
    # freeze the (forward-only) class weights by disabling entropy so the loss is smooth in W
    o = O.head_forward_backward(X, rois, obn, L, p, noise=True, entropy=False)
    g = o["grads"]["fc7_w"]
    idx = np.unravel_index(np.argmax(np.abs(g)), g.shape)
    eps = 5e-3
    def total(pp):
        r = O.head_forward_backward(X, rois, obn, L, pp, noise=True, entropy=False, backward=False)
        return float(r["loss_cls"]) + float(r["loss_cls_noise"])
    pp = {k: v.copy() for k, v in p.items()}; pp["fc7_w"][idx] += eps; up = total(pp)
    pp["fc7_w"][idx] -= 2 * eps; dn = total(pp)
    num = (up - dn) / (2 * eps)
    assert abs(num - g[idx]) <= 5e-3 * max(1.0, abs(num))
