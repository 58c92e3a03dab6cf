"""GPU parity tests (run with -m gpu on the B200 box): every kernel of libnawsod.so, called
through the C ABI via the operator mirror, against the CPU oracle on the same seeded inputs
and against the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import nawsod_oracle as O
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def _ops():
    from nafwebsod_b200 import ops
    return ops


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


# ----------------------------------------------------------------------------------------------
# RoIPoolF
# ----------------------------------------------------------------------------------------------
def test_roi_pool_golden_bit_exact(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "roi_pool.npz"))
    # C = 6 is not a multiple of 4: pad channels to 8 (the pad planes are ignored)
    X = np.concatenate([g["X"], np.zeros((2, 2) + g["X"].shape[2:], np.float32)], axis=1)
    for rk, yk, ak, scale in (("rois", "Y", "argmax", 1 / 16), ("rois8", "Y8", "argmax8", 1 / 8)):
        Y, A = ops.RoIPoolF(dev(X), dev(g[rk]), spatial_scale=scale)
        assert np.array_equal(Y.cpu().numpy()[:, :6], g[yk])
        assert np.array_equal(A.cpu().numpy()[:, :6], g[ak])


@pytest.mark.parametrize("force_global", [0, 1])
@pytest.mark.parametrize("shape", [(1, 512, 38, 50, 2000), (2, 512, 38, 50, 4000), (1, 64, 75, 125, 300)])
def test_roi_pool_vs_oracle_bit_exact(shape, force_global):
    ops = _ops()
    import nafwebsod_b200 as pkg
    N, C, H, W, R = shape
    X = O.synth_conv5(N, C, H, W, seed=0)
    per = R // N
    rois = np.concatenate([O.synth_rois(per, H * 16, W * 16, b, seed=1 + b) for b in range(N)])
    obn = (np.random.default_rng(9).random(R) + 1).astype(np.float32)
    Yo, Ao = CO.roi_pool_f(X, rois, 1 / 16)
    pkg.set_tuning("pool_force_global", force_global)
    try:
        # reference layout in and out
        Y, A = ops.RoIPoolF(dev(X), dev(rois), spatial_scale=1 / 16)
        assert torch.equal(Y.cpu(), torch.from_numpy(Yo))
        assert torch.equal(A.cpu(), torch.from_numpy(Ao))
        # native layout + fused boost == RoIFeatureBoost(RoIPoolF(.))
        Xcl = ops.to_channels_last(dev(X))
        Yb, Ab = ops.RoIPoolF(Xcl, dev(rois), spatial_scale=1 / 16, boost=dev(obn), x_layout="NHWC", y_layout="NHWC")
        ref = O.roi_feature_boost(Yo, obn).transpose(0, 2, 3, 1)
        assert np.array_equal(Yb.cpu().numpy(), ref)
        assert np.array_equal(Ab.cpu().numpy(), Ao.transpose(0, 2, 3, 1))
        # inference: no argmax
        Yt, At = ops.RoIPoolF(Xcl, dev(rois), spatial_scale=1 / 16, is_test=True, x_layout="NHWC", y_layout="NHWC")
        assert At is None and np.array_equal(Yt.cpu().numpy(), Yo.transpose(0, 2, 3, 1))
    finally:
        pkg.set_tuning("pool_force_global", 0)


def test_roi_pool_bf16_path():
    """bf16 map in / bf16 out: max-pooling selects an input element, so the result equals the
    oracle run on the bf16-rounded map exactly (tolerance 0), argmax included."""
    ops = _ops()
    X = O.synth_conv5(1, 256, 38, 50, seed=3)
    rois = O.synth_rois(500, 608, 800, seed=4)
    Xb = dev(X).to(torch.bfloat16)
    Xr = Xb.float().cpu().numpy()
    Yo, Ao = CO.roi_pool_f(Xr, rois, 1 / 16)
    Xcl = Xb.permute(0, 2, 3, 1).contiguous()
    Y, A = ops.RoIPoolF(Xcl, dev(rois), spatial_scale=1 / 16, x_layout="NHWC", y_layout="NHWC")
    assert Y.dtype == torch.bfloat16
    assert np.array_equal(Y.float().cpu().numpy(), Yo.transpose(0, 2, 3, 1))
    assert np.array_equal(A.cpu().numpy(), Ao.transpose(0, 2, 3, 1))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roi_pool_rois_interleaved_across_images(dtype):
    """The kernels filter RoIs by batch index inside every (slab, chunk, image) CTA and CTAs whose chunk holds no RoI of
    their image leave before staging the map (pool_skip_idle): RoIs grouped by image (the loader's order) and randomly
    interleaved across three images must both equal the oracle."""
    ops = _ops()
    X = O.synth_conv5(3, 128, 38, 50, seed=0)
    grouped = np.concatenate([O.synth_rois(400, 608, 800, b, seed=1 + b) for b in range(3)])
    inter = grouped[np.random.default_rng(3).permutation(grouped.shape[0])]
    Xd = dev(X).to(dtype)
    Xr = Xd.float().cpu().numpy()
    Xcl = Xd.permute(0, 2, 3, 1).contiguous()
    for rois in (grouped, inter):
        Yo, Ao = CO.roi_pool_f(Xr, rois, 1 / 16)
        Y, A = ops.RoIPoolF(Xcl, dev(rois), spatial_scale=1 / 16, x_layout="NHWC", y_layout="NHWC")
        assert np.array_equal(Y.float().cpu().numpy(), Yo.transpose(0, 2, 3, 1))
        assert np.array_equal(A.cpu().numpy(), Ao.transpose(0, 2, 3, 1))


def _mixed_rois(R, img_h, img_w, batch, seed):
    """RoI mixture of BASELINE config 3: MCG-like boxes, boxes smaller than one cell / one bin row
    (bins repeat a map row), full-image boxes, boxes hanging over the border or fully outside,
    inverted boxes, and half-integer coordinates that exercise roundf()."""
    rng = np.random.default_rng(seed)
    base = O.synth_rois(R, img_h, img_w, batch, seed=seed)
    kind = rng.integers(0, 8, size=R)
    out = base.copy()
    for i in range(R):
        k = kind[i]
        if k == 0:      # tiny: 1..40 px on a side (less than 7 cells at 1/16 and often at 1/8)
            x1, y1 = rng.integers(0, img_w - 41), rng.integers(0, img_h - 41)
            out[i, 1:] = (x1, y1, x1 + rng.integers(0, 40), y1 + rng.integers(0, 40))
        elif k == 1:    # full image
            out[i, 1:] = (0, 0, img_w - 1, img_h - 1)
        elif k == 2:    # overhanging the border
            out[i, 1:] = (-60, -35, rng.integers(20, img_w // 2), rng.integers(20, img_h // 2))
        elif k == 3:    # overhanging bottom / right, or completely outside
            out[i, 1:] = (img_w - rng.integers(1, 90), img_h - rng.integers(1, 90), img_w + 100, img_h + 70)
            if rng.random() < 0.3:
                out[i, 1:] += (img_w, img_h, img_w, img_h)
        elif k == 4:    # inverted (x2 < x1): width clamps to 1
            out[i, 1:] = (out[i, 3], out[i, 4], out[i, 1], out[i, 2])
        elif k == 5:    # half-way coordinates: x*scale lands on .5 exactly
            out[i, 1:] = np.floor(out[i, 1:] / 16) * 16 + 8
    return out.astype(np.float32)


@pytest.mark.parametrize("knobs", [{}, {"pool_rows2": 0}, {"pool_rows2": 0, "pool_rowcache": 0}, {"pool_generic": 1},
                                   {"pool_force_global": 1}, {"pool_slab_bytes": 32 * 1024}, {"pool_rows2": 1}, {"pool_rows2": 2},
                                   {"pool_rows2": 2, "pool_force_global": 1}, {"pool_chunks": 3}, {"pool_skip_idle": 0}])
@pytest.mark.parametrize("cfg", [(2, 512, 38, 50, 1 / 16, 16), (1, 128, 75, 125, 1 / 16, 16), (1, 64, 60, 80, 1 / 8, 8),
                                 (1, 32, 150, 250, 1 / 8, 8)])      # the largest config-3 map: short side 1200, max side 2000 at 1/8
def test_roi_pool_mixed_rois_all_variants(cfg, knobs):
    """Every forward variant (both bin-row kernels, with / without the row cache, generic kernel, direct
    global reads, small slabs) on the config-3 RoI mixture, fp32 and bf16 maps, with and without
    argmax: bit-exact values and argmax against the C oracle."""
    ops = _ops()
    import nafwebsod_b200 as pkg
    N, C, H, W, scale, stride = cfg
    R = 700
    X = O.synth_conv5(N, C, H, W, seed=11)
    X[:, ::3] -= 0.25                    # negative planes: the running maximum must start below zero
    rois = np.concatenate([_mixed_rois(R // N, H * stride, W * stride, b, seed=20 + b) for b in range(N)])
    Xb = dev(X).to(torch.bfloat16)
    Yo, Ao = CO.roi_pool_f(X, rois, scale)
    Yob, Aob = CO.roi_pool_f(Xb.float().cpu().numpy(), rois, scale)
    for k, v in knobs.items():
        pkg.set_tuning(k, v)
    try:
        Xcl = ops.to_channels_last(dev(X))
        Y, A = ops.RoIPoolF(Xcl, dev(rois), spatial_scale=scale, x_layout="NHWC", y_layout="NHWC")
        assert np.array_equal(Y.cpu().numpy(), Yo.transpose(0, 2, 3, 1))
        assert np.array_equal(A.cpu().numpy(), Ao.transpose(0, 2, 3, 1))
        Yt, _ = ops.RoIPoolF(Xcl, dev(rois), spatial_scale=scale, is_test=True, x_layout="NHWC", y_layout="NHWC")
        assert np.array_equal(Yt.cpu().numpy(), Yo.transpose(0, 2, 3, 1))
        Xbcl = Xb.permute(0, 2, 3, 1).contiguous()
        Y2, A2 = ops.RoIPoolF(Xbcl, dev(rois), spatial_scale=scale, x_layout="NHWC", y_layout="NHWC")
        assert np.array_equal(Y2.float().cpu().numpy(), Yob.transpose(0, 2, 3, 1))
        assert np.array_equal(A2.cpu().numpy(), Aob.transpose(0, 2, 3, 1))
        Y3, _ = ops.RoIPoolF(Xbcl, dev(rois), spatial_scale=scale, is_test=True, x_layout="NHWC", y_layout="NHWC")
        assert np.array_equal(Y3.float().cpu().numpy(), Yob.transpose(0, 2, 3, 1))
    finally:
        for k in knobs:
            pkg.set_tuning(k, {"pool_rowcache": 1, "pool_slab_bytes": 200 * 1024, "pool_rows2": -1, "pool_skip_idle": 1}.get(k, 0))


def test_roi_pool_empty_and_errors():
    ops = _ops()
    X = dev(O.synth_conv5(1, 8, 6, 6))
    Y, A = ops.RoIPoolF(X, torch.zeros((0, 5), device="cuda"))
    assert tuple(Y.shape) == (0, 8, 7, 7) and tuple(A.shape) == (0, 8, 7, 7)
    with pytest.raises(RuntimeError):
        ops.RoIPoolF(X, torch.zeros((3, 4), device="cuda"))
    with pytest.raises(RuntimeError):          # C not a multiple of 4
        ops.RoIPoolF(dev(O.synth_conv5(1, 6, 6, 6)), torch.zeros((1, 5), device="cuda"))


@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
def test_roi_pool_backward(layout):
    ops = _ops()
    N, C, H, W, R = 2, 64, 38, 50, 600
    X = O.synth_conv5(N, C, H, W, seed=5)
    rois = np.concatenate([O.synth_rois(R // 2, 608, 800, b, seed=6 + b) for b in range(N)])
    obn = (np.random.default_rng(7).random(R) + 1).astype(np.float32)
    Yo, Ao = CO.roi_pool_f(X, rois, 1 / 16)
    dY = np.random.default_rng(8).standard_normal(Yo.shape).astype(np.float32)
    ref = CO.roi_pool_f_grad(X.shape, rois, Ao, O.roi_feature_boost_grad(dY, obn))
    if layout == "NCHW":
        dX = ops.RoIPoolFGradient(dev(X), dev(rois), dev(Ao), dev(dY), boost=dev(obn), layout="NCHW").cpu().numpy()
    else:
        dX = ops.RoIPoolFGradient(dev(X).permute(0, 2, 3, 1).contiguous(), dev(rois),
                                  dev(Ao.transpose(0, 2, 3, 1)), dev(dY.transpose(0, 2, 3, 1)), boost=dev(obn),
                                  layout="NHWC").cpu().numpy().transpose(0, 3, 1, 2)
    # atomics: summation order differs from the sequential oracle -> fp32 rounding only
    np.testing.assert_allclose(dX, ref, rtol=1e-4, atol=1e-4)
    assert np.array_equal(dX == 0, ref == 0)


def test_boost_op_matches_reference_golden(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "ref_ops.npz"))
    Y = ops.RoIFeatureBoost(dev(g["boost_X"]), dev(g["boost_S"]))
    assert np.array_equal(Y.cpu().numpy(), g["boost_Y"])
    x = dev(g["boost_X"])
    ops.RoIFeatureBoost(x, dev(g["boost_S"]), out=x)            # in place (AllowInplace {{0,0}})
    assert np.array_equal(x.cpu().numpy(), g["boost_Y"])
    assert np.array_equal(ops.RoIFeatureBoostGradient(dev(g["boost_X"]), dev(g["boost_S"])).cpu().numpy(), g["boost_dX"])
    with pytest.raises(RuntimeError):
        ops.RoIFeatureBoost(dev(g["boost_X"]), dev(g["boost_S"][:3]))


# ----------------------------------------------------------------------------------------------
# RoIIoU / CE / SGD against reference-pinned fixtures
# ----------------------------------------------------------------------------------------------
def test_roi_iou_bit_exact():
    ops = _ops()
    rois = O.synth_rois(777, 1200, 2000, seed=3)
    rois[:, 1:] *= np.float32(1.37)
    J = ops.RoIIoU(dev(rois)).cpu().numpy()
    assert np.array_equal(J, O.roi_iou(rois))


def test_cross_entropy_matches_reference_golden(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "ref_ops.npz"))
    one = torch.ones(1, device="cuda")
    for k in range(int(g["ce_count"])):
        pre = "ce%d_" % k
        x, l, w, im = dev(g[pre + "x"]), dev(g[pre + "l"]), dev(g[pre + "w"]), bool(g[pre + "is_mean"])
        # forward: double log on the GPU vs glibc -> allow 1 ulp of float32
        np.testing.assert_allclose(ops.WeightedCrossEntropyWithLogits(x, l, w, is_mean=im).item(), g[pre + "loss_w"], rtol=2e-7)
        np.testing.assert_allclose(ops.CrossEntropyWithLogits(x, l, is_mean=im).item(), g[pre + "loss_u"], rtol=2e-7)
        # gradient: pure float32 arithmetic -> bit-exact
        assert np.array_equal(ops.WeightedCrossEntropyWithLogitsGradient(x, l, w, one, is_mean=im).cpu().numpy(), g[pre + "grad_w"])
        assert np.array_equal(ops.CrossEntropyWithLogitsGradient(x, l, one, is_mean=im).cpu().numpy(), g[pre + "grad_u"])


def test_sgd_matches_reference_golden(golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "ref_ops.npz"))
    for ci, (isz, gn, wd, lm) in enumerate(g["sgd_cfgs"]):
        p, m, acc = dev(g["sgd%d_p0" % ci]), dev(g["sgd%d_m0" % ci]), dev(g["sgd%d_acc0" % ci])
        G = g["sgd%d_G" % ci]
        shadow = torch.zeros_like(p, dtype=torch.bfloat16)
        for s in range(G.shape[0]):
            lr = torch.tensor([1e-3 if s < 4 else 1e-4], dtype=torch.float32, device="cuda")
            ops.ACMWeightDecayMomentumSGDUpdate(dev(G[s]), m, lr, p, acc, momentum=0.9, iter_size=int(isz),
                                                gpu_num=int(gn), lr_mult=lm, weight_decay=wd, iter_count=s,
                                                p_shadow=shadow)
            assert np.array_equal(p.cpu().numpy(), g["sgd%d_P" % ci][s])       # bit-exact with the reference op
            assert np.array_equal(m.cpu().numpy(), g["sgd%d_M" % ci][s])
            assert np.array_equal(acc.cpu().numpy(), g["sgd%d_A" % ci][s])
            if (s + 1) % int(isz) == 0:
                assert torch.equal(shadow, p.to(torch.bfloat16))




def test_sgd_large_no_acc():
    ops = _ops()
    rng = np.random.default_rng(0)
    n = 4096 * 1024 + 3
    p0, g0 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    m0 = rng.standard_normal(n).astype(np.float32)
    p, m = dev(p0), dev(m0)
    lr = torch.tensor([1e-3], device="cuda")
    ops.ACMWeightDecayMomentumSGDUpdate(dev(g0), m, lr, p, None, weight_decay=5e-4, gpu_num=8, iter_count=5)
    mo, po, _, _ = O.acm_sgd_update(g0, m0, 1e-3, p0, np.zeros(n, np.float32), weight_decay=5e-4, gpu_num=8, iter_count=5)
    assert np.array_equal(p.cpu().numpy(), po) and np.array_equal(m.cpu().numpy(), mo)


# ----------------------------------------------------------------------------------------------
# fused MIL head
# ----------------------------------------------------------------------------------------------
TOL = 1e-3   # north_star: MIL scores, loss and gradients rel <= 1e-3 in fp32


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.abs(b).max() + 1e-30
    assert np.abs(a - b).max() <= tol * scale, (np.abs(a - b).max(), scale)


def _run_mil(fc8c, fc8d, rois, L, nfc8c, nfc8d, offs, **kw):
    ops = _ops()
    o = ops.mil_head(dev(fc8c), dev(fc8d), dev(rois), torch.tensor(offs, dtype=torch.int32, device="cuda"), dev(L),
                     None if nfc8c is None else dev(nfc8c), None if nfc8d is None else dev(nfc8d), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in o.items()}


def test_mil_head_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "head_small.npz"))
    for k in range(int(g["count"])):
        pre = "h%d_" % k
        R = g[pre + "fc8c"].shape[0]
        o = _run_mil(g[pre + "fc8c"], g[pre + "fc8d"], g[pre + "rois"], g[pre + "L"], g[pre + "nfc8c"],
                     g[pre + "nfc8d"], [0, R])
        _close(o["rois_pred"], g[pre + "ref_P"])
        _close(o["cls_prob"], g[pre + "ref_y"])
        _close(o["cls_prob_noise"], g[pre + "ref_yn"])
        _close(o["class_weight_noise"], g[pre + "w_noise64"])
        _close(o["loss"][0, 0], g[pre + "ref_loss"])
        _close(o["loss"][0, 1], g[pre + "ref_loss_n"])
        for name in ("d_fc8c", "d_fc8d", "d_nfc8c", "d_nfc8d"):
            _close(o[name], g[pre + "ref_" + name])


@pytest.mark.parametrize("cfg", [(2000, 20, 2, False), (4000, 80, 1, True), (333, 20, 3, False)])
def test_mil_head_vs_oracle(cfg):
    """B images per call == the reference run once per image."""
    R, C, B, soft = cfg
    rng = np.random.default_rng(R + C)
    per = R // B
    offs = [i * per for i in range(B)] + [R]
    rois = np.concatenate([O.synth_rois(offs[b + 1] - offs[b], 608, 800, b, seed=20 + b) for b in range(B)])
    fc8c, fc8d, nfc8c, nfc8d = [(rng.standard_normal((R, C)) * 1.5).astype(np.float32) for _ in range(4)]
    L = np.zeros((B, C), np.float32)
    for b in range(B):
        L[b, rng.integers(C)] = 1
        if soft:
            lam = np.float32(rng.beta(1.5, 1.5)); L[b] *= lam; L[b, rng.integers(C)] += np.float32(1) - lam
    o = _run_mil(fc8c, fc8d, rois, L, nfc8c, nfc8d, offs)
    for b in range(B):
        s = slice(offs[b], offs[b + 1])
        ref = O.mil_head_forward_backward(fc8c[s], fc8d[s], rois[s], L[b:b + 1], nfc8c[s], nfc8d[s])
        _close(o["rois_pred"][s], ref["rois_pred"])
        _close(o["rois_pred_noise"][s], ref["rois_pred_noise"])
        _close(o["cls_prob"][b], ref["cls_prob"][0])
        _close(o["cls_prob_noise"][b], ref["cls_prob_noise"][0])
        _close(o["class_weight"][b], ref["class_weight"][0])
        _close(o["class_weight_noise"][b], ref["class_weight_noise"][0])
        _close(o["loss"][b, 0], ref["loss_cls"])
        _close(o["loss"][b, 1], ref["loss_cls_noise"])
        for name in ("d_fc8c", "d_fc8d", "d_nfc8c", "d_nfc8d"):
            _close(o[name][s], ref[name])


def test_mil_head_plain_wsddn_and_no_entropy():
    R, C = 500, 20
    rng = np.random.default_rng(5)
    rois = O.synth_rois(R, 608, 800, seed=2)
    fc8c, fc8d, nfc8c, nfc8d = [(rng.standard_normal((R, C))).astype(np.float32) for _ in range(4)]
    L = np.zeros((1, C), np.float32); L[0, 7] = 1
    o = _run_mil(fc8c, fc8d, rois, L, None, None, [0, R])                     # single-stack WSDDN
    ref = O.mil_head_forward_backward(fc8c, fc8d, rois, L)
    _close(o["loss"][0, 0], ref["loss_cls"]); _close(o["d_fc8c"], ref["d_fc8c"]); _close(o["d_fc8d"], ref["d_fc8d"])
    o = _run_mil(fc8c, fc8d, rois, L, nfc8c, nfc8d, [0, R], entropy=False, is_mean=False)
    ref = O.mil_head_forward_backward(fc8c, fc8d, rois, L, nfc8c, nfc8d, entropy=False, is_mean=False)
    _close(o["loss"][0, 0], ref["loss_cls"]); _close(o["loss"][0, 1], ref["loss_cls_noise"])
    for name in ("d_fc8c", "d_fc8d", "d_nfc8c", "d_nfc8d"):
        _close(o[name], ref[name])
    # size-independent property: y_c = sum_r a_cls*a_det with sum_r a_det = 1 and a_cls <= 1 -> each y_c in [0,1]
    assert (o["cls_prob"] >= 0).all() and (o["cls_prob"] <= 1 + 1e-5).all()
