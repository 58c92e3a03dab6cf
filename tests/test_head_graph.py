"""The oracle's head against the reference's OWN graph builders (tests/golden/head_graph.npz, made by
tests/golden/make_golden_head_graph.py): add_VGG16_roi_2fc_noise_head, add_webly_outputs, add_webly_losses,
add_spatial_entropy_weight, add_cls_pred, add_cross_entropy_loss and RoIFeatureTransform were imported unmodified from
/root/reference and executed operator by operator on an eager NumPy workspace, with the reference's own code for the
operators that live in its tree.  This pins the WIRING of SURVEY.md section 8 rows a3-a8 (which operator, on which blobs,
in which order, with which axes / flags) to the reference; the arithmetic of the Caffe2 built-ins stays a float32
restatement of their documented defaults.  The GPU test of the same vectors is in tests/test_gpu_zzz_reference_vectors.py."""
import os

import numpy as np
import pytest

from oracle import nawsod_oracle as O


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "head_graph.npz"))


def _maker():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_head_graph", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                                       "golden", "make_golden_head_graph.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                 # NumPy + the oracle package only; /root/reference is touched by its main() alone
    return mod


def case_inputs(gold, i):
    """Inputs of golden case i: the stored blobs, the bit-packed dropout masks, the parameters regenerated from the case's
    seed (checksums stored in the file guard the regeneration)."""
    X, rois, obn, L, params, masks, cfg = _maker().load_case(gold, i)
    return X, rois, obn, L, params, masks, cfg["train"]


def out(gold, i, name):
    return gold["case%d_out_%s" % (i, name)]


def close(a, b, rtol=2e-5, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("i", [0, 1])
def test_training_graph_blobs(gold, i):
    X, rois, obn, L, params, masks, train = case_inputs(gold, i)
    assert train
    ref = O.head_forward_backward(X, rois, obn, L, params, masks=masks, noise=True, entropy=True, is_mean=True, backward=False)
    # RoIPoolF -> RoIFeatureBoost -> StopGradient: the same code on both sides, exact
    assert np.array_equal(ref["roi_feat"], out(gold, i, "roi_feat").reshape(ref["roi_feat"].shape))
    # FC / Relu / Dropout x2 per stack, fc8c | fc8d per stack (float32 GEMMs: summation order only)
    close(ref["acts"]["fc6"], out(gold, i, "fc6"))
    close(ref["acts"]["drop7"], out(gold, i, "drop7"))
    close(ref["noisy_acts"]["drop7"], out(gold, i, "_[noisy]_drop7"))
    for mine, theirs in (("fc8c", "fc8c"), ("fc8d", "fc8d"), ("nfc8c", "noisy_fc8c"), ("nfc8d", "noisy_fc8d")):
        close(ref[mine], out(gold, i, theirs), rtol=1e-4, atol=1e-6)
    # the two-stream MIL outputs of both streams, the image scores
    for k in ("rois_pred", "rois_pred_noise", "cls_prob", "cls_prob_noise"):
        close(ref[k], out(gold, i, k), rtol=1e-4, atol=1e-9)
    # the noise-aware class weights and both losses
    close(ref["class_weight"], out(gold, i, "rois_class_weight"), rtol=1e-4, atol=1e-6)
    close(ref["class_weight_noise"], out(gold, i, "rois_class_weight_noise"), rtol=1e-4, atol=1e-6)
    close(ref["loss_cls"], out(gold, i, "loss_cls"), rtol=1e-4)
    close(ref["loss_cls_noise"], out(gold, i, "loss_cls_noise"), rtol=1e-4)
    # loss-gradient seeds are 1.0 per loss, no 1/NUM_GPUS (utils/blob.py:167-173)
    assert float(out(gold, i, "loss_cls_grad")) == 1.0 and float(out(gold, i, "loss_cls_noise_grad")) == 1.0
    assert list(gold["case%d_losses" % i]) == ["loss_cls", "loss_cls_noise"]


@pytest.mark.parametrize("i", [0, 1])
def test_entropy_weight_intermediates(gold, i):
    """add_spatial_entropy_weight blob by blob: J, E = -P log P (NaN -> 0), D = LeakyRelu(J E), sum_r E^2 / D and its
    normalisation by y (log R - log y), clipped to [0, 1]."""
    P, y = out(gold, i, "rois_pred"), out(gold, i, "cls_prob")
    rois, L = gold["case%d_in_rois" % i], gold["case%d_in_labels" % i]
    assert np.array_equal(O.roi_iou(rois), out(gold, i, "rois_J"))
    w_clean, w_noise, parts = O.spatial_entropy_weight(P, y, rois, L, return_parts=True)
    close(parts["E"], out(gold, i, "rois_pred_E"), rtol=1e-6, atol=0)
    close(parts["D"], out(gold, i, "rois_pred_D"), rtol=1e-4, atol=1e-9)
    close(parts["hatE_sum"], out(gold, i, "rois_pred_hatE_sum"), rtol=1e-4)
    close(parts["norm"], out(gold, i, "rois_pred_hatE_sum_norm"), rtol=1e-4, atol=1e-7)
    close(w_noise, out(gold, i, "rois_class_weight_noise"), rtol=1e-4, atol=1e-7)
    close(w_clean, out(gold, i, "rois_class_weight"), rtol=1e-4, atol=1e-7)


def test_test_mode_graph(gold):
    """model.train == False: no Dropout, no losses, cls_prob = Concat(Split(rois_pred)[0], rois_pred) [R, C+1]."""
    X, rois, obn, L, params, masks, train = case_inputs(gold, 2)
    assert not train and masks is None
    ref = O.head_forward_backward(X, rois, obn, L, params, masks=None, noise=True, backward=False)
    close(ref["rois_pred"], out(gold, 2, "rois_pred"), rtol=1e-4, atol=1e-9)
    cp = O.test_cls_prob(ref["rois_pred"])
    assert cp.shape == out(gold, 2, "cls_prob").shape == (rois.shape[0], L.shape[1] + 1)
    close(cp, out(gold, 2, "cls_prob"), rtol=1e-4, atol=1e-9)
    assert not any(t.startswith("Dropout") for t in gold["case2_trace"])


def test_every_emitted_operator_is_accounted_for(gold):
    """The builders' operator trace: nothing outside what the oracle restates (or deliberately ignores), the live branch of
    the entropy normalisation, and the order FC -> Relu -> Dropout of both stacks."""
    trace = [str(t) for t in gold["case0_trace"]]
    kinds = [t.split("(")[0] for t in trace]
    restated = {"RoIPoolF", "RoIFeatureBoost", "FC", "Relu", "Dropout", "Softmax", "Transpose", "Add", "Sub", "Mul", "Div",
                "ReduceSum", "RoIIoU", "Log", "Scale", "ReplaceNaN", "MatMul", "LeakyRelu", "Shape", "Cast", "Clip",
                "ConstantFill", "WeightedCrossEntropyWithLogits", "AveragedLoss"}
    ignored = {"StopGradient", "Stat", "Accuracy"}          # identity in the forward pass / logging / metric
    assert set(kinds) <= restated | ignored, set(kinds) - restated - ignored
    assert "Tile" not in kinds                               # the `if True and False` branch (webly_heads.py:294-332) is dead
    assert any(t.startswith("Div(rois_pred_hatE_sum,rois_pred_y_logN__logy)") for t in trace)
    assert kinds[:9] == ["RoIPoolF", "RoIFeatureBoost", "StopGradient", "FC", "Relu", "Dropout", "FC", "Relu", "Dropout"]
    assert kinds.count("FC") == 8 and kinds.count("WeightedCrossEntropyWithLogits") == 2 and kinds.count("RoIIoU") == 1
    ce = [t for t in trace if t.startswith("WeightedCrossEntropyWithLogits")]
    assert ce[0].startswith("WeightedCrossEntropyWithLogits(cls_prob,labels_oh,rois_class_weight)->(cross_entropy)")
    assert ce[1].startswith("WeightedCrossEntropyWithLogits(cls_prob_noise,labels_oh,rois_class_weight_noise)->(cross_entropy_noise)")
    assert all("('is_mean', 'True')" in t for t in ce)
    # both class-weight vectors are constants for the backward pass (webly_heads.py:390-391); roi_feat too (FREEZE_CONV_BODY)
    for blob in ("rois_class_weight", "rois_class_weight_noise", "roi_feat"):
        assert "StopGradient(%s)->(%s)" % (blob, blob) in trace


# ---------------------------------------------------------------------------------------------------------------
# row a10 wiring: add_single_gpu_param_update_ops (modeling/optimizer_wsl.py:75-137) run on the same eager helper
# ---------------------------------------------------------------------------------------------------------------
def test_optimizer_builder_arguments_and_trajectories(gold):
    params = [str(p) for p in gold["opt_params"]]
    assert int(gold["opt_gpu_num"]) == 4 and int(gold["opt_iter_size"]) == 1          # the shipped flickr_voc config
    for p in params:
        mom, isz, gn, lr_mult, wd = gold["opt_args_" + p]
        bias = p.endswith("_b")
        # biases: no weight decay, 2x learning rate; weights: SOLVER.WEIGHT_DECAY, 1x (optimizer_wsl.py:106-123)
        assert (mom, isz, gn, lr_mult, wd) == (0.9, 1.0, 4.0, 2.0 if bias else 1.0, 0.0 if bias else 5e-4), (p, gold["opt_args_" + p])
        # three runs of the update net == the oracle's restatement with those arguments, bit for bit
        pv, mv = gold["opt_p0_" + p].copy(), np.zeros_like(gold["opt_p0_" + p])
        for s, lr in enumerate(gold["opt_lrs"]):
            mv, pv, _, _ = O.acm_sgd_update(gold["opt_G_" + p][s], mv, lr, pv, np.zeros_like(pv), momentum=0.9, weight_decay=wd,
                                            lr_mult=lr_mult, gpu_num=int(gn), iter_count=s)
            assert np.array_equal(pv, gold["opt_P_" + p][s]) and np.array_equal(mv, gold["opt_M_" + p][s]), (p, s)
    kinds = [str(t).split("(")[0] for t in gold["opt_trace"]]
    assert kinds.count("ACMWeightDecayMomentumSGDUpdate") == len(params) == 16
    assert set(kinds) == {"ConstantFill", "ACMWeightDecayMomentumSGDUpdate"}


def test_host_update_wiring_matches_the_reference_builder(gold, monkeypatch):
    """dp.DataParallelHead._update_slice (the product's per-bucket update call): weights and biases get the arguments the
    reference's builder gives them, gpu_num is the world size."""
    import torch
    from nafwebsod_b200 import dp, ops
    seen = []
    monkeypatch.setattr(ops, "ACMWeightDecayMomentumSGDUpdate", lambda g, m, lr, p, acc, **kw: seen.append(kw))

    class M:
        flat_grad = flat_mom = flat_param = flat_lp = torch.zeros(64)
        lr, iter_count = torch.zeros(1), 5
    head = dp.DataParallelHead.__new__(dp.DataParallelHead)
    head.model, head.world, head._hyper = M(), 4, dict(momentum=0.9, weight_decay=5e-4)
    for tag in ("fc6_panel", "small_weights", "biases"):
        head._update_slice(0, 64, tag, 0, 64)
    want_w = gold["opt_args_fc6_w"]
    want_b = gold["opt_args_fc6_b"]
    for kw, want in zip(seen, (want_w, want_w, want_b)):
        assert (kw["momentum"], 1.0, float(kw["gpu_num"]), kw["lr_mult"], kw["weight_decay"]) == tuple(want)
        assert kw["iter_count"] == 5


def test_exported_parameter_names_are_the_builders_blob_names(gold):
    """Checkpoint contract (detectron/utils/net_wsl.py:140-181 saves every parameter under its blob name): the names the
    host model exports are exactly the weight / bias blobs the reference's builders hand to their FC operators."""
    import torch
    from nafwebsod_b200.heads import WeblyHeadModel
    fc = [str(t) for t in gold["case0_trace"] if str(t).startswith("FC(")]
    blobs = set()
    for t in fc:
        _, w, b = t[3:t.index(")")].split(",")
        blobs |= {w, b}
    m = WeblyHeadModel(21, 16, 7, 64, noise=True, dtype=torch.float32, device="cpu")
    exported = m.export_reference_params()
    assert set(exported) == blobs, set(exported) ^ blobs
    assert tuple(exported["_[noisy]_fc6_w"].shape) == (64, 16 * 49) and tuple(exported["noisy_fc8d_b"].shape) == (20,)


def _tf32(x):
    """Round to the nearest TF32 value (10-bit mantissa, ties away: cvt.rna.tf32.f32), kept in float32."""
    x = np.asarray(x, np.float32)
    u = (x.view(np.uint32).astype(np.uint64) + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def test_tf32_rounding_points_leave_headroom_on_the_gpu_case(gold):
    """The product's fp32 path feeds the tensor cores TF32 operands rounded to nearest at four points (the pooled features,
    the weight shadow, the fc6 and the fc7 outputs: DESIGN.md section 2, 'TF32').  Emulating exactly those roundings on
    the CPU (everything else float32) predicts how far the GPU run of golden case 0 can be from the reference-built
    float32 vectors: it must stay at most half the 1e-3 bar the GPU test asserts, so that test is not a coin flip."""
    X, rois, obn, L, params, masks, _ = case_inputs(gold, 0)
    Y, _ = O.roi_pool_f(X, rois, 1.0 / 16)
    feat = _tf32(O.roi_feature_boost(Y, obn).reshape(Y.shape[0], -1))

    def stack(pfx):
        f6 = np.maximum(feat @ _tf32(params[pfx + "fc6_w"]).T + params[pfx + "fc6_b"], 0) * masks[pfx + "drop6"] * 2
        f7 = np.maximum(_tf32(f6.astype(np.float32)) @ _tf32(params[pfx + "fc7_w"]).T + params[pfx + "fc7_b"], 0) * masks[pfx + "drop7"] * 2
        return _tf32(f7.astype(np.float32))
    d7, nd7 = stack(""), stack("noisy_")
    fc8 = lambda x, k: (x @ _tf32(params[k + "_w"]).T + params[k + "_b"]).astype(np.float32)
    got = O.mil_head_forward_backward(fc8(d7, "fc8c"), fc8(d7, "fc8d"), rois, L, fc8(nd7, "noisy_fc8c"), fc8(nd7, "noisy_fc8d"),
                                      backward=False)
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel()) /
                             np.linalg.norm(np.asarray(b, np.float64).ravel()))
    worst = max(rel(got[k], out(gold, 0, k)) for k in ("rois_pred", "rois_pred_noise", "cls_prob", "cls_prob_noise"))
    assert 1e-5 < worst <= 6e-4, worst                  # TF32 does cost ~5e-4 here -- and no more
    assert rel(got["class_weight_noise"], out(gold, 0, "rois_class_weight_noise")) <= 5e-4
    assert abs(got["loss_cls"] - out(gold, 0, "loss_cls")) <= 2e-4 * abs(out(gold, 0, "loss_cls"))


def test_weights_file_initialisation_follows_the_reference(gold):
    """heads.WeblyHeadModel.initialize_from_weights against initialize_gpu_from_weights_file (utils/net_wsl.py:53-137) run
    unmodified on a dictionary workspace: the `]_` rule (`_[noisy]_fc6/fc7` start from the clean stack's blobs unless the
    file holds them), missing blobs keep their initialisation, momentum blobs are loaded along."""
    import torch
    from nafwebsod_b200.heads import WeblyHeadModel
    Cc, hidden, C = (int(v) for v in gold["winit_cfg"])
    src = {k[len("winit_src_"):]: gold[k] for k in gold.files if k.startswith("winit_src_")}
    ws = {k[len("winit_ws_"):]: gold[k] for k in gold.files if k.startswith("winit_ws_") and k != "winit_ws_names"}
    m = WeblyHeadModel(C + 1, Cc, 7, hidden, noise=True, dtype=torch.float32, device="cpu")
    loaded = m.initialize_from_weights({"blobs": src})
    params = m.export_reference_params()
    assert set(params) == set(str(p) for p in gold["winit_params"])
    fed = {k for k in ws if not k.startswith("__preserve__/") and not k.endswith("_momentum")}
    assert {d for d, _, _ in loaded} == fed
    for name, value in params.items():
        if name in fed:
            assert np.array_equal(value.numpy(), ws[name]), name
        else:
            assert float(value.abs().sum()) == 0.0, name                  # not in the file: initialisation untouched
    targets = m._param_targets()
    for name in ("fc7_w", "_[noisy]_fc7_w"):
        assert np.array_equal(targets[name][1].numpy(), ws[name + "_momentum"]), name
    assert float(targets["fc6_w"][1].abs().sum()) == 0.0
    # the explicitly stored noisy blob wins over the rule; the clean one fills the noisy fc6 / fc7 weights
    assert ("_[noisy]_fc7_b", "_[noisy]_fc7_b", False) in loaded and ("_[noisy]_fc6_w", "fc6_w", False) in loaded
    with pytest.raises(RuntimeError, match="does not match"):
        m.initialize_from_weights({"fc7_w": np.zeros((3, 3), np.float32)})

    # save_model_to_weights_file's dictionary (utils/net_wsl.py:140-181) round-trips through the loader
    blobs = m.weights_file_blobs()
    assert set(blobs) == set(params) | {p + "_momentum" for p in params}
    m2 = WeblyHeadModel(C + 1, Cc, 7, hidden, noise=True, dtype=torch.float32, device="cpu")
    m2.initialize_from_weights({"blobs": blobs, "cfg": "unused"})
    assert torch.equal(m2.flat_param, m.flat_param) and torch.equal(m2.flat_mom, m.flat_mom)


def test_lr_change_scales_the_update_history_like_the_reference(gold):
    """heads.lr_change_correction / WeblyHeadModel.UpdateWorkspaceLr against DetectionModelHelper.UpdateWorkspaceLr,
    _SetNewLr and _CorrectMomentum (modeling/detector.py:509-586) run on a dictionary workspace: the `lr` blob after each
    call and the factor the momentum blobs were scaled by (0.1 at the schedule's 1e-3 -> 1e-4 step; none for a change
    below 10 %, from a rate <= 1e-7, or on the very first call from the blob's initial 0)."""
    from nafwebsod_b200.heads import lr_change_correction
    cur = np.float32(0.0)
    for new, blob, factor in zip(gold["lrseq_new"], gold["lrseq_blob"], gold["lrseq_momentum_factor"]):
        got = lr_change_correction(cur, new)
        assert np.float32(got) == factor, (cur, new, got, factor)
        cur = new
        assert cur == blob
    assert lr_change_correction(np.float32(1e-3), np.float32(1e-4), scale_momentum=False) == 1.0
