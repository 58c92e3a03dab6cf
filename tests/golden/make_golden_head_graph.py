"""Generate tests/golden/head_graph.npz: the forward blobs of the head AS THE REFERENCE'S OWN GRAPH BUILDERS WIRE THEM
(SURVEY.md section 8 rows a3-a8).  Run in the BUILD container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden_head_graph.py

The reference describes the head as a Caffe2 graph: `add_VGG16_roi_2fc_noise_head` (modeling/webly_heads.py:463-502) over
`add_VGG16_roi_2fc_head` (modeling/wsl_heads.py:654-681) and `DetectionModelHelper.RoIFeatureTransform`
(modeling/detector.py:268-331), `add_webly_outputs` (webly_heads.py:32-74) over `add_wsl_outputs` (wsl_heads.py:23-56),
`add_webly_losses` (webly_heads.py:123-197) with `add_cls_pred` (wsl_heads.py:213-227), `add_spatial_entropy_weight`
(webly_heads.py:265-391) and `add_cross_entropy_loss` (wsl_heads.py:292-302).  Caffe2 itself cannot be installed here, so
those functions are imported UNMODIFIED and run against an EAGER model helper: every `model.net.<Op>(inputs, outputs, **args)`
they emit is executed at once on a NumPy workspace.  Which operator runs on which blobs, in which order, with which axes /
flags / broadcast arguments -- everything a re-reading of the builders could get wrong -- therefore comes from the
reference's code.  What the interpreter supplies is the arithmetic of each operator:
  * the operators that live in the reference tree run the reference's own code: RoIFeatureBoost and
    [Weighted]CrossEntropyWithLogits are the unmodified CPU operators of oracle/_ref/libnawsod_ref.so, RoIIoU is the
    unmodified CUDA kernel run on the host (oracle/_ref/libnawsod_ref_kernels.so), RoIPoolF is the C restatement that is
    pinned bit for bit to the in-tree pooling kernel (tests/test_ref_kernels.py);
  * the Caffe2 built-ins (FC, Relu, Dropout, Softmax, Transpose, Add, Sub, Mul, Div, ReduceSum, Log, Scale, ReplaceNaN,
    MatMul, LeakyRelu, Shape, Cast, Clip, ConstantFill, StopGradient, AveragedLoss, Split, Concat) are float32 NumPy with
    the documented pytorch v1.3.0 defaults (LeakyRelu alpha 0.01, ReplaceNaN value 0, Dropout scale 1/(1-ratio),
    max-subtracted Softmax, NumPy-style broadcasting) -- the one part that stays a restatement.
The operator trace (type + blob names, in emission order) is stored next to the blobs, so the test can also show that the
oracle covers every operator the builders emit.

The same file holds three more reference-run cases around the head's parameters: `opt_*` -- add_single_gpu_param_update_ops
(modeling/optimizer_wsl.py:75-137) on the eager helper, the update net run three times through the reference's CPU operator;
`winit_*` -- initialize_gpu_from_weights_file (utils/net_wsl.py:53-137) on a dictionary workspace; `lrseq_*` --
UpdateWorkspaceLr / _SetNewLr / _CorrectMomentum (modeling/detector.py:509-586) over a sequence of learning rates.
Inputs that are cheap to regenerate are not stored: `load_case()` rebuilds a case's parameters from its seed (checksums in
the file guard the regeneration) and unpacks the bit-packed dropout masks.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
F32 = np.float32


def _load_roi_data_maker():
    spec = importlib.util.spec_from_file_location("make_golden_roi_data", os.path.join(HERE, "make_golden_roi_data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def head_case_params(ncls, Cc, hidden, seed):
    """The case's parameters, regenerated from its seed instead of being stored (they dominate the file size): the oracle's
    synthetic initialisation (gauss 0.01 / Xavier, zero biases), the narrow layers scaled so that activations keep the
    statistics of the 25088- and 4096-wide originals.  Keys use the oracle's `noisy_` prefix."""
    from oracle import nawsod_oracle as O
    params = O.synth_params(ncls - 1, Cc * 49, hidden, noise=True, seed=seed + 2)
    for k in list(params):
        if k.endswith("fc6_w"):
            params[k] = (params[k] * F32(np.sqrt(25088.0 / (Cc * 49)))).astype(F32)
        if k.endswith("fc7_w"):
            params[k] = (params[k] * F32(np.sqrt(4096.0 / hidden))).astype(F32)
    return params


def param_checksums(params):
    return np.array([np.asarray(params[k], np.float64).sum() for k in sorted(params)], np.float64)


def load_case(gold, i):
    """(X, rois, obn, labels, params, masks-or-None, cfg dict) of golden case i -- what both the CPU and the GPU test feed.
    Needs the oracle package only (not /root/reference)."""
    ncls, hidden, soft, train, Cc, seed = (int(v) for v in gold["case%d_cfg" % i])
    pre = "case%d_in_" % i
    params = head_case_params(ncls, Cc, hidden, seed)
    if not np.array_equal(param_checksums(params), gold["case%d_param_checksums" % i]):
        raise RuntimeError("the regenerated parameters of golden case %d differ from the ones the vectors were made with" % i)
    R = gold[pre + "rois"].shape[0]
    masks = None
    if train:
        masks = {}
        for k in gold.files:
            if k.startswith(pre + "maskbits_"):
                name = k[len(pre) + 9:].replace("_[noisy]_", "noisy_")
                masks[name] = np.unpackbits(gold[k])[:R * hidden].reshape(R, hidden).astype(F32)
    cfg = dict(ncls=ncls, hidden=hidden, soft=bool(soft), train=bool(train), Cc=Cc, seed=seed)
    return gold[pre + "X"], gold[pre + "rois"], gold[pre + "obn"], gold[pre + "labels"], params, masks, cfg


class EagerNet:
    """`model.net` / `model.param_init_net`: attribute access yields an operator that executes immediately."""

    def __init__(self, model):
        self._m = model

    def __getattr__(self, op_type):
        if op_type.startswith("_"):
            raise AttributeError(op_type)
        return lambda inputs, outputs=None, **args: self._m.run_op(op_type, inputs, outputs, **args)

    def Proto(self):
        raise RuntimeError("the eager net has no NetDef")


class EagerModel:
    """Stands for detectron.modeling.detector.DetectionModelHelper (a caffe2 CNNModelHelper): the builder-facing surface
    the head uses, executing on `self.ws` (blob name -> float32 / int array)."""

    def __init__(self, num_classes, train, ws, dropout_masks=None):
        from oracle import c_oracle, ref_kernels, ref_ops
        self.num_classes, self.train, self.ws = num_classes, train, ws
        self.net = EagerNet(self)
        self.param_init_net = self.net
        self.trace = []
        self.losses, self.metrics = [], []
        self.masks = dropout_masks or {}
        self.update_ops, self._last_ins = [], []
        self.biases, self.gn_params, self.param_to_grad, self.weights = [], [], {}, []
        self._co, self._rk, self._ro = c_oracle, ref_kernels, ref_ops

    # ---- CNNModelHelper / DetectionModelHelper helpers the builders call (thin forwards to net ops, as in caffe2's brew) ----
    def FC(self, blob_in, blob_out, dim_in, dim_out, **kw):
        W, b = self.ws[blob_out + "_w"], self.ws[blob_out + "_b"]
        assert W.shape == (dim_out, dim_in) and b.shape == (dim_out,), (blob_out, W.shape, dim_out, dim_in)
        return self.run_op("FC", [blob_in, blob_out + "_w", blob_out + "_b"], blob_out)

    def Relu(self, i, o, **kw):
        return self.run_op("Relu", i, o, **kw)

    def Dropout(self, i, o, **kw):
        return self.run_op("Dropout", i, [o, "_" + str(o) + "_mask"], **kw)[0]       # brew.dropout returns the data output

    def Softmax(self, i, o, **kw):
        return self.run_op("Softmax", i, o, **kw)

    def Transpose(self, i, o, **kw):
        return self.run_op("Transpose", i, o, **kw)

    def StopGradient(self, i, o):
        return self.run_op("StopGradient", i, o)

    def Accuracy(self, i, o, **kw):
        return self.run_op("Accuracy", i, o, **kw)

    def AddLosses(self, losses):
        self.losses += [losses] if isinstance(losses, str) else list(losses)

    def AddMetrics(self, metrics):
        self.metrics += [metrics] if isinstance(metrics, str) else list(metrics)

    def TrainableParams(self, gpu_id=-1):
        return list(self.weights) + list(self.biases)          # cnn.CNNModelHelper.TrainableParams: weights, then biases

    # ---- the interpreter ----
    def run_op(self, op_type, inputs, outputs=None, **args):
        ins = [inputs] if isinstance(inputs, str) else [str(i) for i in inputs]
        if outputs is None:
            raise RuntimeError("%s: the builders always name their outputs" % op_type)
        outs = [outputs] if isinstance(outputs, str) else [str(o) for o in outputs]
        self.trace.append("%s(%s)->(%s)%s" % (op_type, ",".join(ins), ",".join(outs),
                                               "" if not args else " " + repr(sorted((k, str(v)) for k, v in args.items()))))
        x = [self.ws[i] for i in ins if i in self.ws] if op_type == "Accuracy" else [self.ws[i] for i in ins]
        self._last_ins = ins
        res = getattr(self, "op_" + op_type)(x, args, outs)
        res = res if isinstance(res, (list, tuple)) else [res]
        for name, val in zip(outs, res):
            self.ws[name] = val
        return outs[0] if len(outs) == 1 else tuple(outs)     # core.Net._CreateAndAddToSelf: one output -> one BlobReference

    # reference-tree operators: the reference's own code
    def op_RoIPoolF(self, x, a, outs):
        assert a["pooled_w"] == a["pooled_h"] == 7
        Y, A = self._co.roi_pool_f(x[0], x[1], float(a["spatial_scale"]))
        return [Y, A]

    def op_RoIFeatureBoost(self, x, a, outs):
        return self._ro.RefOp("RoIFeatureBoost").run([x[0], x[1]], 1)[0].reshape(x[0].shape)

    def op_RoIIoU(self, x, a, outs):
        return self._rk.roi_iou(x[0])

    def op_WeightedCrossEntropyWithLogits(self, x, a, outs):
        return self._ro.RefOp("WeightedCrossEntropyWithLogits", is_mean=float(bool(a.get("is_mean", False)))).run(x, 1, out_cap=4)[0]

    def op_CrossEntropyWithLogits(self, x, a, outs):
        return self._ro.RefOp("CrossEntropyWithLogits", is_mean=float(bool(a.get("is_mean", False)))).run(x, 1, out_cap=4)[0]

    # Caffe2 built-ins (pytorch v1.3.0 caffe2/operators/*), float32 NumPy
    def op_FC(self, x, a, outs):
        return (x[0].reshape(x[0].shape[0], -1) @ x[1].T + x[2]).astype(F32)

    def op_Relu(self, x, a, outs):
        return np.maximum(x[0], F32(0))

    def op_Dropout(self, x, a, outs):
        assert not a.get("is_test", False)
        ratio = F32(a["ratio"])
        mask = self.masks[outs[0]].astype(F32)
        return [(x[0] * mask * (F32(1) / (F32(1) - ratio))).astype(F32), mask]

    def op_Softmax(self, x, a, outs):
        assert a.get("axis", 1) == 1 and x[0].ndim == 2
        e = np.exp(x[0] - x[0].max(axis=1, keepdims=True), dtype=F32)
        return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)

    def op_Transpose(self, x, a, outs):
        return np.ascontiguousarray(np.transpose(x[0], a["axes"]))

    def op_Add(self, x, a, outs):
        return (x[0] + x[1]).astype(F32)

    def op_Sub(self, x, a, outs):
        return (x[0] - x[1]).astype(F32)

    def op_Mul(self, x, a, outs):
        return (x[0] * x[1]).astype(F32)

    def op_Div(self, x, a, outs):
        with np.errstate(divide="ignore", invalid="ignore"):
            return (x[0] / x[1]).astype(F32)

    def op_ReduceSum(self, x, a, outs):
        return x[0].sum(axis=tuple(a["axes"]), keepdims=bool(a.get("keepdims", True)), dtype=F32)

    def op_Log(self, x, a, outs):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.log(x[0]).astype(F32)

    def op_Scale(self, x, a, outs):
        return (x[0] * F32(a.get("scale", 1.0))).astype(F32)

    def op_ReplaceNaN(self, x, a, outs):
        return np.where(np.isnan(x[0]), F32(a.get("value", 0.0)), x[0]).astype(F32)

    def op_MatMul(self, x, a, outs):
        assert not a.get("trans_a") and not a.get("trans_b")
        return (x[0] @ x[1]).astype(F32)

    def op_LeakyRelu(self, x, a, outs):
        return np.where(x[0] > 0, x[0], x[0] * F32(a.get("alpha", 0.01))).astype(F32)

    def op_Shape(self, x, a, outs):
        shp = np.array(x[0].shape, np.int64)
        return shp[list(a["axes"])] if "axes" in a else shp

    def op_Cast(self, x, a, outs):
        assert a["to"] == 1                              # caffe2_pb2.TensorProto.FLOAT
        return x[0].astype(F32)

    def op_Clip(self, x, a, outs):
        return np.clip(x[0], F32(a["min"]), F32(a["max"])).astype(F32)

    def op_ConstantFill(self, x, a, outs):
        shape = tuple(a["shape"]) if not x else x[0].shape             # ConstantFill([], shape=...) or ConstantFill([like])
        return np.full(shape, F32(a.get("value", 0.0)), F32)

    def op_ACMWeightDecayMomentumSGDUpdate(self, x, a, outs):
        """The reference's own CPU operator (one instance per emitted op, like a Caffe2 net: its hidden iter_count_ lives
        across runs).  The instance and its blob names are kept so that `run_updates()` can run the update net again."""
        op = self._ro.RefOp("ACMWeightDecayMomentumSGDUpdate", **{k: float(v) for k, v in a.items()})
        self.update_ops.append((op, list(self._last_ins), list(outs), dict(a)))
        return op.run(x, 4, out_alias=[0, 1, 3, 4])

    def run_updates(self):
        for op, ins, outs, _ in self.update_ops:
            res = op.run([self.ws[i] for i in ins], 4, out_alias=[0, 1, 3, 4])
            for name, val in zip(outs, res):
                self.ws[name] = val

    def op_StopGradient(self, x, a, outs):
        return x[0]

    def op_AveragedLoss(self, x, a, outs):
        return np.asarray(x[0], F32).mean(dtype=F32).reshape(())

    def op_Accuracy(self, x, a, outs):
        return np.zeros((), F32)                         # metric only; needs labels_int32, which the head's parity does not feed

    def op_Stat(self, x, a, outs):
        return [np.zeros((), F32) for _ in outs]         # logging only (detectron/ops/stat_op.*): prints running means every `display` calls

    def op_Split(self, x, a, outs):
        assert a["axis"] == 1
        return np.split(x[0], np.cumsum(a["split"])[:-1], axis=1)

    def op_Concat(self, x, a, outs):
        assert a["axis"] == 1
        return [np.concatenate(x, axis=1), np.array([t.shape[1] for t in x], np.int32)]


def reference_roi_feature_transform():
    """detector.DetectionModelHelper.RoIFeatureTransform as a plain function (the class itself derives from a caffe2
    class that does not exist here; only this one method is needed and it only touches `self.net`)."""
    import ast
    import inspect
    import textwrap
    src = open("/root/reference/detectron/modeling/detector.py").read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "RoIFeatureTransform":
            lines = src.split("\n")[node.lineno - 1: node.end_lineno]
            code = textwrap.dedent("\n".join(lines))
            from detectron.core.config import cfg
            ns = {"cfg": cfg}
            exec(compile(code, "/root/reference/detectron/modeling/detector.py::RoIFeatureTransform", "exec"), ns)
            assert "self.net.__getattr__(method)" in code and inspect.isfunction(ns["RoIFeatureTransform"])
            return ns["RoIFeatureTransform"]
    raise RuntimeError("RoIFeatureTransform not found in the reference")


def build_case(rng, mods, cfg, *, Cc, Hc, Wc, R, hidden, ncls, soft, train, seed):
    """One head problem of reduced width (the builders hard-code 4096 hidden units; see `hidden` below), run through the
    reference's builders.  Returns (inputs, outputs, trace)."""
    from oracle import nawsod_oracle as O
    webly, wsl_heads, xform = mods
    C = ncls - 1
    X = O.synth_conv5(1, Cc, Hc, Wc, seed=seed)
    rois = O.synth_rois(R, Hc * 16, Wc * 16, seed=seed + 1)
    obn = (rng.random((R, 1)) + 1).astype(F32)
    labels = np.zeros((1, C), F32)
    labels[0, rng.integers(0, C)] = 1
    if soft:                                                    # bagging-mixup: (lam, 1 - lam) on two classes
        lam = F32(rng.beta(1.5, 1.5))
        a, b = rng.choice(C, 2, replace=False)
        labels[:] = 0
        labels[0, a], labels[0, b] = lam, F32(1) - lam
    params = head_case_params(ncls, Cc, hidden, seed)
    ws = {"conv5_3": X, "rois": rois, "obn_scores": obn, "labels_oh": labels}
    for k, v in params.items():                                 # the oracle's 'noisy_' prefix -> the reference's blob names
        name = k
        if k.startswith("noisy_fc6") or k.startswith("noisy_fc7"):
            name = "_[noisy]_" + k[len("noisy_"):]
        ws[name] = v
    masks = {}
    if train:
        for nme in ("drop6", "drop7", "_[noisy]_drop6", "_[noisy]_drop7"):
            masks[nme] = (rng.random((R, hidden)) < 0.5).astype(F32)
    model = EagerModel(ncls, train, ws, masks)
    model.RoIFeatureTransform = lambda *a, **k: xform(model, *a, **k)

    # The builders hard-code 4096 hidden units (wsl_heads.py:674-678, webly_heads.py:490-498).  The golden case keeps
    # the test small by running them at `hidden` units: FC() checks each weight against the dims it is GIVEN, so the
    # literal 4096 is mapped onto the width of the weights that were fed -- the wiring is untouched.
    real_fc = model.FC
    model.FC = lambda bi, bo, di, do, **kw: real_fc(bi, bo, hidden if di == 4096 else di, hidden if do == 4096 else do, **kw)

    blobs, dims = webly.add_VGG16_roi_2fc_noise_head(model, "conv5_3", Cc, 1.0 / 16)
    assert dims == [4096, 4096]
    webly.add_webly_outputs(model, blobs, dims)
    if train:
        webly.add_webly_losses(model)
    inputs = dict(X=X, rois=rois, obn=obn, labels=labels)
    inputs.update({"maskbits_" + k: np.packbits(v.astype(np.uint8)) for k, v in masks.items()})
    inputs["__param_checksums"] = param_checksums(params)
    keep = ["roi_feat", "fc6", "fc7", "_[noisy]_fc6", "_[noisy]_fc7", "fc8c", "fc8d", "noisy_fc8c", "noisy_fc8d", "rois_pred",
            "rois_pred_noise", "cls_prob", "cls_prob_noise", "rois_J", "rois_pred_E", "rois_pred_D", "rois_pred_hatE_sum",
            "rois_pred_hatE_sum_norm", "rois_class_weight", "rois_class_weight_noise", "cross_entropy", "cross_entropy_noise",
            "loss_cls", "loss_cls_noise", "loss_cls_grad", "loss_cls_noise_grad", "drop7", "_[noisy]_drop7"]
    outputs = {k: np.asarray(ws[k]) for k in keep if k in ws}
    return inputs, outputs, model.trace, model.losses


def build_optimizer_case(rng, cfg):
    """add_single_gpu_param_update_ops (modeling/optimizer_wsl.py:75-137) over the head's parameter blobs: which blob gets
    which weight decay / lr multiplier / gpu_num / iter_size comes from the reference; the update itself is the reference's
    CPU operator.  Three runs of the update net with fresh gradients (lr 1e-3, 1e-3, 1e-4)."""
    import detectron.modeling.optimizer_wsl as opt
    weights = ["fc6_w", "fc7_w", "_[noisy]_fc6_w", "_[noisy]_fc7_w", "fc8c_w", "fc8d_w", "noisy_fc8c_w", "noisy_fc8d_w"]
    biases = [w[:-2] + "_b" for w in weights]
    ws = {}
    for k, name in enumerate(weights + biases):
        ws[name] = rng.standard_normal(37 + 3 * k).astype(F32)
    model = EagerModel(21, True, ws)
    model.weights, model.biases = weights, biases
    model.param_to_grad = {p: p + "_grad" for p in weights + biases}
    steps, G = 3, {}
    for p in weights + biases:
        G[p] = rng.standard_normal((steps, ws[p].size)).astype(F32)
        ws[p + "_grad"] = G[p][0].copy()
    p0 = {p: ws[p].copy() for p in weights + biases}
    P, M = {p: [] for p in p0}, {p: [] for p in p0}
    lrs = [1e-3, 1e-3, 1e-4]
    for s in range(steps):
        if s == 0:
            opt.add_single_gpu_param_update_ops(model, 0)        # builds AND (eagerly) runs the update net once, with lr = 0 ...
            # ... so rewind: the dummy lr of the builder is "set properly at the start of training" (optimizer_wsl.py:81-85)
            for p in p0:
                ws[p] = p0[p].copy()
            model.update_ops_args = [(o[1], o[2], o[3]) for o in model.update_ops]
            # fresh operator instances (iter_count_ = 0) with exactly the blobs / arguments the builder emitted
            model.update_ops = [(model._ro.RefOp("ACMWeightDecayMomentumSGDUpdate", **{k: float(v) for k, v in a.items()}), i, o, a)
                                for (_, i, o, a) in model.update_ops]
        for p in p0:
            ws[p + "_grad"] = G[p][s].copy()
        ws["lr"] = np.array([lrs[s]], F32)
        model.run_updates()
        for p in p0:
            P[p].append(ws[p].copy()); M[p].append(ws[p + "_momentum"].copy())
    out = {"opt_params": np.array(weights + biases), "opt_lrs": np.array(lrs, F32), "opt_trace": np.array(model.trace),
           "opt_gpu_num": np.int32(cfg.NUM_GPUS), "opt_iter_size": np.int32(cfg.WSL.ITER_SIZE)}
    for p in p0:
        out["opt_p0_" + p], out["opt_G_" + p] = p0[p], G[p]
        out["opt_P_" + p], out["opt_M_" + p] = np.stack(P[p]), np.stack(M[p])
    args = {}
    for ins, outs, a in model.update_ops_args:
        args[ins[3]] = a                                         # keyed by the parameter blob
        assert ins == [ins[3] + "_grad", ins[3] + "_momentum", "lr", ins[3], ins[3] + "_acmgrad"], ins
        assert outs == [ins[0], ins[1], ins[3], ins[4]], outs
    for p in p0:
        out["opt_args_" + p] = np.array([args[p]["momentum"], args[p]["iter_size"], args[p]["gpu_num"], args[p]["lr_mult"],
                                         args[p]["weight_decay"]], np.float64)
    return out


def build_weights_init_case(rng):
    """initialize_gpu_from_weights_file (utils/net_wsl.py:53-137) run unmodified against a dictionary workspace: which
    source blob initialises which parameter blob (the `]_` rule for `_[noisy]_fc6/fc7`, missing blobs left alone, momentum
    loaded along, unused blobs preserved)."""
    import contextlib
    import detectron.utils.net_wsl as nu
    params = ["fc6_w", "fc6_b", "fc7_w", "fc7_b", "_[noisy]_fc6_w", "_[noisy]_fc6_b", "_[noisy]_fc7_w", "_[noisy]_fc7_b",
              "fc8c_w", "fc8c_b", "fc8d_w", "fc8d_b", "noisy_fc8c_w", "noisy_fc8c_b", "noisy_fc8d_w", "noisy_fc8d_b"]
    Cc, hidden, C = 4, 8, 5
    shapes = {"fc6_w": (hidden, Cc * 49), "fc6_b": (hidden,), "fc7_w": (hidden, hidden), "fc7_b": (hidden,),
              "fc8c_w": (C, hidden), "fc8c_b": (C,), "fc8d_w": (C, hidden), "fc8d_b": (C,)}
    # an ImageNet-style file: the clean fc6 / fc7 only (+ a momentum blob, + a conv blob this model does not use), and one
    # explicitly stored noisy blob that must win over the `]_` rule
    src = {k: rng.standard_normal(shapes[k]).astype(F32) for k in ("fc6_w", "fc6_b", "fc7_w", "fc7_b")}
    src["fc7_w_momentum"] = rng.standard_normal(shapes["fc7_w"]).astype(F32)
    src["_[noisy]_fc7_b"] = rng.standard_normal(shapes["fc7_b"]).astype(F32)
    src["conv1_1_w"] = rng.standard_normal((3,)).astype(F32)
    ws = {}

    class Model:
        pass
    model = Model()
    model.params = list(params)
    model.GetComputedParams = lambda: []
    nu.load_object = lambda f: {"blobs": dict(src)}
    nu.workspace.Blobs = lambda: list(ws)
    nu.workspace.FeedBlob = lambda name, value: ws.__setitem__(str(name), np.array(value))
    nu.workspace.FetchBlob = lambda name: ws[str(name)]
    nu.core.ScopedName = lambda s: s
    nu.c2_utils.UnscopeName = lambda s: s
    nu.c2_utils.NamedCudaScope = lambda gpu_id: contextlib.nullcontext()
    nu.c2_utils.CpuScope = lambda: contextlib.nullcontext()
    nu.initialize_gpu_from_weights_file(model, "weights.pkl", gpu_id=0)
    out = {"winit_params": np.array(params), "winit_cfg": np.array([Cc, hidden, C], np.int32),
           "winit_ws_names": np.array(sorted(ws))}
    for k, v in src.items():
        out["winit_src_" + k] = v
    for k, v in ws.items():
        out["winit_ws_" + k] = v
    return out


def build_lr_change_case(cfg):
    """DetectionModelHelper.UpdateWorkspaceLr / _SetNewLr / _CorrectMomentum and _get_lr_change_ratio
    (modeling/detector.py:509-586) as plain functions on a dictionary workspace: for a sequence of learning rates, the value
    the `lr` blob takes and the factor every `<param>_momentum` blob is scaled by (or 1.0 when the reference leaves the
    update history alone)."""
    import ast
    import contextlib
    import logging
    import textwrap
    import types
    src = open("/root/reference/detectron/modeling/detector.py").read()
    tree = ast.parse(src)
    lines = src.split("\n")
    code = []
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("UpdateWorkspaceLr", "_SetNewLr", "_CorrectMomentum", "_get_lr_change_ratio"):
            code.append(textwrap.dedent("\n".join(lines[node.lineno - 1: node.end_lineno])))
    assert len(code) == 4
    ws = {"gpu_0/lr": np.array([0.0], F32), "gpu_0/fc6_w_momentum": np.ones(3, F32)}
    scales = []

    def run_once(op):
        scales.append(op["scale"])
        ws[op["out"]] = (ws[op["in"]] * F32(op["scale"])).astype(F32)      # caffe2 Scale: float argument, float32 arithmetic
    one_gpu = types.SimpleNamespace(NUM_GPUS=1, SOLVER=cfg.SOLVER)          # the blobs of gpu_0 stand for every replica
    ns = {"np": np, "cfg": one_gpu, "logger": logging.getLogger("lr"),
          "workspace": types.SimpleNamespace(FetchBlob=lambda n: ws[n], FeedBlob=lambda n, v: ws.__setitem__(n, np.array(v)),
                                             RunOperatorOnce=run_once),
          "core": types.SimpleNamespace(CreateOperator=lambda typ, i, o, scale: {"type": typ, "in": i[0], "out": o[0], "scale": scale}),
          "c2_utils": types.SimpleNamespace(CudaScope=lambda i: contextlib.nullcontext())}
    for c in code:
        exec(compile(c, "/root/reference/detectron/modeling/detector.py", "exec"), ns)
    helper = types.SimpleNamespace(TrainableParams=lambda gpu_id=-1: ["gpu_0/fc6_w"])
    helper._SetNewLr = lambda cur, new: ns["_SetNewLr"](helper, cur, new)
    helper._CorrectMomentum = lambda corr: ns["_CorrectMomentum"](helper, corr)
    seq = [F32(1e-3), F32(1e-3), F32(1e-4), F32(1.05e-4), F32(5e-8), F32(1e-3), F32(2e-3), F32(1e-5)]
    lr_after, factor = [], []
    for it, new_lr in enumerate(seq):
        before = ws["gpu_0/fc6_w_momentum"].copy()
        n_scales = len(scales)
        ns["UpdateWorkspaceLr"](helper, it, new_lr)
        lr_after.append(ws["gpu_0/lr"][0])
        factor.append(F32(scales[-1]) if len(scales) > n_scales else F32(1.0))
        assert np.array_equal(ws["gpu_0/fc6_w_momentum"], (before * factor[-1]).astype(F32))
    assert cfg.SOLVER.SCALE_MOMENTUM and cfg.SOLVER.SCALE_MOMENTUM_THRESHOLD == 1.1
    return {"lrseq_new": np.array(seq, F32), "lrseq_blob": np.array(lr_after, F32), "lrseq_momentum_factor": np.array(factor, F32)}


def main():
    maker = _load_roi_data_maker()
    sys.meta_path.insert(0, maker._Absent())
    import future.utils
    future.utils.iteritems = lambda d: iter(d.items())
    sys.path.insert(0, "/root/reference")
    from detectron.core.config import cfg, merge_cfg_from_file
    import detectron.modeling.webly_heads as webly
    import detectron.modeling.wsl_heads as wsl_heads
    import yaml
    import detectron.utils.env as envu
    envu.yaml_load = lambda f: yaml.load(f, Loader=yaml.SafeLoader)     # PyYAML >= 6 needs the Loader the reference omits
    merge_cfg_from_file("/root/reference/configs/flickr_voc/na_wsddn_V-16-C5_1x.yaml")      # the shipped NA-fWebSOD config
    assert cfg.WEBLY.ENTROPY and cfg.WSL.MEAN_LOSS and cfg.FAST_RCNN.ROI_XFORM_METHOD == "RoIPoolF"
    assert cfg.FAST_RCNN.ROI_XFORM_RESOLUTION == 7 and cfg.TRAIN.FREEZE_CONV_BODY and cfg.TRAIN.IMS_PER_BATCH == 1
    mods = (webly, wsl_heads, reference_roi_feature_transform())
    rng = np.random.default_rng(2024)
    out = {}
    cases = [dict(Cc=16, Hc=12, Wc=16, R=96, hidden=256, ncls=21, soft=False, train=True, seed=11),
             dict(Cc=8, Hc=10, Wc=14, R=130, hidden=32, ncls=81, soft=True, train=True, seed=21),
             dict(Cc=16, Hc=12, Wc=16, R=64, hidden=64, ncls=21, soft=False, train=False, seed=31)]
    for i, c in enumerate(cases):
        inputs, outputs, trace, losses = build_case(rng, mods, cfg, **c)
        pre = "case%d_" % i
        out[pre + "param_checksums"] = inputs.pop("__param_checksums")
        for k, v in inputs.items():
            out[pre + "in_" + k] = v
        for k, v in outputs.items():
            out[pre + "out_" + k] = v
        out[pre + "trace"] = np.array(trace)
        out[pre + "losses"] = np.array(losses)
        out[pre + "cfg"] = np.array([c["ncls"], c["hidden"], int(c["soft"]), int(c["train"]), c["Cc"], c["seed"]], np.int32)
        print("case", i, c, "->", len(trace), "operators;", {k: v.shape for k, v in outputs.items() if k in ("rois_pred", "cls_prob", "loss_cls")})
    out["cases"] = np.int32(len(cases))
    out.update(build_optimizer_case(rng, cfg))
    out.update(build_weights_init_case(rng))
    out.update(build_lr_change_case(cfg))
    np.savez_compressed(os.path.join(HERE, "head_graph.npz"), **out)
    print("wrote head_graph.npz (%d arrays)" % len(out))


if __name__ == "__main__":
    main()
