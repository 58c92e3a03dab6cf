"""CPU oracle for the NA-fWebSOD per-proposal head (TEST INFRASTRUCTURE ONLY).

This module is a NumPy restatement of the reference's algorithm for the hot path
(SURVEY.md section 8a).  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``na-fwebsod_b200/`` imports it, and the product
path raises if the CUDA library is missing -- there is no CPU fallback.

Parity status
-------------
* RoIFeatureBoost, (Weighted)CrossEntropyWithLogits fwd/bwd and
  ACMWeightDecayMomentumSGDUpdate are PINNED against the reference's own C++
  CPU operators compiled from ``/root/reference`` (``oracle/build_ref.sh`` ->
  ``oracle/_ref/libnawsod_ref.so``; vectors in ``tests/golden/ref_ops.npz``).
* RoIPoolF is a Caffe2 built-in whose source is not in the reference tree
  (pytorch v1.3.0 ``modules/detectron/roi_pool_f_op.cu``); the arithmetic is
  restated from the in-tree clone ``detectron/ops/roi_loop_pool_op.cu:19-140``
  with the three deltas of SURVEY.md row a1 and cross-checked bit-exactly
  (values and int32 argmax) against ``torch.ops.torchvision.roi_pool`` (CPU),
  which descends from the same Caffe2 kernel (``tests/golden/roi_pool.npz``).
* RoIIoU is GPU-only in the reference and the head graph is built from Caffe2
  built-ins: both are restated here and anchored by torch-autograd fp64
  cross-checks (``tests/golden/head_small.npz``) -> "parity unpinned" by the
  reference's own tests for those rows (the reference has no tests on this path).

All arithmetic is float32 unless noted; every function cites the reference
file:line (relative to /root/reference/detectron) it follows.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
FLT_MAX = np.finfo(np.float32).max
LOG_THRESHOLD = F32(1e-20)   # ops/cross_entropy_wsl_op.h:56
DIFF_THRESHOLD = F32(1e4)    # ops/cross_entropy_wsl_op.h:74


# --------------------------------------------------------------------------- #
# a1/a2  RoIPoolF / RoIPoolFGradient
# --------------------------------------------------------------------------- #
def _round_half_away(x):
    """C roundf(): half away from zero, on float32 (ops/roi_loop_pool_op.cu:42-45)."""
    x = np.asarray(x, dtype=F32)
    return (np.sign(x) * np.floor(np.abs(x) + F32(0.5))).astype(np.int32)


def roi_bin_bounds(roi, spatial_scale, height, width, pooled_h=7, pooled_w=7):
    """Integer bin bounds of one RoI (ops/roi_loop_pool_op.cu:41-68, stride-5 rois).

    Returns (batch_ind, hstart[ph], hend[ph], wstart[pw], wend[pw]) already shifted
    by the RoI origin and clipped to the map.
    """
    scale = F32(spatial_scale)
    b = int(roi[0])
    sw = int(_round_half_away(F32(roi[1]) * scale))
    sh = int(_round_half_away(F32(roi[2]) * scale))
    ew = int(_round_half_away(F32(roi[3]) * scale))
    eh = int(_round_half_away(F32(roi[4]) * scale))
    roi_w = max(ew - sw + 1, 1)            # :54
    roi_h = max(eh - sh + 1, 1)            # :55
    bin_h = F32(roi_h) / F32(pooled_h)     # :56 (fp32 division)
    bin_w = F32(roi_w) / F32(pooled_w)     # :57
    ph = np.arange(pooled_h, dtype=np.int32)
    pw = np.arange(pooled_w, dtype=np.int32)
    hstart = np.floor(ph.astype(F32) * bin_h).astype(np.int32)          # :59
    wstart = np.floor(pw.astype(F32) * bin_w).astype(np.int32)          # :60
    hend = np.ceil((ph + 1).astype(F32) * bin_h).astype(np.int32)       # :61
    wend = np.ceil((pw + 1).astype(F32) * bin_w).astype(np.int32)       # :62
    hstart = np.minimum(np.maximum(hstart + sh, 0), height)             # :65-68
    hend = np.minimum(np.maximum(hend + sh, 0), height)
    wstart = np.minimum(np.maximum(wstart + sw, 0), width)
    wend = np.minimum(np.maximum(wend + sw, 0), width)
    return b, hstart, hend, wstart, wend


def roi_pool_f(X, rois, spatial_scale, pooled_h=7, pooled_w=7):
    """Caffe2 ``RoIPoolF([X, rois] -> [Y, argmax])`` (modeling/detector.py:321-329).

    Arithmetic per ops/roi_loop_pool_op.cu:19-102 with RoIPoolF's deltas: rois stride 5,
    no inner rectangle, ``maxval = is_empty ? 0 : -FLT_MAX`` (the commented line :72).
    Scan order h then w with strict ``>`` -> first maximum in row-major order, which is
    what ``argmax`` over the row-major flattened window returns.  Inputs must be finite.
    X [N,C,H,W] f32, rois [R,5] f32 -> Y [R,C,ph,pw] f32, argmax [R,C,ph,pw] i32
    (index h*W+w inside the image plane, -1 for an empty bin).
    """
    X = np.ascontiguousarray(X, dtype=F32)
    rois = np.asarray(rois, dtype=F32).reshape(-1, 5)
    N, C, H, W = X.shape
    R = rois.shape[0]
    Y = np.zeros((R, C, pooled_h, pooled_w), dtype=F32)
    A = np.full((R, C, pooled_h, pooled_w), -1, dtype=np.int32)
    for r in range(R):
        b, hs, he, ws, we = roi_bin_bounds(rois[r], spatial_scale, H, W, pooled_h, pooled_w)
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                h0, h1, w0, w1 = int(hs[ph]), int(he[ph]), int(ws[pw]), int(we[pw])
                if h1 <= h0 or w1 <= w0:       # :69  empty -> 0 / -1
                    continue
                win = X[b, :, h0:h1, w0:w1].reshape(C, -1)
                k = np.argmax(win, axis=1)     # first max, row-major == strict '>' scan
                v = win[np.arange(C), k]
                ww = w1 - w0
                idx = (h0 + k // ww) * W + (w0 + k % ww)
                never = v <= -FLT_MAX          # a value equal to -FLT_MAX is never selected
                Y[r, :, ph, pw] = v
                A[r, :, ph, pw] = np.where(never, -1, idx)
    return Y, A


def roi_pool_f_grad(X_shape, rois, argmax, dY):
    """``RoIPoolFGradient([X, rois, argmax, dY] -> dX)`` (ops/roi_loop_pool_op.cu:105-140,
    zero fill :199-201): dX[b,c,argmax] += dY[r,c,ph,pw], skipping argmax == -1.
    Accumulated in float64 then rounded: the reference's atomicAdd order is unspecified."""
    N, C, H, W = X_shape
    rois = np.asarray(rois, dtype=F32).reshape(-1, 5)
    R = rois.shape[0]
    dX = np.zeros((N, C, H * W), dtype=np.float64)
    dYf = np.asarray(dY, dtype=F32).reshape(R, C, -1)
    Af = np.asarray(argmax).reshape(R, C, -1)
    cidx = np.arange(C)[:, None]
    for r in range(R):
        b = int(rois[r, 0])
        a = Af[r]
        m = a >= 0
        np.add.at(dX[b], (np.broadcast_to(cidx, a.shape)[m], a[m]), dYf[r][m].astype(np.float64))
    return dX.reshape(N, C, H, W).astype(F32)


# --------------------------------------------------------------------------- #
# a3  RoIFeatureBoost
# --------------------------------------------------------------------------- #
def roi_feature_boost(X, S):
    """``RoIFeatureBoost([X, S] -> Y)``: Y[r,:] = X[r,:] * S[r] (ops/roi_feature_boost_op.cc:8-35)."""
    X = np.asarray(X, dtype=F32)
    S = np.asarray(S, dtype=F32).reshape(-1)
    assert X.shape[0] == S.shape[0]
    return X * S.reshape((-1,) + (1,) * (X.ndim - 1))


def roi_feature_boost_grad(dY, S):
    """``RoIFeatureBoostGradient([dY, S] -> dX)`` (ops/roi_feature_boost_op.cc:37-64)."""
    return roi_feature_boost(dY, S)


# --------------------------------------------------------------------------- #
# a4  FC / Relu / Dropout (Caffe2 built-ins; wiring modeling/wsl_heads.py:654-681)
# --------------------------------------------------------------------------- #
def fc(X, W, b):
    """Caffe2 ``FC([X, W, b] -> Y)``: Y = X.reshape(R,-1) @ W.T + b, W is [out, in]."""
    X2 = np.asarray(X, dtype=F32).reshape(X.shape[0], -1)
    return (X2 @ np.asarray(W, dtype=F32).T + np.asarray(b, dtype=F32)[None, :]).astype(F32)


def fc_grad(X, W, dY):
    """Caffe2 ``FCGradient([X, W, dY] -> [dW, db, dX])``."""
    X2 = np.asarray(X, dtype=F32).reshape(X.shape[0], -1)
    dY = np.asarray(dY, dtype=F32)
    dW = (dY.T @ X2).astype(F32)
    db = dY.sum(axis=0, dtype=np.float64).astype(F32)
    dX = (dY @ np.asarray(W, dtype=F32)).astype(F32)
    return dW, db, dX


def relu(X):
    return np.maximum(np.asarray(X, dtype=F32), F32(0))


def relu_grad(Y, dY):
    """Caffe2 ReluGradient uses the output: dX = dY * (Y > 0)."""
    return np.where(np.asarray(Y) > 0, np.asarray(dY, dtype=F32), F32(0))


def dropout(X, mask, ratio=0.5):
    """Caffe2 ``Dropout(is_test=0)``: Y = X * mask / (1 - ratio); the mask is injected
    (modeling/wsl_heads.py:1259-1267; Caffe2's RNG stream is not reproducible)."""
    scale = F32(1.0) / (F32(1.0) - F32(ratio))
    return np.asarray(X, dtype=F32) * np.asarray(mask, dtype=F32) * scale


def dropout_grad(dY, mask, ratio=0.5):
    return dropout(dY, mask, ratio)


# --------------------------------------------------------------------------- #
# a5/a6  two-stream MIL outputs
# --------------------------------------------------------------------------- #
def softmax(x, axis):
    """Caffe2 Softmax: max-subtracted."""
    x = np.asarray(x, dtype=F32)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def softmax_grad(Y, dY, axis):
    """Caffe2 SoftmaxGradient: dX = Y * (dY - sum(dY*Y))."""
    s = (dY * Y).sum(axis=axis, keepdims=True, dtype=F32)
    return (Y * (dY - s)).astype(F32)


def wsl_outputs(fc8c, fc8d):
    """``add_wsl_outputs`` (modeling/wsl_heads.py:49-55): alpha_cls = softmax over classes,
    alpha_det = softmax over RoIs (Transpose/Softmax/Transpose), rois_pred = product."""
    a_cls = softmax(fc8c, axis=1)
    a_det = softmax(fc8d, axis=0)
    return a_cls, a_det, (a_cls * a_det).astype(F32)


def cls_pred(rois_pred):
    """``add_cls_pred`` (modeling/wsl_heads.py:213-227): ReduceSum(axes=[0], keepdims)."""
    return rois_pred.sum(axis=0, keepdims=True, dtype=F32).astype(F32)


def test_cls_prob(rois_pred):
    """Test-time ``cls_prob`` [R, C+1] (modeling/wsl_heads.py:57-67): column 0 repeated."""
    return np.concatenate([rois_pred[:, :1], rois_pred], axis=1)


# --------------------------------------------------------------------------- #
# a7  RoIIoU + noise-aware class weights
# --------------------------------------------------------------------------- #
def roi_iou(rois):
    """``RoIIoU([rois] -> J)`` (ops/roi_iou_op.cu:28-62): coords truncated to int, +1 widths,
    intersection w/h truncated to int, union evaluated in double then rounded to float,
    diagonal forced to 1."""
    r = np.asarray(rois, dtype=F32).reshape(-1, 5)
    x1 = r[:, 1].astype(np.int32).astype(np.int64)   # C float->int: truncation
    y1 = r[:, 2].astype(np.int32).astype(np.int64)
    x2 = r[:, 3].astype(np.int32).astype(np.int64)
    y2 = r[:, 4].astype(np.int32).astype(np.int64)
    xmin = np.maximum(x1[:, None], x1[None, :])
    ymin = np.maximum(y1[:, None], y1[None, :])
    xmax = np.minimum(x2[:, None], x2[None, :])
    ymax = np.minimum(y2[:, None], y2[None, :])
    w = np.maximum(xmax - xmin + 1, 0)
    h = np.maximum(ymax - ymin + 1, 0)
    inters = (w * h).astype(F32)                       # float inters = w * h
    area = ((x2 - x1 + 1) * (y2 - y1 + 1)).astype(np.float64)
    uni = (area[:, None] + area[None, :] - inters.astype(np.float64)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        J = (inters / uni).astype(F32)
    np.fill_diagonal(J, F32(1.0))
    return J


def spatial_entropy_weight(rois_pred, cls_prob, rois, labels_oh, return_parts=False):
    """``add_spatial_entropy_weight`` (modeling/webly_heads.py:265-391, live ``else`` branch
    :334-347).  Forward only: both outputs are StopGradient-ed (:390-391).

    rois_pred [R,C], cls_prob [1,C], rois [R,5], labels_oh [1,C] ->
    (class_weight [1,C], class_weight_noise [1,C]).
    """
    P = np.asarray(rois_pred, dtype=F32)
    y = np.asarray(cls_prob, dtype=F32).reshape(1, -1)
    L = np.asarray(labels_oh, dtype=F32).reshape(1, -1)
    R = P.shape[0]
    J = roi_iou(rois)                                              # :266
    with np.errstate(divide="ignore", invalid="ignore"):
        E = (P * np.log(P, dtype=F32)) * F32(-1.0)                # :276-278
        E = np.where(np.isnan(E), F32(0), E).astype(F32)          # :279 ReplaceNaN(0)
        D = (J.astype(np.float64) @ E.astype(np.float64)).astype(F32)   # :280 MatMul
        D = np.where(D >= 0, D, F32(0.01) * D).astype(F32)        # :281 LeakyRelu(0.01)
        G = (E / D).astype(F32)                                   # :282
        hatE = (E * G).astype(F32)                                # :283
        hatE_sum = hatE.sum(axis=0, keepdims=True, dtype=np.float64).astype(F32)  # :285-288
        logy = np.log(y, dtype=F32)                               # :335
        logN = np.log(F32(R), dtype=F32)                          # :336
        denom = ((logN - logy) * y).astype(F32)                   # :337-340
        norm = (hatE_sum / denom).astype(F32)                     # :345-347
    norm = np.minimum(np.maximum(norm, F32(0)), F32(1))           # :350-353 Clip
    w_noise = (norm * (F32(1) - L)).astype(F32)                   # :368-371
    w_clean = (F32(1) - w_noise).astype(F32)                      # :373-374
    if return_parts:
        return w_clean, w_noise, dict(J=J, E=E, D=D, hatE_sum=hatE_sum, norm=norm)
    return w_clean, w_noise


# --------------------------------------------------------------------------- #
# a8  (Weighted)CrossEntropyWithLogits  -- CPU code is what the reference runs
# --------------------------------------------------------------------------- #
def cross_entropy_with_logits(X, L, W=None, is_mean=True):
    """``(Weighted)CrossEntropyWithLogits`` forward (ops/cross_entropy_wsl_op.cc:8-45, 88-129).
    Sequential float32 accumulation exactly like the reference loop."""
    X = np.asarray(X, dtype=F32)
    L = np.asarray(L, dtype=F32)
    N, C = X.shape
    Wf = np.ones_like(X) if W is None else np.asarray(W, dtype=F32).reshape(X.shape)
    norm = F32(C) if is_mean else F32(1)
    loss = F32(0)
    xf, lf, wf = X.reshape(-1), L.reshape(-1), Wf.reshape(-1)
    # The reference calls the unqualified C ``log`` on a float: the double overload is the
    # one in scope, so each term is evaluated in double and ``loss -= term`` rounds the
    # running float sum once per class (verified bit-exact against oracle/_ref).
    for i in range(xf.size):
        prob = max(xf[i], LOG_THRESHOLD)
        one_prob = max(F32(1) - xf[i], LOG_THRESHOLD)
        term = float(lf[i]) * np.log(np.float64(prob)) + float(F32(1) - lf[i]) * np.log(np.float64(one_prob))
        if W is not None:
            term = term * float(wf[i])
        loss = F32(np.float64(loss) - term)
    y = F32(loss / norm)
    return F32(y * F32(1.0 / N))


def cross_entropy_with_logits_grad(X, L, dY, W=None, is_mean=True):
    """Gradient (ops/cross_entropy_wsl_op.cc:47-85, 131-180): the 1e4 clamp is upper-side only
    and applied before the per-class weight."""
    X = np.asarray(X, dtype=F32)
    L = np.asarray(L, dtype=F32)
    N, C = X.shape
    norm = F32(C) if is_mean else F32(1)
    g = F32(dY)
    prob = np.maximum(X, LOG_THRESHOLD)
    one_prob = np.maximum(F32(1) - X, LOG_THRESHOLD)
    d = (g * (F32(-1) * L / prob - F32(-1) * (F32(1) - L) / one_prob) / norm).astype(F32)
    d = np.minimum(d, DIFF_THRESHOLD)
    if W is not None:
        d = (d * np.asarray(W, dtype=F32).reshape(X.shape)).astype(F32)
    return (d * F32(1.0 / N)).astype(F32)


# --------------------------------------------------------------------------- #
# a10  ACMWeightDecayMomentumSGDUpdate
# --------------------------------------------------------------------------- #
def acm_sgd_update(g, m, lr, p, acc, *, momentum=0.9, weight_decay=0.0, lr_mult=1.0,
                   iter_size=1, gpu_num=1, iter_count=0):
    """``ACMWeightDecayMomentumSGDUpdate([g, m, lr, p, acc] -> [g, m, p, acc])``
    (ops/acm_weightdecay_momentum_sgd_op.h:48-112; non-nesterov branch :19-22).
    ``iter_count`` is the op's call counter before this call (the reference keeps it as
    hidden state; here it is explicit).  Returns (m, p, acc, iter_count+1)."""
    g = np.asarray(g, dtype=F32)
    m = np.array(m, dtype=F32, copy=True)
    p = np.array(p, dtype=F32, copy=True)
    acc = np.array(acc, dtype=F32, copy=True)
    if iter_count == 0:                                   # :62-69
        acc[...] = 0
        m[...] = 0
    acc = (g + acc).astype(F32)                           # :72-75
    iter_count += 1
    if iter_count % iter_size == 0:
        acc = (acc * F32(1.0 / (iter_size * gpu_num))).astype(F32)   # :79-84
        acc = (acc + F32(weight_decay) * p).astype(F32)             # :88-90 Axpy
        LR = F32(F32(lr) * F32(lr_mult))                            # :15
        adj = (LR * acc + F32(momentum) * m).astype(F32)            # :19
        m = adj
        p = (p - adj).astype(F32)                                   # :30
        acc = np.zeros_like(acc)                                    # :106-109
    return m, p, acc, iter_count


# --------------------------------------------------------------------------- #
# whole head, one image  (modeling/webly_heads.py:463-502, 32-74, 123-197)
# --------------------------------------------------------------------------- #
def fc_stack_forward(feat, W6, b6, W7, b7, mask6=None, mask7=None):
    """fc6 -> Relu -> Dropout -> fc7 -> Relu -> Dropout (modeling/wsl_heads.py:674-679)."""
    fc6 = relu(fc(feat, W6, b6))
    drop6 = dropout(fc6, mask6) if mask6 is not None else fc6
    fc7 = relu(fc(drop6, W7, b7))
    drop7 = dropout(fc7, mask7) if mask7 is not None else fc7
    return dict(fc6=fc6, drop6=drop6, fc7=fc7, drop7=drop7)


def fc_stack_backward(feat, acts, W6, W7, d_drop7, mask6=None, mask7=None, need_dfeat=False,
                      relu_pattern=None):
    """relu_pattern: optional (pattern6, pattern7) boolean arrays that replace ``Y > 0`` in the two
    ReluGradients.  Tests use it to evaluate the gradient arithmetic on the activation pattern
    of the implementation under test: an element whose pre-activation is within rounding error
    of zero may legitimately fall on either side of the ReLU, and its gradient then flips
    between 0 and its full value."""
    d_fc7 = dropout_grad(d_drop7, mask7) if mask7 is not None else d_drop7
    d_fc7 = relu_grad(acts["fc7"] if relu_pattern is None else relu_pattern[1], d_fc7)
    dW7, db7, d_drop6 = fc_grad(acts["drop6"], W7, d_fc7)
    d_fc6 = dropout_grad(d_drop6, mask6) if mask6 is not None else d_drop6
    d_fc6 = relu_grad(acts["fc6"] if relu_pattern is None else relu_pattern[0], d_fc6)
    dW6, db6, d_feat = fc_grad(feat, W6, d_fc6)
    out = dict(fc6_w=dW6, fc6_b=db6, fc7_w=dW7, fc7_b=db7)
    if need_dfeat:
        out["d_feat"] = d_feat
    return out


def mil_head_forward_backward(fc8c, fc8d, rois, labels_oh, nfc8c=None, nfc8d=None,
                              entropy=True, is_mean=True, backward=True):
    """One image: a5 -> a6 -> a7 -> a8 -> a9 on the fc8 logits.

    fc8c/fc8d [R,C]; nfc8c/nfc8d [R,C] or None (plain WSDDN).  Returns a dict with
    rois_pred, cls_prob, (rois_pred_noise, cls_prob_noise), class_weight(_noise),
    loss_cls(_noise) and, if ``backward``, d_fc8c, d_fc8d, d_nfc8c, d_nfc8d
    (loss-gradient seeds are 1.0 each, utils/blob.py:167-173).
    """
    out = {}
    a_cls, a_det, P = wsl_outputs(fc8c, fc8d)
    y = cls_pred(P)
    out.update(alpha_cls=a_cls, alpha_det=a_det, rois_pred=P, cls_prob=y)
    noise = nfc8c is not None
    if noise:
        lc = (np.asarray(fc8c, F32) + np.asarray(nfc8c, F32)).astype(F32)   # webly_heads.py:57-61
        ld = (np.asarray(fc8d, F32) + np.asarray(nfc8d, F32)).astype(F32)
        a_cls_n, a_det_n, Pn = wsl_outputs(lc, ld)
        yn = cls_pred(Pn)
        out.update(rois_pred_noise=Pn, cls_prob_noise=yn)
    L = np.asarray(labels_oh, dtype=F32).reshape(1, -1)
    if noise and entropy:
        w_clean, w_noise = spatial_entropy_weight(P, y, rois, L)
    else:
        w_clean, w_noise = None, None
    out.update(class_weight=w_clean, class_weight_noise=w_noise)
    out["loss_cls"] = cross_entropy_with_logits(y, L, w_clean, is_mean)           # webly_heads.py:167-175
    if noise:
        out["loss_cls_noise"] = cross_entropy_with_logits(yn, L, w_noise, is_mean)  # :183-193
    if not backward:
        return out
    # a9: ReduceSumGradient -> MulGradient -> SoftmaxGradient(s) -> AddGradient
    dy = cross_entropy_with_logits_grad(y, L, F32(1.0), w_clean, is_mean)
    dP = np.broadcast_to(dy, P.shape).astype(F32)
    d_fc8c = softmax_grad(a_cls, (dP * a_det).astype(F32), axis=1)
    d_fc8d = softmax_grad(a_det, (dP * a_cls).astype(F32), axis=0)
    if noise:
        dyn = cross_entropy_with_logits_grad(yn, L, F32(1.0), w_noise, is_mean)
        dPn = np.broadcast_to(dyn, Pn.shape).astype(F32)
        d_lc = softmax_grad(a_cls_n, (dPn * a_det_n).astype(F32), axis=1)
        d_ld = softmax_grad(a_det_n, (dPn * a_cls_n).astype(F32), axis=0)
        out.update(d_nfc8c=d_lc, d_nfc8d=d_ld)
        d_fc8c = (d_fc8c + d_lc).astype(F32)        # Add fans the gradient to both inputs
        d_fc8d = (d_fc8d + d_ld).astype(F32)
    out.update(d_fc8c=d_fc8c, d_fc8d=d_fc8d)
    return out


def head_forward_backward(X, rois, obn_scores, labels_oh, params, spatial_scale=1.0 / 16,
                          masks=None, noise=True, entropy=True, is_mean=True, backward=True,
                          need_dX=False, relu_patterns=None):
    """The whole per-proposal head for ONE image (the reference asserts 1 image per GPU,
    modeling/wsl_heads.py:214): RoIPoolF -> RoIFeatureBoost -> fc6/fc7 (x2 stacks if
    ``noise``) -> fc8c/fc8d (+noisy) -> MIL -> noise-aware losses -> gradients of every
    parameter (as the reference runs it: roi_feat is StopGradient-ed unless ``need_dX``).

    params: dict with fc6_w,fc6_b,fc7_w,fc7_b,fc8c_w,fc8c_b,fc8d_w,fc8d_b and, if noise,
    the same keys prefixed ``noisy_`` (reference blob names ``_[noisy]_fc6_w`` ... and
    ``noisy_fc8c_w`` ...).  masks: dict drop6, drop7(, noisy_drop6, noisy_drop7) or None.
    """
    masks = masks or {}
    Y, A = roi_pool_f(X, rois, spatial_scale)
    feat = roi_feature_boost(Y, obn_scores).reshape(Y.shape[0], -1)
    acts = fc_stack_forward(feat, params["fc6_w"], params["fc6_b"], params["fc7_w"], params["fc7_b"],
                            masks.get("drop6"), masks.get("drop7"))
    fc8c = fc(acts["drop7"], params["fc8c_w"], params["fc8c_b"])
    fc8d = fc(acts["drop7"], params["fc8d_w"], params["fc8d_b"])
    nfc8c = nfc8d = None
    if noise:
        nacts = fc_stack_forward(feat, params["noisy_fc6_w"], params["noisy_fc6_b"],
                                 params["noisy_fc7_w"], params["noisy_fc7_b"],
                                 masks.get("noisy_drop6"), masks.get("noisy_drop7"))
        nfc8c = fc(nacts["drop7"], params["noisy_fc8c_w"], params["noisy_fc8c_b"])
        nfc8d = fc(nacts["drop7"], params["noisy_fc8d_w"], params["noisy_fc8d_b"])
    out = mil_head_forward_backward(fc8c, fc8d, rois, labels_oh, nfc8c, nfc8d,
                                    entropy=entropy, is_mean=is_mean, backward=backward)
    out.update(roi_feat=feat, argmax=A, fc8c=fc8c, fc8d=fc8d, drop7=acts["drop7"], acts=acts)
    if noise:
        out.update(nfc8c=nfc8c, nfc8d=nfc8d, noisy_acts=nacts)
    relu_patterns = relu_patterns or {}
    if not backward:
        return out
    grads = {}
    dWc, dbc, dx_c = fc_grad(acts["drop7"], params["fc8c_w"], out["d_fc8c"])
    dWd, dbd, dx_d = fc_grad(acts["drop7"], params["fc8d_w"], out["d_fc8d"])
    grads.update(fc8c_w=dWc, fc8c_b=dbc, fc8d_w=dWd, fc8d_b=dbd)
    g = fc_stack_backward(feat, acts, params["fc6_w"], params["fc7_w"], (dx_c + dx_d).astype(F32),
                          masks.get("drop6"), masks.get("drop7"), need_dfeat=need_dX,
                          relu_pattern=relu_patterns.get("clean"))
    d_feat = g.pop("d_feat", None)
    grads.update(g)
    if noise:
        dWc, dbc, dx_c = fc_grad(nacts["drop7"], params["noisy_fc8c_w"], out["d_nfc8c"])
        dWd, dbd, dx_d = fc_grad(nacts["drop7"], params["noisy_fc8d_w"], out["d_nfc8d"])
        grads.update(noisy_fc8c_w=dWc, noisy_fc8c_b=dbc, noisy_fc8d_w=dWd, noisy_fc8d_b=dbd)
        g = fc_stack_backward(feat, nacts, params["noisy_fc6_w"], params["noisy_fc7_w"],
                              (dx_c + dx_d).astype(F32), masks.get("noisy_drop6"),
                              masks.get("noisy_drop7"), need_dfeat=need_dX,
                              relu_pattern=relu_patterns.get("noisy"))
        d_feat_n = g.pop("d_feat", None)
        grads.update({"noisy_" + k: v for k, v in g.items()})
        if need_dX:
            d_feat = (d_feat + d_feat_n).astype(F32)
    if need_dX:
        d_pool = roi_feature_boost_grad(d_feat.reshape(Y.shape), obn_scores)
        out["dX"] = roi_pool_f_grad(X.shape, rois, A, d_pool)
    out["grads"] = grads
    return out


# --------------------------------------------------------------------------- #
# synthetic inputs (SURVEY.md section 8d / BASELINE.md section 5)
# --------------------------------------------------------------------------- #
def synth_conv5(n, c, h, w, seed=0):
    """Post-ReLU-like map: U[0,1) * Bernoulli(0.5)."""
    rng = np.random.default_rng(seed)
    return (rng.random((n, c, h, w), dtype=F32) * (rng.random((n, c, h, w)) < 0.5)).astype(F32)


def synth_rois(r, img_h, img_w, batch_idx=0, seed=1, im_scale=1.0):
    """MCG-like integer boxes: side 16 px ... image/2, as (batch_idx, x1, y1, x2, y2)."""
    rng = np.random.default_rng(seed)
    x1 = np.floor(rng.random(r) * (img_w - 17))
    y1 = np.floor(rng.random(r) * (img_h - 17))
    bw = np.floor(16 + rng.random(r) * (img_w / 2 - 16))
    bh = np.floor(16 + rng.random(r) * (img_h / 2 - 16))
    x2 = np.minimum(x1 + bw, img_w - 1)
    y2 = np.minimum(y1 + bh, img_h - 1)
    rois = np.stack([np.full(r, batch_idx), x1, y1, x2, y2], axis=1).astype(F32)
    rois[:, 1:] *= F32(im_scale)
    return rois


def synth_params(c_classes, d_in=25088, hidden=4096, noise=True, seed=2):
    rng = np.random.default_rng(seed)

    def gauss(*s):
        return (rng.standard_normal(s, dtype=F32) * F32(0.01)).astype(F32)

    def xavier(o, i):
        lim = np.sqrt(3.0 / i)     # Caffe2 XavierFill: U(-sqrt(3/fan_in), +)
        return ((rng.random((o, i), dtype=F32) * 2 - 1) * F32(lim)).astype(F32)

    p = {}
    for pre in (["", "noisy_"] if noise else [""]):
        p[pre + "fc6_w"] = gauss(hidden, d_in)
        p[pre + "fc6_b"] = np.zeros(hidden, F32)
        p[pre + "fc7_w"] = gauss(hidden, hidden)
        p[pre + "fc7_b"] = np.zeros(hidden, F32)
        p[pre + "fc8c_w"] = xavier(c_classes, hidden)
        p[pre + "fc8c_b"] = np.zeros(c_classes, F32)
        p[pre + "fc8d_w"] = xavier(c_classes, hidden)
        p[pre + "fc8d_b"] = np.zeros(c_classes, F32)
    return p
