"""Operator-level mirror of the reference's Caffe2 ops on the hot path.

Each function keeps the reference operator's name, blob order and argument meaning
(``[inputs] -> [outputs]; args``) and calls the C ABI of libnawsod.so on the current CUDA
stream.  Tensors are torch CUDA tensors used as device-memory handles only; no arithmetic is
done by PyTorch here.  Errors surface as RuntimeError (the reference's CAFFE_ENFORCE ->
RuntimeError pattern, detectron/tests/test_zero_even_op.py:48-51).

Citations are relative to /root/reference/detectron.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import BF16, F32, NCHW, NHWC

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, name, dtype=None, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libnawsod has no CPU path)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise RuntimeError("%s has dtype %s, expected %s" % (name, t.dtype, dtype))
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError("%s must be %d-d, got shape %s" % (name, ndim, tuple(t.shape)))
    return t


def _layout(s):
    if s in (NCHW, "NCHW"):
        return NCHW
    if s in (NHWC, "NHWC"):
        return NHWC
    raise RuntimeError("unknown layout %r" % (s,))


def transpose_batched(src, out_dtype=None):
    """[B, rows, cols] -> [B, cols, rows] for 4-byte elements (float32 / int32); a float32
    source may be converted to bfloat16 on the way out."""
    _req(src, "src", (torch.float32, torch.int32), 3)
    B, rows, cols = src.shape
    out_dtype = out_dtype or src.dtype
    if out_dtype not in (src.dtype, torch.bfloat16) or (out_dtype == torch.bfloat16 and src.dtype != torch.float32):
        raise RuntimeError("transpose_batched: unsupported conversion %s -> %s" % (src.dtype, out_dtype))
    out = torch.empty((B, cols, rows), dtype=out_dtype, device=src.device)
    _lib.call("nawsod_transpose_batched", _ptr(src), B, rows, cols, _ptr(out),
              BF16 if out_dtype == torch.bfloat16 else F32, _stream())
    return out


def to_channels_last(X, dtype=None):
    """NCHW float32 map -> NHWC (optionally bfloat16) with the library's transpose kernel."""
    _req(X, "X", torch.float32, 4)
    N, C, H, W = X.shape
    return transpose_batched(X.view(N, C, H * W), dtype).view(N, H, W, C)


# --------------------------------------------------------------------------------------------
# RoIPoolF / RoIPoolFGradient / RoIFeatureBoost
# --------------------------------------------------------------------------------------------
def RoIPoolF(X, rois, *, pooled_h=7, pooled_w=7, spatial_scale=1.0 / 16, is_test=False, boost=None,
             x_layout="NCHW", y_layout="NCHW", out_dtype=None):
    """``RoIPoolF([X, rois] -> [Y, argmax]; pooled_h, pooled_w, spatial_scale)``
    (modeling/detector.py:321-329; ``sampling_ratio`` is ignored there too).

    Defaults reproduce the reference blobs exactly: X [N,C,H,W] float32, Y [R,C,ph,pw],
    argmax int32 (None when ``is_test``).  ``x_layout='NHWC'`` / ``y_layout='NHWC'`` select the
    native channels-last kernel directly (X [N,H,W,C], Y [R,ph,pw,C]); the NCHW defaults go
    through the library's transpose kernels around the same pooling kernel, so values and
    argmax are bit-identical in every layout.  ``boost`` ([R] or [R,1], obn_scores+1) fuses
    ``RoIFeatureBoost`` (modeling/wsl_heads.py:668) into the epilogue.
    """
    xl, yl = _layout(x_layout), _layout(y_layout)
    _req(X, "X", (torch.float32, torch.bfloat16), 4)
    _req(rois, "rois", torch.float32, 2)
    if rois.shape[1] != 5:
        raise RuntimeError("rois must be [R,5], got %s" % (tuple(rois.shape),))
    if boost is not None:
        _req(boost, "boost", torch.float32)
        if boost.numel() != rois.shape[0]:
            raise RuntimeError("boost must have one entry per RoI")
    if xl == NCHW:
        if X.dtype != torch.float32:
            raise RuntimeError("an NCHW map must be float32 (the reference layout)")
        N, C, H, W = X.shape
        Xcl = to_channels_last(X)
    else:
        N, H, W, C = X.shape
        Xcl = X
    R = rois.shape[0]
    out_dtype = out_dtype or (torch.float32 if xl == NCHW else X.dtype)
    if yl == NCHW and out_dtype != torch.float32:
        raise RuntimeError("pooled NCHW output is float32 only")
    Y = torch.empty((R, pooled_h, pooled_w, C), dtype=out_dtype, device=X.device)
    A = None if is_test else torch.empty((R, pooled_h, pooled_w, C), dtype=torch.int32, device=X.device)
    _lib.call("nawsod_roi_pool_f_fwd", _ptr(Xcl), _DT[Xcl.dtype], NHWC, _ptr(rois), _ptr(boost), N, C, H, W, R,
              float(spatial_scale), pooled_h, pooled_w, _ptr(Y), _DT[out_dtype], NHWC, _ptr(A), _stream())
    if yl == NCHW:
        bins = pooled_h * pooled_w
        Y = transpose_batched(Y.view(R, bins, C)).view(R, C, pooled_h, pooled_w)
        if A is not None:
            A = transpose_batched(A.view(R, bins, C)).view(R, C, pooled_h, pooled_w)
    return Y, A


def RoIPoolFGradient(X, rois, argmax, dY, *, boost=None, layout="NCHW"):
    """``RoIPoolFGradient([X, rois, argmax, dY] -> dX)`` (grad maker ops/roi_loop_pool_op.cc:85-96).
    X is only consulted for its shape, as in the reference.  layout: layout of X/dX and of
    dY/argmax (NCHW: [N,C,H,W] / [R,C,ph,pw]; NHWC: [N,H,W,C] / [R,ph,pw,C])."""
    lay = _layout(layout)
    _req(argmax, "argmax", torch.int32, 4)
    _req(dY, "dY", (torch.float32, torch.bfloat16), 4)
    _req(rois, "rois", torch.float32, 2)
    if tuple(argmax.shape) != tuple(dY.shape):
        raise RuntimeError("argmax and dY must have the same shape")
    if lay == NCHW:
        N, C, H, W = X.shape
        R, C2, ph, pw = dY.shape
    else:
        N, H, W, C = X.shape
        R, ph, pw, C2 = dY.shape
    if C2 != C or R != rois.shape[0]:
        raise RuntimeError("shape mismatch between X, rois and dY")
    dX = torch.empty(tuple(X.shape), dtype=torch.float32, device=dY.device)
    _lib.call("nawsod_roi_pool_f_bwd", _ptr(dY), _DT[dY.dtype], lay, _ptr(argmax), _ptr(rois), _ptr(boost), N, C, H, W,
              R, ph, pw, _ptr(dX), lay, _stream())
    return dX


def RoIFeatureBoost(X, S, out=None):
    """``RoIFeatureBoost([X, S] -> [Y])``, in place when ``out is X`` (ops/roi_feature_boost_op.cc:8-35,74-82)."""
    _req(X, "X", torch.float32)
    _req(S, "S", torch.float32)
    if S.numel() != S.shape[0] or X.shape[0] != S.shape[0]:      # CAFFE_ENFORCE_EQ(S.dim32(0), S.numel()) ...
        raise RuntimeError("RoIFeatureBoost: S must be [R] or [R,1] with R == X.shape[0]")
    Y = torch.empty_like(X) if out is None else out
    R = X.shape[0]
    _lib.call("nawsod_roi_feature_boost", _ptr(X), _ptr(S), R, X.numel() // max(R, 1), _ptr(Y), _stream())
    return Y


def RoIFeatureBoostGradient(dY, S, out=None):
    """``RoIFeatureBoostGradient([dY, S] -> [dX])`` (ops/roi_feature_boost_op.cc:37-64)."""
    return RoIFeatureBoost(dY, S, out)


# --------------------------------------------------------------------------------------------
# RoIIoU, [Weighted]CrossEntropyWithLogits
# --------------------------------------------------------------------------------------------
def RoIIoU(rois):
    """``RoIIoU([rois] -> [J])`` (ops/roi_iou_op.cc:11-18)."""
    _req(rois, "rois", torch.float32, 2)              # CAFFE_ENFORCE_EQ(R.dim(), 2)
    if rois.shape[1] != 5:                            # CAFFE_ENFORCE_EQ(R.dim32(1), 5)
        raise RuntimeError("RoIIoU: rois must be [R,5]")
    n = rois.shape[0]
    J = torch.empty((n, n), dtype=torch.float32, device=rois.device)
    _lib.call("nawsod_roi_iou", _ptr(rois), n, _ptr(J), _stream())
    return J


def _ce_check(X, L, W):
    _req(X, "X", torch.float32, 2)                    # CAFFE_ENFORCE_EQ(X.dim(), 2)
    _req(L, "L", torch.float32)
    if tuple(X.shape) != tuple(L.shape):              # CAFFE_ENFORCE_EQ(X.sizes(), L.sizes())
        raise RuntimeError("CrossEntropyWithLogits: X and L shapes differ")
    if W is not None:
        _req(W, "W", torch.float32)
        if tuple(X.shape) != tuple(W.shape):
            raise RuntimeError("WeightedCrossEntropyWithLogits: X and W shapes differ")


def CrossEntropyWithLogits(X, L, *, is_mean=False):
    """``CrossEntropyWithLogits([X, L] -> [Y]; is_mean)`` (ops/cross_entropy_wsl_op.cc:8-45)."""
    _ce_check(X, L, None)
    Y = torch.empty((), dtype=torch.float32, device=X.device)
    _lib.call("nawsod_cross_entropy_fwd", _ptr(X), _ptr(L), None, X.shape[0], X.shape[1], int(bool(is_mean)), _ptr(Y),
              _stream())
    return Y


def WeightedCrossEntropyWithLogits(X, L, W, *, is_mean=False):
    """``WeightedCrossEntropyWithLogits([X, L, W] -> [Y]; is_mean)`` (ops/cross_entropy_wsl_op.cc:88-129)."""
    _ce_check(X, L, W)
    Y = torch.empty((), dtype=torch.float32, device=X.device)
    _lib.call("nawsod_cross_entropy_fwd", _ptr(X), _ptr(L), _ptr(W), X.shape[0], X.shape[1], int(bool(is_mean)),
              _ptr(Y), _stream())
    return Y


def CrossEntropyWithLogitsGradient(X, L, dY, *, is_mean=False):
    """``([X, L, dY] -> [dX])`` (ops/cross_entropy_wsl_op.cc:47-85)."""
    _ce_check(X, L, None)
    _req(dY, "dY", torch.float32)
    if dY.numel() != 1:                               # CAFFE_ENFORCE_EQ(dY.numel(), 1)
        raise RuntimeError("dY must have one element")
    dX = torch.empty_like(X)
    _lib.call("nawsod_cross_entropy_bwd", _ptr(X), _ptr(L), None, _ptr(dY), X.shape[0], X.shape[1],
              int(bool(is_mean)), _ptr(dX), _stream())
    return dX


def WeightedCrossEntropyWithLogitsGradient(X, L, W, dY, *, is_mean=False):
    """``([X, L, W, dY] -> [dX])`` (ops/cross_entropy_wsl_op.cc:131-180)."""
    _ce_check(X, L, W)
    _req(dY, "dY", torch.float32)
    if dY.numel() != 1:
        raise RuntimeError("dY must have one element")
    dX = torch.empty_like(X)
    _lib.call("nawsod_cross_entropy_bwd", _ptr(X), _ptr(L), _ptr(W), _ptr(dY), X.shape[0], X.shape[1],
              int(bool(is_mean)), _ptr(dX), _stream())
    return dX


# --------------------------------------------------------------------------------------------
# fused MIL head + losses (a5..a9)
# --------------------------------------------------------------------------------------------
_ws_cache = {}


def _workspace(nbytes, device, tag):
    key = (tag, str(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def mil_head(fc8c, fc8d, rois, roi_offsets, labels_oh, nfc8c=None, nfc8d=None, *, entropy=True, is_mean=True,
             backward=True, grads_out=None):
    """The whole of ``add_wsl_outputs`` + ``add_webly_outputs`` + ``add_webly_losses`` and their
    gradient ops in one kernel (modeling/wsl_heads.py:23-56,213-227; modeling/webly_heads.py:32-74,
    123-197,265-391).  ``roi_offsets`` [B+1] int32 (device): rows of image b are
    roi_offsets[b]:roi_offsets[b+1].  The logits may be column slices of wider matrices (all
    with one common row pitch), e.g. the halves of a fused [R,2C] fc8 output.  ``grads_out``:
    optional dict of preallocated (possibly sliced) d_fc8c/d_fc8d/d_nfc8c/d_nfc8d tensors.
    Returns a dict of the reference's blob names."""
    R, C, ldl = _mat(fc8c, "fc8c")
    B = labels_oh.shape[0]
    noise = nfc8c is not None
    logits = [fc8d] + ([nfc8c, nfc8d] if noise else [])
    for t in [fc8c] + logits:
        r2, c2, l2 = _mat(t, "logits")
        if t.dtype != torch.float32 or (r2, c2, l2) != (R, C, ldl):
            raise RuntimeError("mil_head: logits must be float32 [R,C] with one common row pitch")
    _req(rois, "rois", torch.float32, 2)
    _req(roi_offsets, "roi_offsets", torch.int32, 1)
    _req(labels_oh, "labels_oh", torch.float32, 2)
    if rois.shape[0] != R or labels_oh.shape[1] != C or roi_offsets.numel() != B + 1:
        raise RuntimeError("mil_head: inconsistent shapes")
    dev = fc8c.device
    new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = {"rois_pred": new(R, C), "cls_prob": new(B, C), "loss": new(B, 2)}
    if noise:
        out.update(rois_pred_noise=new(R, C), cls_prob_noise=new(B, C))
        if entropy:
            out.update(class_weight=new(B, C), class_weight_noise=new(B, C))
    ldg = C
    if backward:
        names = ["d_fc8c", "d_fc8d"] + (["d_nfc8c", "d_nfc8d"] if noise else [])
        if grads_out is None:
            out.update({n: new(R, C) for n in names})
        else:
            ldg = None
            for n in names:
                r2, c2, l2 = _mat(grads_out[n], n)
                if grads_out[n].dtype != torch.float32 or (r2, c2) != (R, C) or (ldg is not None and l2 != ldg):
                    raise RuntimeError("mil_head: grads_out tensors must be float32 [R,C] with one row pitch")
                ldg = l2
                out[n] = grads_out[n]
    flags = (_lib.MIL_ENTROPY if entropy else 0) | (_lib.MIL_MEAN if is_mean else 0) | \
            (_lib.MIL_BACKWARD if backward else 0)
    ws = _workspace(_lib.load().nawsod_mil_workspace_bytes(R, C, B), dev, "mil")
    g = out.get
    _lib.call("nawsod_mil_head_fwd_bwd", _ptr(fc8c), _ptr(fc8d), _ptr(nfc8c), _ptr(nfc8d), ldl, _ptr(rois),
              _ptr(roi_offsets), _ptr(labels_oh), R, C, B, flags, _ptr(g("rois_pred")), _ptr(g("cls_prob")),
              _ptr(g("rois_pred_noise")), _ptr(g("cls_prob_noise")), _ptr(g("class_weight")),
              _ptr(g("class_weight_noise")), _ptr(g("loss")), _ptr(g("d_fc8c")), _ptr(g("d_fc8d")),
              _ptr(g("d_nfc8c")), _ptr(g("d_nfc8d")), ldg, _ptr(ws), _stream())
    return out


# --------------------------------------------------------------------------------------------
# ACMWeightDecayMomentumSGDUpdate
# --------------------------------------------------------------------------------------------
def ACMWeightDecayMomentumSGDUpdate(g, m, lr, p, acc, *, momentum=0.9, iter_size=1, gpu_num=1, lr_mult=1.0,
                                    weight_decay=0.0, iter_count=0, p_shadow=None):
    """``ACMWeightDecayMomentumSGDUpdate([g, m, lr, p, acc] -> [g, m, p, acc])`` in place
    (ops/acm_weightdecay_momentum_sgd_op.cc:7-22; wiring modeling/optimizer_wsl.py:127-136).
    ``iter_count`` replaces the op's hidden ``iter_count_`` member; ``acc`` may be None when
    iter_size == 1 (the accumulator is then identically zero between calls)."""
    for t, nme in ((g, "g"), (m, "m"), (p, "p")):
        _req(t, nme, torch.float32)
    _req(lr, "lr", torch.float32)
    if lr.numel() != 1:                               # CAFFE_ENFORCE_EQ(Input(LR).numel(), 1)
        raise RuntimeError("lr must have one element")
    if g.numel() != m.numel() or g.numel() != p.numel() or (acc is not None and acc.numel() != g.numel()):
        raise RuntimeError("g, m, p, acc must have the same number of elements")
    if p_shadow is not None:
        _req(p_shadow, "p_shadow", (torch.bfloat16, torch.float32))   # bf16 copy, or float rounded to TF32
    _lib.call("nawsod_sgd_update", _ptr(g), _ptr(m), _ptr(lr), _ptr(p), _ptr(acc), g.numel(), float(momentum),
              float(weight_decay), float(lr_mult), int(iter_size), int(gpu_num), int(iter_count), _ptr(p_shadow),
              _DT[p_shadow.dtype] if p_shadow is not None else F32, _stream())
    return g, m, p, acc


def ACMWeightDecayMomentumSGDUpdateReduce(grads, m, lr, p, *, momentum=0.9, gpu_num=1, lr_mult=1.0, weight_decay=0.0,
                                          iter_count=0, p_shadow=None, abort_flag=None):
    """Data-parallel owner's update: ``g = grads[0] + grads[1] + ...`` (in that order: the ranks'
    contributions to this parameter slice), then ``ACMWeightDecayMomentumSGDUpdate`` with
    iter_size 1 -- i.e. the reference's NCCLAllreduce + update pair (modeling/optimizer_wsl.py:52-72,
    96-137) restricted to the slice this rank owns, in one pass over HBM.  A gradient source may be a tensor or the raw
    device address of ``m.numel()`` floats (a peer's gradient slice mapped into this process: the kernel then reads it
    over NVLink, all sources' loads in flight together).  ``abort_flag`` (int32 CUDA word): a
    non-zero value at launch time (a peer exchange whose watchdog fired) makes the call a no-op."""
    n = m.numel()
    for t, nme in ((m, "m"), (p, "p")):
        _req(t, nme, torch.float32)
    _req(lr, "lr", torch.float32)
    if not grads or p.numel() != n:
        raise RuntimeError("need at least one gradient source and matching m / p sizes")
    for i, g in enumerate(grads):
        if isinstance(g, int):                  # raw device address of n floats (a peer-mapped gradient slice)
            if g == 0 or g % 16:
                raise RuntimeError("grads[%d]: null or misaligned device address" % i)
            continue
        _req(g, "grads[%d]" % i, torch.float32)
        if g.numel() != n:
            raise RuntimeError("grads[%d] has %d elements, expected %d" % (i, g.numel(), n))
    if p_shadow is not None:
        _req(p_shadow, "p_shadow", (torch.bfloat16, torch.float32))
    table = (ctypes.c_void_p * len(grads))(*[g if isinstance(g, int) else g.data_ptr() for g in grads])
    _lib.call("nawsod_sgd_update_reduce", table, len(grads), _ptr(m), _ptr(lr), _ptr(p), n, float(momentum),
              float(weight_decay), float(lr_mult), int(gpu_num), int(iter_count), _ptr(p_shadow),
              _DT[p_shadow.dtype] if p_shadow is not None else F32, _ptr(abort_flag), _stream())
    return m, p


# --------------------------------------------------------------------------------------------
# peer-to-peer plumbing of the gradient exchange (raw device addresses; see csrc/p2p.cu)
# --------------------------------------------------------------------------------------------
def p2p_copy(dst_ptr: int, src_ptr: int, nbytes: int):
    """Copy-engine transfer between local / peer-mapped device buffers on the current stream."""
    _lib.call("nawsod_p2p_copy", ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), int(nbytes), _stream())


def p2p_signal(flag_ptrs, value: int):
    """Publish ``value`` into every flag word (system-scope release), in stream order."""
    table = (ctypes.c_void_p * len(flag_ptrs))(*flag_ptrs)
    _lib.call("nawsod_p2p_signal", table, len(flag_ptrs), int(value) & 0xFFFFFFFF, _stream())


def p2p_scatter(srcs, dsts, nbytes: int, flag_ptrs, value: int, slot: int):
    """One SM-driven launch: copy ``nbytes`` from srcs[i] to dsts[i] (raw addresses, local or peer-mapped) for every
    peer, then publish ``value`` into the flag words."""
    ts = (ctypes.c_void_p * max(len(srcs), 1))(*srcs)
    td = (ctypes.c_void_p * max(len(dsts), 1))(*dsts)
    tf = (ctypes.c_void_p * max(len(flag_ptrs), 1))(*flag_ptrs)
    _lib.call("nawsod_p2p_scatter", ts, td, len(srcs), int(nbytes), tf, len(flag_ptrs), int(value) & 0xFFFFFFFF, int(slot), _stream())


def p2p_wait(flags, value: int, timeout_ms: int = 20000, status=None):
    """Block the current stream until every word of ``flags`` (int32 CUDA tensor) reached ``value``."""
    _req(flags, "flags", torch.int32)
    _lib.call("nawsod_p2p_wait", _ptr(flags), flags.numel(), int(value) & 0xFFFFFFFF, int(timeout_ms), _ptr(status), _stream())


# --------------------------------------------------------------------------------------------
# FC / FCGradient on the tcgen05 tensor cores (+ fused Relu / Dropout and their gradients)
# --------------------------------------------------------------------------------------------
def _mat(t, name):
    """2-D matrix view: rows, cols, leading dimension (elements).  Column slices are allowed."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libnawsod has no CPU path)" % name)
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError("%s must be a row-major 2-d matrix (unit column stride)" % name)
    return t.shape[0], t.shape[1], t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _ab(dtype):
    if dtype not in _DT:
        raise RuntimeError("FC operands must be float32 (TF32 path) or bfloat16, got %s" % dtype)
    return _DT[dtype]


def _mat3(t, name):
    """Stack of equally shaped row-major matrices: (S, rows, cols, leading dimension, stack stride), all in
    elements.  A 2-d matrix is a stack of one.  Stacks are strided views (e.g. the per-stack column blocks
    of an [R, S*H] activation buffer, ``buf.view(R, S, H).permute(1, 0, 2)``) -- nothing is copied."""
    if isinstance(t, torch.Tensor) and t.dim() == 2:
        r, c, ld = _mat(t, name)
        return 1, r, c, ld, 0
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libnawsod has no CPU path)" % name)
    if t.dim() != 3 or t.stride(2) != 1:
        raise RuntimeError("%s must be a 2-d matrix or a 3-d stack of row-major matrices" % name)
    ld = t.stride(1) if t.shape[1] > 1 else max(t.stride(1), t.shape[2])
    return t.shape[0], t.shape[1], t.shape[2], ld, (t.stride(0) if t.shape[0] > 1 else 0)


def _same_stacks(who, *ss):
    S = max(ss)
    if any(x not in (S,) for x in ss):
        raise RuntimeError("%s: operands disagree on the number of stacks %s" % (who, ss))
    return S


# Split-operand (fp32) products sum their dominant high x high term in reduction chunks of this many elements, added up by the
# epilogue's round-to-nearest fp32 adds (FC_ACCUMULATE).  Measured on B200 (profiles/r2q_pytest_gpu.log): a tcgen05 kind::tf32
# accumulation chain loses ~n_mma * 2^-25 of the running sum (the TMEM accumulator is updated with truncation, not rounding): 4e-6
# of the output scale at K = 1568, 9e-6 at K = 4096, ~1e-4 at fc6's K = 25088 (3136 chained MMAs) -- two orders above the 2^-21 the
# operand split itself leaves.  128 chained MMAs per chunk bound the loss at ~4e-6; the two low-order passes are 2^-11 smaller and
# run unchunked.
X3_CHUNK = 1024


def FC(X, W, b=None, *, relu=False, dropout_mask=None, dropout=False, dropout_seed=0, out=None, out_dtype=None,
       round_tf32=False, gate=None, accumulate=False, X_lo=None, W_lo=None):
    """``FC([X, W, b] -> Y)`` with W [out, in] (Caffe2 layout), optionally fused with the
    ``Relu`` and ``Dropout(ratio=0.5, is_test=0)`` that follow it in the head
    (modeling/wsl_heads.py:674-679).  X may be a column slice of a wider matrix.  3-d operands
    ([S, ., .] strided views, b [S, N]) run the S stacks of the head as ONE launch; stack s draws its
    seeded dropout bits from ``dropout_seed + s``.

    ``gate`` (2-d operands only) = dict(flags=int32 CUDA tensor [groups * nflags], nflags, rows, seq, timeout_ms, status):
    rows ``[g*rows, (g+1)*rows)`` of W are read only once flags ``[g*nflags, (g+1)*nflags)`` have reached ``seq`` -- the
    data-parallel peer exchange's "operands of bucket g have landed" words (dp.P2PExchange), so the GEMM runs on the weights
    that are there and meets the rest as they arrive.

    ``accumulate`` (float32 ``out``): the prior contents of ``out`` are added to X.W^T before bias / Relu / Dropout.
    ``X_lo`` / ``W_lo`` (float32 operands, both or neither; see :func:`split_tf32`): the fp32 path -- X, W are the TF32 high
    parts, and the product is summed from three tensor-core passes ``X_lo.W^T + X.W_lo^T + X.W^T`` (the epilogue runs in the last)."""
    if (X_lo is None) != (W_lo is None):
        raise RuntimeError("FC: pass both X_lo and W_lo (split-operand fp32 path) or neither")
    if X_lo is not None:
        if gate is not None or accumulate:
            raise RuntimeError("FC: the split-operand path takes neither a gate nor accumulate")
        if out is None:
            out = torch.empty((X.shape[0], W.shape[0]) if X.dim() == 2 else (X.shape[0], X.shape[1], W.shape[1]),
                              dtype=torch.float32, device=X.device)
        FC(X_lo, W, out=out)
        FC(X, W_lo, out=out, accumulate=True)
        K = X.shape[-1]
        for k0 in range(0, K, X3_CHUNK):             # high x high in reduction chunks (see X3_CHUNK)
            k1 = min(K, k0 + X3_CHUNK)
            if k1 < K:
                FC(X[..., k0:k1], W[..., k0:k1], out=out, accumulate=True)
        return FC(X[..., k0:K], W[..., k0:K], b, relu=relu, dropout_mask=dropout_mask, dropout=dropout, dropout_seed=dropout_seed,
                  out=out, round_tf32=round_tf32, accumulate=True)
    S, M, K, lda, sA = _mat3(X, "X")
    S2, N, K2, ldw, sW = _mat3(W, "W")
    if K != K2 or X.dtype != W.dtype:
        raise RuntimeError("FC: X [%d,%d] %s and W [%d,%d] %s do not match" % (M, K, X.dtype, N, K2, W.dtype))
    sb = 0
    if b is not None:
        _req(b, "b", torch.float32) if b.dim() == 1 else None
        if b.dtype != torch.float32 or not b.is_cuda or b.shape[-1] != N or b.stride(-1) != 1 or b.numel() != S * N:
            raise RuntimeError("FC: bias must be float32 with %d elements per stack" % N)
        sb = b.stride(0) if b.dim() == 2 and S > 1 else 0
    out_dtype = out_dtype or (out.dtype if out is not None else X.dtype)
    Y = torch.empty((M, N) if X.dim() == 2 else (S, M, N), dtype=out_dtype, device=X.device) if out is None else out
    S3, M3, N2, ldy, sY = _mat3(Y, "Y")
    _same_stacks("FC", S, S2, S3)
    if N2 != N or M3 != M:
        raise RuntimeError("FC: out has the wrong shape")
    if dropout and dropout_mask is None and not dropout_seed:
        raise RuntimeError("FC: dropout=True needs a dropout_mask or a non-zero dropout_seed")
    flags = (_lib.FC_RELU if relu else 0) | \
            (_lib.FC_DROPOUT if (dropout or dropout_mask is not None or dropout_seed) else 0) | \
            (_lib.FC_ROUND_TF32 if round_tf32 else 0) | (_lib.FC_ACCUMULATE if accumulate else 0)
    if accumulate and (out is None or Y.dtype != torch.float32):
        raise RuntimeError("FC: accumulate adds into a given float32 out")
    ldm = sm = 0
    if dropout_mask is not None:
        if dropout_mask.dtype != torch.uint8:
            raise RuntimeError("dropout_mask must be uint8 (0/1)")
        S4, _, _, ldm, sm = _mat3(dropout_mask, "dropout_mask")
        _same_stacks("FC", S, S4)
    if gate is not None:
        if S != 1:
            raise RuntimeError("FC: a gated launch takes 2-d operands")
        gf = gate["flags"]
        _req(gf, "gate flags", torch.int32)
        groups = gf.numel() // int(gate["nflags"])
        _lib.call("nawsod_fc_fwd_gated", _ptr(X), lda, _ptr(W), ldw, _ptr(b), _ptr(dropout_mask), ldm, int(dropout_seed), M, N, K,
                  _ab(X.dtype), _ptr(Y), ldy, _DT[Y.dtype], flags, _ptr(gf), groups, int(gate["nflags"]), int(gate["rows"]),
                  int(gate["seq"]) & 0xFFFFFFFF, int(gate.get("timeout_ms", 20000)), _ptr(gate.get("status")), _stream())
        return Y
    _lib.call("nawsod_fc_fwd_stacks", _ptr(X), lda, sA, _ptr(W), ldw, sW, _ptr(b), sb, _ptr(dropout_mask), ldm, sm,
              int(dropout_seed), S, M, N, K, _ab(X.dtype), _ptr(Y), ldy, sY, _DT[Y.dtype], flags, _stream())
    return Y


def FCGradientX(dY, W, *, act_below=None, mask_below=None, dropout=False, out=None, out_dtype=None, round_tf32=False,
                accumulate=False, dY_lo=None, W_lo=None):
    """dX of ``FCGradient([X, W, dY] -> [dW, db, dX])`` fused with the ``DropoutGradient`` and
    ``ReluGradient`` of the layer below: dX = (dY . W) * 2[dropout] * (act_below > 0).  3-d operands = stacks.
    ``accumulate`` / ``dY_lo`` + ``W_lo``: as in :func:`FC` (prior contents of ``out`` added before the gating; three passes)."""
    if (dY_lo is None) != (W_lo is None):
        raise RuntimeError("FCGradientX: pass both dY_lo and W_lo (split-operand fp32 path) or neither")
    if dY_lo is not None:
        if accumulate:
            raise RuntimeError("FCGradientX: the split-operand path does not take accumulate")
        if out is None:
            out = torch.empty((dY.shape[0], W.shape[1]) if dY.dim() == 2 else (dY.shape[0], dY.shape[1], W.shape[2]),
                              dtype=torch.float32, device=dY.device)
        FCGradientX(dY_lo, W, out=out)
        FCGradientX(dY, W_lo, out=out, accumulate=True)
        N = dY.shape[-1]
        for n0 in range(0, N, X3_CHUNK):             # the reduction runs over the layer's outputs
            n1 = min(N, n0 + X3_CHUNK)
            if n1 < N:
                FCGradientX(dY[..., n0:n1], W[..., n0:n1, :], out=out, accumulate=True)
        return FCGradientX(dY[..., n0:N], W[..., n0:N, :], act_below=act_below, mask_below=mask_below, dropout=dropout, out=out,
                           round_tf32=round_tf32, accumulate=True)
    S, M, N, lddy, sdY = _mat3(dY, "dY")
    S2, N2, K, ldw, sW = _mat3(W, "W")
    if N != N2 or dY.dtype != W.dtype:
        raise RuntimeError("FCGradientX: dY and W do not match")
    out_dtype = out_dtype or (out.dtype if out is not None else dY.dtype)
    dX = torch.empty((M, K) if dY.dim() == 2 else (S, M, K), dtype=out_dtype, device=dY.device) if out is None else out
    S3, M3, K2, ldda, sdA = _mat3(dX, "dX")
    _same_stacks("FCGradientX", S, S2, S3)
    if K2 != K or M3 != M:
        raise RuntimeError("FCGradientX: out has the wrong shape")
    flags = (_lib.FC_RELU if act_below is not None else 0) | (_lib.FC_DROPOUT if (dropout or mask_below is not None) else 0) | \
            (_lib.FC_ROUND_TF32 if round_tf32 else 0) | (_lib.FC_ACCUMULATE if accumulate else 0)
    if accumulate and (out is None or dX.dtype != torch.float32):
        raise RuntimeError("FCGradientX: accumulate adds into a given float32 out")
    ldact, sact, act_dt, ldm, sm = 0, 0, F32, 0, 0
    if act_below is not None:
        S4, _, _, ldact, sact = _mat3(act_below, "act_below")
        _same_stacks("FCGradientX", S, S4)
        act_dt = _DT[act_below.dtype]
    if mask_below is not None:
        S5, _, _, ldm, sm = _mat3(mask_below, "mask_below")
        _same_stacks("FCGradientX", S, S5)
    _lib.call("nawsod_fc_bwd_x_stacks", _ptr(dY), lddy, sdY, _ptr(W), ldw, sW, _ptr(act_below), ldact, sact, act_dt,
              _ptr(mask_below), ldm, sm, S, M, N, K, _ab(dY.dtype), _ptr(dX), ldda, sdA, _DT[dX.dtype], flags, _stream())
    return dX


def FCGradientW(dY, X, *, dW=None, db=None, want_db=True, accumulate=False, dY_lo=None, X_lo=None):
    """dW, db of ``FCGradient``: dW [N,K] = dY^T . X (float32), db [N] = column sums of dY.  3-d operands = stacks
    (dW [S,N,K], db [S,N]).  ``dY_lo`` + ``X_lo``: the split-operand fp32 path (see :func:`FC`): three accumulating passes
    ``dY_lo^T.X + dY^T.X_lo + dY^T.X``; db sums the high and the low part of dY."""
    if (dY_lo is None) != (X_lo is None):
        raise RuntimeError("FCGradientW: pass both dY_lo and X_lo (split-operand fp32 path) or neither")
    if dY_lo is not None:
        dW, db = FCGradientW(dY_lo, X, dW=dW, db=db, want_db=want_db, accumulate=accumulate)
        FCGradientW(dY, X_lo, dW=dW, want_db=False, accumulate=True)
        M = dY.shape[-2]
        for m0 in range(0, M, X3_CHUNK):             # the reduction runs over the RoIs; each chunk adds its rows' column sums to db
            m1 = min(M, m0 + X3_CHUNK)
            FCGradientW(dY[..., m0:m1, :], X[..., m0:m1, :], dW=dW, db=db, want_db=want_db, accumulate=True)
        return dW, db
    S, M, N, lddy, sdY = _mat3(dY, "dY")
    S2, M2, K, lda, sA = _mat3(X, "X")
    if M != M2 or dY.dtype != X.dtype:
        raise RuntimeError("FCGradientW: dY and X do not match")
    if dW is None:
        dW = torch.empty((N, K) if dY.dim() == 2 else (S, N, K), dtype=torch.float32, device=dY.device)
    S3, N3, K3, lddw, sdW = _mat3(dW, "dW")
    _same_stacks("FCGradientW", S, S2, S3)
    if dW.dtype != torch.float32 or (N3, K3) != (N, K):
        raise RuntimeError("FCGradientW: dW must be float32 [%d,%d] per stack" % (N, K))
    if db is None and want_db:
        db = torch.empty((N,) if dY.dim() == 2 else (S, N), dtype=torch.float32, device=dY.device)
    sdb = 0
    if db is not None:
        if db.dtype != torch.float32 or db.shape[-1] != N or db.stride(-1) != 1 or db.numel() != S * N:
            raise RuntimeError("FCGradientW: db must be float32 with %d elements per stack" % N)
        sdb = db.stride(0) if db.dim() == 2 and S > 1 else 0
    _lib.call("nawsod_fc_bwd_w_stacks", _ptr(dY), lddy, sdY, _ptr(X), lda, sA, S, M, N, K, _ab(dY.dtype), _ptr(dW), lddw, sdW,
              _ptr(db), sdb, _lib.FC_ACCUMULATE if accumulate else 0, _stream(), extra_kernels=S if db is not None else 0)
    return dW, db


def FCBiasGradient(dY, db, *, accumulate=False):
    """db of ``FCGradient`` on its own: db [N] (or [S,N] for a stack dY [S,M,N]) = column sums of dY, float32 -- the same
    kernel ``FCGradientW`` runs for its ``db``, callable on another stream than the GEMMs."""
    S, M, N, lddy, sdY = _mat3(dY, "dY")
    if db.dtype != torch.float32 or not db.is_cuda or db.shape[-1] != N or db.stride(-1) != 1 or db.numel() != S * N:
        raise RuntimeError("FCBiasGradient: db must be float32 with %d elements per stack" % N)
    sdb = db.stride(0) if db.dim() == 2 and S > 1 else 0
    _lib.call("nawsod_fc_bias_grad", _ptr(dY), lddy, sdY, S, M, N, _ab(dY.dtype), _ptr(db), sdb,
              _lib.FC_ACCUMULATE if accumulate else 0, _stream(), extra_kernels=S)
    return db


def to_bf16(src, out=None):
    """float32 [rows, cols] (may be a column slice) -> bfloat16, by the library's conversion kernel."""
    rows, cols, lds = _mat(src, "src")
    if src.dtype != torch.float32:
        raise RuntimeError("to_bf16: source must be float32")
    dst = torch.empty((rows, cols), dtype=torch.bfloat16, device=src.device) if out is None else out
    _, _, ldd = _mat(dst, "dst")
    _lib.call("nawsod_convert_f32_to_bf16", _ptr(src), lds, rows, cols, _ptr(dst), ldd, _stream())
    return dst


def round_to_tf32(src, out=None):
    """float32 -> nearest TF32 value in a float32 container (``out`` may be ``src``)."""
    rows, cols, lds = _mat(src, "src")
    if src.dtype != torch.float32:
        raise RuntimeError("round_to_tf32: source must be float32")
    dst = torch.empty((rows, cols), dtype=torch.float32, device=src.device) if out is None else out
    _, _, ldd = _mat(dst, "dst")
    _lib.call("nawsod_round_to_tf32", _ptr(src), lds, rows, cols, _ptr(dst), ldd, _stream())
    return dst


def split_tf32(src, hi=None, lo=None):
    """float32 ``src`` -> (hi, lo): hi = nearest TF32 of src (``hi`` may be ``src``), lo = nearest TF32 of the exact remainder
    -- the operand pair of the fp32 path's three-pass products (``FC(..., X_lo=, W_lo=)``)."""
    rows, cols, lds = _mat(src, "src")
    if src.dtype != torch.float32:
        raise RuntimeError("split_tf32: source must be float32")
    hi = torch.empty((rows, cols), dtype=torch.float32, device=src.device) if hi is None else hi
    lo = torch.empty((rows, cols), dtype=torch.float32, device=src.device) if lo is None else lo
    _, _, ldh = _mat(hi, "hi")
    _, _, ldl = _mat(lo, "lo")
    if hi.dtype != torch.float32 or lo.dtype != torch.float32 or tuple(hi.shape) != (rows, cols) or tuple(lo.shape) != (rows, cols):
        raise RuntimeError("split_tf32: hi and lo must be float32 of the source's shape")
    _lib.call("nawsod_split_tf32", _ptr(src), lds, rows, cols, _ptr(hi), ldh, _ptr(lo), ldl, _stream())
    return hi, lo


# --------------------------------------------------------------------------------------------
# Test-time wrapper and MinEntropyLoss (N4); see csrc/post.cu
# --------------------------------------------------------------------------------------------
def project_rois(boxes, im_scale, *, flip_width=None, batch_idx=0, obn_scores=None):
    """``_get_rois_blob(im_rois, im_scale)`` (core/test_wsl.py:998-1027), optionally on the horizontally
    flipped boxes (utils/boxes.py:246-251), plus ``obn_scores + 1`` (core/test_wsl.py:1058).
    Returns rois [R,5] (and the boosted obn_scores [R] when given)."""
    _req(boxes, "boxes", torch.float32, 2)
    if boxes.shape[1] != 4:
        raise RuntimeError("boxes must be [R,4]")
    R = boxes.shape[0]
    rois = torch.empty((R, 5), dtype=torch.float32, device=boxes.device)
    obn_out = None
    if obn_scores is not None:
        _req(obn_scores, "obn_scores", torch.float32)
        if obn_scores.numel() != R:
            raise RuntimeError("obn_scores must have one entry per box")
        obn_out = torch.empty(R, dtype=torch.float32, device=boxes.device)
    _lib.call("nawsod_project_rois", _ptr(boxes), R, float(im_scale), -1.0 if flip_width is None else float(flip_width),
              int(batch_idx), _ptr(rois), _ptr(obn_scores), _ptr(obn_out), _stream())
    return rois if obn_scores is None else (rois, obn_out)


def dedup_rois(rois, dedup_boxes=1.0 / 16):
    """The dedup block of ``im_detect_bbox`` (core/test_wsl.py:125-133).  Returns device tensors
    (index [R], inv_index [R], num_unique [1], roi_offsets [2] = {0, num_unique})."""
    _req(rois, "rois", torch.float32, 2)
    if rois.shape[1] != 5:
        raise RuntimeError("rois must be [R,5]")
    R = rois.shape[0]
    new = lambda n: torch.empty(n, dtype=torch.int32, device=rois.device)
    index, inv, nu, offs = new(R), new(R), new(1), new(2)
    _lib.call("nawsod_dedup_rois", _ptr(rois), R, float(dedup_boxes), _ptr(index), _ptr(inv), _ptr(nu), _ptr(offs), _stream())
    return index, inv, nu, offs


def gather_rows(src, index, n=None):
    """dst[i, :] = src[index[i], :] for i < n (``rois[index, :]``, core/test_wsl.py:131-133)."""
    _req(src, "src", torch.float32)
    _req(index, "index", torch.int32, 1)
    n = index.numel() if n is None else int(n)
    cols = src.numel() // max(src.shape[0], 1)
    dst = torch.empty((n,) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    _lib.call("nawsod_gather_rows", _ptr(src), _ptr(index), n, cols, _ptr(dst), _stream())
    return dst


def scatter_scores(rois_pred, inv_index=None, *, R=None, out=None, accumulate=False):
    """``cls_prob = concat(rois_pred[:, :1], rois_pred)`` (modeling/wsl_heads.py:57-67) mapped back to the
    original boxes, ``scores[inv_index, :]`` (core/test_wsl.py:173-176); ``accumulate`` adds into ``out``
    (the running sum of the 'AVG' score heuristic, core/test_wsl.py:262-263)."""
    Ru, C, ld = _mat(rois_pred, "rois_pred")
    if rois_pred.dtype != torch.float32:
        raise RuntimeError("rois_pred must be float32")
    if inv_index is not None:
        _req(inv_index, "inv_index", torch.int32, 1)
        R = inv_index.numel()
    elif R is None:
        R = Ru
    if out is None:
        if accumulate:
            raise RuntimeError("scatter_scores: accumulate needs an existing out tensor")
        out = torch.empty((R, C + 1), dtype=torch.float32, device=rois_pred.device)
    _req(out, "out", torch.float32, 2)
    if tuple(out.shape) != (R, C + 1):
        raise RuntimeError("scatter_scores: out must be [R, C+1]")
    _lib.call("nawsod_scatter_scores", _ptr(rois_pred), ld, _ptr(inv_index), R, C, int(bool(accumulate)), _ptr(out), _stream())
    return out


def scores_finalize(acc, count):
    """Divide the accumulated scores by the number of passes (np.mean's final step)."""
    _req(acc, "acc", torch.float32)
    _lib.call("nawsod_scores_finalize", _ptr(acc), acc.numel(), int(count), _stream())
    return acc


def nms_and_limit(scores, boxes, *, score_thresh=0.05, nms_thresh=0.3, detections_per_im=100):
    """Device part of ``box_results_with_nms_and_limit`` (core/test_wsl.py:803-863): per-class threshold,
    greedy NMS (utils/cython_nms.pyx:38-93) and the detections-per-image limit.  scores [R, num_classes],
    boxes [R, 4].  Returns (keep [num_classes, R] uint8, num_keep [num_classes] int32, image_thresh [1])."""
    _req(scores, "scores", torch.float32, 2)
    _req(boxes, "boxes", torch.float32, 2)
    R, K1 = scores.shape
    if tuple(boxes.shape) != (R, 4):
        raise RuntimeError("boxes must be [R,4] (COORD_HEUR 'ID': one box per proposal)")
    keep = torch.empty((K1, R), dtype=torch.uint8, device=scores.device)
    num_keep = torch.empty(K1, dtype=torch.int32, device=scores.device)
    thr = torch.empty(1, dtype=torch.float32, device=scores.device)
    _lib.call("nawsod_nms_and_limit", _ptr(scores), _ptr(boxes), R, K1, float(score_thresh), float(nms_thresh),
              int(detections_per_im), _ptr(keep), _ptr(num_keep), _ptr(thr), _stream())
    return keep, num_keep, thr


def MinEntropyLoss(X, L):
    """``MinEntropyLoss([X, L] -> Y)`` (ops/min_entropy_loss_op.cu:70-104)."""
    _req(X, "X", torch.float32, 2)                    # CAFFE_ENFORCE_EQ(X.dim(), 2)
    _req(L, "L", torch.float32, 2)                    # CAFFE_ENFORCE_EQ(L.dim(), 2)
    if X.shape[1] != L.shape[1]:                      # CAFFE_ENFORCE_EQ(X.dim32(1), L.dim32(1))
        raise RuntimeError("MinEntropyLoss: X and L disagree on the number of classes")
    Y = torch.empty((), dtype=torch.float32, device=X.device)
    norm = torch.empty(1, dtype=torch.float32, device=X.device)
    _lib.call("nawsod_min_entropy_loss_fwd", _ptr(X), _ptr(L), X.shape[0], X.shape[1], L.shape[0], _ptr(Y), _ptr(norm), _stream())
    return Y


def MinEntropyLossGradient(X, L, dY):
    """``MinEntropyLossGradient([X, L, dY] -> dX)`` (ops/min_entropy_loss_op.cu:106-152)."""
    _req(X, "X", torch.float32, 2)
    _req(L, "L", torch.float32, 2)
    _req(dY, "dY", torch.float32)
    if X.shape[1] != L.shape[1]:
        raise RuntimeError("MinEntropyLoss: X and L disagree on the number of classes")
    if dY.numel() != 1:                               # CAFFE_ENFORCE_EQ(dY.numel(), 1)
        raise RuntimeError("dY must have one element")
    dX = torch.empty_like(X)
    norm = torch.empty(1, dtype=torch.float32, device=X.device)
    _lib.call("nawsod_min_entropy_loss_bwd", _ptr(X), _ptr(L), _ptr(dY), X.shape[0], X.shape[1], L.shape[0], _ptr(dX),
              _ptr(norm), _stream())
    return dX


# ---------------------------------------------------------------------------------------------
# the training-input contract (roi_data/wsl.py, roi_data/loader_wsl.py, tools/convert_mcg.py)
# ---------------------------------------------------------------------------------------------
def sample_rois(boxes, im_scale, im_crop, batch_idx=0, *, obn_scores=None, out_rois=None, out_obn=None):
    """``_sample_rois`` + ``_project_im_rois`` for one image (roi_data/wsl.py:101-111, 212-225): boxes [R,4]
    float32 in original-image pixels, ``im_crop`` = (x1, y1, x2, y2) integers (minibatch_wsl.py:63-64).
    Returns rois [R,5] (and obn_scores + 1 as [R,1] when given).  ``out_rois`` / ``out_obn``: row slices of the
    minibatch blobs to fill in place (add_wsl_blobs concatenates the images, roi_data/wsl.py:59-85)."""
    _req(boxes, "boxes", torch.float32, 2)
    if boxes.shape[1] != 4:
        raise RuntimeError("boxes must be [R,4]")
    R = boxes.shape[0]
    crop = [int(v) for v in im_crop]
    if len(crop) != 4:
        raise RuntimeError("im_crop must be (x1, y1, x2, y2)")
    rois = out_rois if out_rois is not None else torch.empty((R, 5), dtype=torch.float32, device=boxes.device)
    _req(rois, "rois", torch.float32, 2)
    if tuple(rois.shape) != (R, 5):
        raise RuntimeError("rois must be [R,5]")
    obn_out = None
    if obn_scores is not None:
        _req(obn_scores, "obn_scores", torch.float32)
        if obn_scores.numel() != R:
            raise RuntimeError("obn_scores must have one entry per box")
        obn_out = out_obn if out_obn is not None else torch.empty((R, 1), dtype=torch.float32, device=boxes.device)
        _req(obn_out, "obn_out", torch.float32)
        if obn_out.numel() != R:
            raise RuntimeError("obn_out must have one entry per box")
    _lib.call("nawsod_sample_rois", _ptr(boxes), R, float(im_scale), crop[0], crop[1], crop[2], crop[3], int(batch_idx),
              _ptr(rois), _ptr(obn_scores), _ptr(obn_out), _stream())
    return rois if obn_scores is None else (rois, obn_out)


def image_labels(gt_classes, num_classes, *, out_oh=None, out_int=None):
    """The label half of ``_sample_rois`` (roi_data/wsl.py:139-155): gt_classes [n] int32 (0 = proposal row).
    Returns (labels_oh [1, num_classes-1] float32, labels_int32 [1] int32: -1 if no ground-truth row)."""
    _req(gt_classes, "gt_classes", torch.int32, 1)
    dev = gt_classes.device
    oh = out_oh if out_oh is not None else torch.empty((1, num_classes - 1), dtype=torch.float32, device=dev)
    li = out_int if out_int is not None else torch.empty(1, dtype=torch.int32, device=dev)
    _req(oh, "labels_oh", torch.float32)
    _req(li, "labels_int32", torch.int32)
    if oh.numel() != num_classes - 1 or li.numel() != 1:
        raise RuntimeError("labels_oh must hold num_classes-1 entries and labels_int32 one")
    _lib.call("nawsod_image_labels", _ptr(gt_classes), gt_classes.numel(), int(num_classes), _ptr(oh), _ptr(li), _stream())
    return oh, li


def bagging_mixup(x, lam, out=None):
    """``out[0] = lam * x[0] + (1 - lam) * x[1]`` in float32 (roi_data/loader_wsl.py:152-164) for a blob whose
    leading dimension holds the two images (``data`` [2,3,H,W], ``labels_oh`` [2,C])."""
    _req(x, "x", torch.float32)
    if x.dim() < 1 or x.shape[0] != 2:
        raise RuntimeError("bagging_mixup mixes exactly two images (leading dimension 2)")
    out = out if out is not None else torch.empty((1,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    _req(out, "out", torch.float32)
    n = x[0].numel()
    if out.numel() != n:
        raise RuntimeError("out must have the shape of one image")
    import numpy as np
    l0, l1 = float(np.float32(lam)), float(np.float32(1.0 - float(lam)))     # the scalar takes the array's type (NumPy 1.x casting)
    _lib.call("nawsod_bagging_mixup", _ptr(x[0]), _ptr(x[1]), n, l0, l1, _ptr(out), _stream())
    return out


def set_column(a, col, value):
    """``a[:, col] = value`` in place (``blobs['rois'][:, 0] = 0``, roi_data/loader_wsl.py:165)."""
    _req(a, "a", torch.float32, 2)
    _lib.call("nawsod_set_column", _ptr(a), a.shape[0], a.shape[1], int(col), float(value), _stream())
    return a


def convert_mcg_boxes(bboxes):
    """tools/convert_mcg.py:45-49: 1-indexed (y1,x1,y2,x2) float64 .mat boxes -> 0-indexed (x1,y1,x2,y2).
    Returns an int16 tensor holding the uint16 bit patterns (``.view(torch.uint16)`` / NumPy ``view('uint16')``)."""
    _req(bboxes, "bboxes", torch.float64, 2)
    if bboxes.shape[1] != 4:
        raise RuntimeError("bboxes must be [R,4]")
    out = torch.empty((bboxes.shape[0], 4), dtype=torch.int16, device=bboxes.device)
    _lib.call("nawsod_convert_mcg_boxes", _ptr(bboxes), bboxes.shape[0], _ptr(out), _stream())
    return out


# --------------------------------------------------------------------------------------------
# N4: the frozen VGG16 conv body in channels-last bf16 (EXPERIMENTAL -- see csrc/conv_body.cu)
# --------------------------------------------------------------------------------------------
def _nhwc_bf16(X, name):
    _req(X, name, torch.bfloat16, 4)
    if not X.is_contiguous():
        raise RuntimeError("%s must be a contiguous channels-last [N,H,W,C] tensor" % name)
    return X.shape


def Im2Col3x3(X, *, dilation=1, out=None):
    """Patch matrix of a 3x3 / stride 1 / pad = dilation convolution: X [N,H,W,C] bf16 -> [N*H*W, 9*C] bf16 with the
    K-order (kh, kw, c); taps outside the image are zeros (Caffe2 ``Conv`` zero padding, modeling/VGG16.py:10-56)."""
    N, H, W, C = _nhwc_bf16(X, "X")
    cols = torch.empty((N * H * W, 9 * C), dtype=torch.bfloat16, device=X.device) if out is None else out
    _req(cols, "out", torch.bfloat16, 2)
    if tuple(cols.shape) != (N * H * W, 9 * C) or not cols.is_contiguous():
        raise RuntimeError("Im2Col3x3: out must be a contiguous [%d, %d] matrix" % (N * H * W, 9 * C))
    _lib.call("nawsod_im2col3x3", _ptr(X), N, H, W, C, int(dilation), _ptr(cols), _stream())
    return cols


def Conv3x3Relu(X, Wmat, b, *, dilation=1, relu=True, cols=None, implicit=None):
    """``Conv(kernel=3, pad=dilation, stride=1, dilation)`` + ``Relu`` on a channels-last bf16 map; ``Wmat`` [Cout, 9*Cin]
    is the reference's [Cout, Cin, 3, 3] weight permuted to (kh, kw, c).  Returns Y [N,H,W,Cout] bf16 -- the next layer's
    input as is.

    ``implicit`` (default: whenever Cin is a multiple of 64): the implicit GEMM of csrc/conv_body.cu -- shifted 4-D TMA
    boxes of the map are the A operand, no patch matrix.  Otherwise (conv1_1's 8 padded planes, or ``implicit=False``):
    the patch matrix (``Im2Col3x3``) times ``Wmat`` on the FC GEMM.  Both accumulate the same products in the same
    order: bit-identical outputs."""
    N, H, W, C = _nhwc_bf16(X, "X")
    if Wmat.dim() != 2 or Wmat.shape[1] != 9 * C:
        raise RuntimeError("Conv3x3Relu: Wmat must be [Cout, %d], got %s" % (9 * C, tuple(Wmat.shape)))
    Cout = Wmat.shape[0]
    if implicit is None:
        implicit = C % 64 == 0 and Cout % 32 == 0
    if implicit:
        _req(Wmat, "Wmat", torch.bfloat16, 2)
        _req(b, "b", torch.float32, 1)
        if b.numel() != Cout or not Wmat.is_contiguous():
            raise RuntimeError("Conv3x3Relu: need a contiguous Wmat and a bias of %d elements" % Cout)
        Y = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=X.device)
        _lib.call("nawsod_conv3x3_relu", _ptr(X), N, H, W, C, _ptr(Wmat), _ptr(b), Cout, int(dilation), 1 if relu else 0, _ptr(Y), _stream())
        return Y
    cols = Im2Col3x3(X, dilation=dilation, out=cols)
    Y = FC(cols, Wmat, b, relu=relu, out_dtype=torch.bfloat16)
    return Y.view(N, H, W, Cout)


def MaxPool2x2(X, *, stride=2):
    """``MaxPool(kernel=2, pad=0, stride)`` on a channels-last bf16 map (Caffe2 floor output size)."""
    N, H, W, C = _nhwc_bf16(X, "X")
    if H < 2 or W < 2:
        raise RuntimeError("MaxPool2x2: the map must be at least 2 x 2")
    Y = torch.empty((N, (H - 2) // stride + 1, (W - 2) // stride + 1, C), dtype=torch.bfloat16, device=X.device)
    _lib.call("nawsod_maxpool2x2", _ptr(X), N, H, W, C, int(stride), _ptr(Y), _stream())
    return Y
