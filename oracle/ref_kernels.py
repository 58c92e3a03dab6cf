"""ctypes front-end to oracle/_ref/libnawsod_ref_kernels.so: the reference's own CUDA kernels for RoIIoU, the in-tree
RoI max-pooling clone and MinEntropyLoss, executed on the host (oracle/build_ref_kernels.py, oracle/cuda_host_shim.h).
TEST INFRASTRUCTURE ONLY (see the header of oracle/nawsod_oracle.py)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libnawsod_ref_kernels.so")
_lib = None
_f, _i = ctypes.c_float, ctypes.c_int


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def roi_iou(rois):
    """``iou<float>`` of detectron/ops/roi_iou_op.cu:28-62 on rois [n, 5]."""
    r = np.ascontiguousarray(rois, np.float32)
    n = r.shape[0]
    J = np.empty((n, n), np.float32)
    _load().nawsod_refk_roi_iou(_p(r), _i(n), _p(J))
    return J


def rois9(rois5, inner=None):
    """[R,5] -> the clone's [R,9] rois; ``inner`` None disables the inner rectangle (no cell satisfies
    start_in < h < end_in when start_in >= end_in), which is RoIPoolF's behaviour."""
    r = np.asarray(rois5, np.float32)
    inn = np.tile(np.asarray([[1e6, 1e6, -1e6, -1e6]], np.float32), (r.shape[0], 1)) if inner is None else np.asarray(inner, np.float32)
    return np.ascontiguousarray(np.concatenate([r, inn], axis=1))


def roi_loop_pool(X, r9, spatial_scale, pooled_h=7, pooled_w=7):
    """``ROIPoolForward<float>`` of detectron/ops/roi_loop_pool_op.cu:19-102 (NCHW, rois [R,9])."""
    X = np.ascontiguousarray(X, np.float32)
    r9 = np.ascontiguousarray(r9, np.float32)
    N, C, H, W = X.shape
    R = r9.shape[0]
    Y = np.empty((R, C, pooled_h, pooled_w), np.float32)
    A = np.empty((R, C, pooled_h, pooled_w), np.int32)
    _load().nawsod_refk_roi_loop_pool_fwd(_p(X), _p(r9), _i(R), _i(C), _i(H), _i(W), _i(pooled_h), _i(pooled_w),
                                          _f(spatial_scale), _p(Y), _p(A))
    return Y, A


def roi_loop_pool_grad(X_shape, r9, argmax, dY, spatial_scale):
    """Zero-fill + ``ROIPoolBackward<float>`` (detectron/ops/roi_loop_pool_op.cu:105-140,199-201)."""
    N, C, H, W = X_shape
    r9 = np.ascontiguousarray(r9, np.float32)
    A = np.ascontiguousarray(argmax, np.int32)
    dY = np.ascontiguousarray(dY, np.float32)
    R, _, PH, PW = dY.shape
    dX = np.empty(X_shape, np.float32)
    _load().nawsod_refk_roi_loop_pool_bwd(_p(dY), _p(A), _p(r9), _i(R), _i(N), _i(C), _i(H), _i(W), _i(PH), _i(PW),
                                          _f(spatial_scale), _p(dX))
    return dX


def min_entropy_forward_kernel(X, L, log_threshold=1e-20):
    """``Forward<float>`` (detectron/ops/min_entropy_loss_op.cu:34-49): (sum of -p log p, count) over the classes
    with L[0, c] >= 0.5, both accumulated from zero."""
    X = np.ascontiguousarray(X, np.float32)
    L = np.ascontiguousarray(L, np.float32)
    N, C = X.shape
    Y, norm = np.zeros(1, np.float32), np.zeros(1, np.float32)
    _load().nawsod_refk_min_entropy_fwd(_p(X), _p(L), _i(N), _i(C), _i(L.shape[0]), _f(log_threshold), _p(Y), _p(norm))
    return Y[0], norm[0]


def min_entropy_backward_kernel(X, L, scale, log_threshold=1e-20, diff_threshold=1e4):
    """Zero-fill + ``Backward<float>`` (detectron/ops/min_entropy_loss_op.cu:52-66)."""
    X = np.ascontiguousarray(X, np.float32)
    L = np.ascontiguousarray(L, np.float32)
    N, C = X.shape
    s = np.asarray([scale], np.float32)
    dX = np.empty((N, C), np.float32)
    _load().nawsod_refk_min_entropy_bwd(_p(X), _p(L), _i(N), _i(C), _i(L.shape[0]), _p(s), _f(log_threshold),
                                        _f(diff_threshold), _p(dX))
    return dX
