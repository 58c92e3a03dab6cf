"""ctypes front-end to the C restatement of RoIPoolF (oracle/roi_pool_ref.c).
TEST INFRASTRUCTURE ONLY (see oracle/nawsod_oracle.py header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liboracle_roi_pool.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["bash", os.path.join(_HERE, "build_oracle.sh")])
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def roi_pool_f(X, rois, spatial_scale, pooled_h=7, pooled_w=7, want_argmax=True):
    lib = _load()
    X = np.ascontiguousarray(X, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    N, C, H, W = X.shape
    R = rois.shape[0]
    Y = np.empty((R, C, pooled_h, pooled_w), np.float32)
    A = np.empty((R, C, pooled_h, pooled_w), np.int32) if want_argmax else None
    lib.nawsod_oracle_roi_pool_f(_p(X, ctypes.c_float), _p(rois, ctypes.c_float), N, C, H, W, R,
                                 ctypes.c_float(spatial_scale), pooled_h, pooled_w, _p(Y, ctypes.c_float),
                                 _p(A, ctypes.c_int32) if want_argmax else None)
    return (Y, A) if want_argmax else Y


def roi_pool_f_grad(X_shape, rois, argmax, dY):
    lib = _load()
    N, C, H, W = X_shape
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    R = rois.shape[0]
    dY = np.ascontiguousarray(dY, dtype=np.float32)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    PH, PW = dY.shape[-2], dY.shape[-1]
    dX = np.empty((N, C, H, W), np.float32)
    lib.nawsod_oracle_roi_pool_f_grad(_p(dY, ctypes.c_float), _p(argmax, ctypes.c_int32), _p(rois, ctypes.c_float),
                                      N, C, H, W, R, PH, PW, _p(dX, ctypes.c_float))
    return dX
