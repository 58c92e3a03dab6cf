"""ctypes front-end to oracle/_ref/libnawsod_ref.so -- the reference's OWN CPU operators
(RoIFeatureBoost[Gradient], [Weighted]CrossEntropyWithLogits[Gradient],
ACMWeightDecayMomentumSGDUpdate) compiled unmodified from /root/reference by
oracle/build_ref.sh.  TEST INFRASTRUCTURE ONLY (see oracle/nawsod_oracle.py header)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libnawsod_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_PATH)
        lib.nawsod_ref_create.restype = ctypes.c_void_p
        lib.nawsod_ref_create.argtypes = [ctypes.c_char_p, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double)]
        lib.nawsod_ref_destroy.argtypes = [ctypes.c_void_p]
        lib.nawsod_ref_last_error.restype = ctypes.c_char_p
        lib.nawsod_ref_run.restype = ctypes.c_int
        lib.nawsod_ref_has_op.argtypes = [ctypes.c_char_p]
        _lib = lib
    return _lib


class RefOp:
    """One instance of a reference operator (Caffe2 ``core.CreateOperator`` analogue)."""

    def __init__(self, op_type: str, **args):
        lib = _load()
        names = (ctypes.c_char_p * len(args))(*[k.encode() for k in args])
        vals = (ctypes.c_double * len(args))(*[float(v) for v in args.values()])
        self._h = lib.nawsod_ref_create(op_type.encode(), len(args), names, vals)
        if not self._h:
            raise RuntimeError(lib.nawsod_ref_last_error().decode())
        self.type = op_type

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.nawsod_ref_destroy(self._h)
            self._h = None

    def run(self, inputs, n_out, out_alias=None, out_cap=None):
        """inputs: list of float32 arrays.  out_alias[k] = index of the input that output k
        is in-place with (or -1).  Returns the list of output arrays."""
        lib = _load()
        ins = [np.ascontiguousarray(a, dtype=np.float32) for a in inputs]
        n_in = len(ins)
        in_ptrs = (ctypes.POINTER(ctypes.c_float) * n_in)(
            *[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in ins])
        in_ndims = (ctypes.c_int * n_in)(*[a.ndim for a in ins])
        dims = [d for a in ins for d in a.shape]
        in_dims = (ctypes.c_int64 * max(len(dims), 1))(*dims)
        out_alias = list(out_alias) if out_alias is not None else [-1] * n_out
        cap = max([a.size for a in ins] + [1]) if out_cap is None else out_cap
        outs = [np.zeros(cap, dtype=np.float32) for _ in range(n_out)]
        out_ptrs = (ctypes.POINTER(ctypes.c_float) * n_out)(
            *[o.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for o in outs])
        out_cap_a = (ctypes.c_int64 * n_out)(*[cap] * n_out)
        out_ndims = (ctypes.c_int * n_out)()
        out_dims = (ctypes.c_int64 * (8 * n_out))()
        rc = lib.nawsod_ref_run(ctypes.c_void_p(self._h), n_in, in_ptrs, in_ndims, in_dims, n_out,
                                (ctypes.c_int * n_out)(*out_alias), out_ptrs, out_cap_a, out_ndims, out_dims)
        if rc != 0:
            raise RuntimeError(lib.nawsod_ref_last_error().decode())
        res = []
        for k in range(n_out):
            shape = tuple(out_dims[k * 8 + j] for j in range(out_ndims[k]))
            n = int(np.prod(shape)) if shape else 1
            res.append(outs[k][:n].reshape(shape).copy())
        return res


def has_op(op_type: str) -> bool:
    return bool(_load().nawsod_ref_has_op(op_type.encode()))
