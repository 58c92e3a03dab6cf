"""ctypes binding of libnawsod.so (the C ABI declared in include/nawsod.h).

The library is the product; this module only marshals pointers.  If the shared object is
missing or a call fails, a RuntimeError is raised -- there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnawsod.so")

F32, BF16 = 0, 1
NCHW, NHWC = 0, 1
FC_RELU, FC_DROPOUT, FC_ACCUMULATE, FC_ROUND_TF32 = 1, 2, 4, 8
MIL_ENTROPY, MIL_MEAN, MIL_BACKWARD = 1, 2, 4

_c = ctypes
_vp, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

# name -> (restype, argtypes); must list every symbol include/nawsod.h declares
PROTOTYPES = {
    "nawsod_last_error": (_c.c_char_p, []),
    "nawsod_version": (_i, []),
    "nawsod_set_tuning": (_i, [_c.c_char_p, _i64]),
    "nawsod_transpose_batched": (_i, [_vp, _i64, _i64, _i64, _vp, _i, _vp]),
    "nawsod_roi_pool_f_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _i, _i, _vp, _vp]),
    "nawsod_roi_pool_f_bwd": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "nawsod_roi_feature_boost": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "nawsod_fc_fwd": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _c.c_uint64, _i, _i, _i, _i, _vp, _i64, _i, _i, _vp]),
    "nawsod_fc_fwd_gated": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _c.c_uint64, _i, _i, _i, _i, _vp, _i64, _i, _i,
                                 _vp, _i, _i, _i, _c.c_uint32, _i64, _vp, _vp]),
    "nawsod_fc_bwd_x": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _i, _i, _vp]),
    "nawsod_fc_bwd_w": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _vp, _i, _vp]),
    "nawsod_fc_fwd_stacks": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _i64, _c.c_uint64, _i, _i, _i, _i, _i,
                                  _vp, _i64, _i64, _i, _i, _vp]),
    "nawsod_fc_bwd_x_stacks": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i, _vp, _i64, _i64, _i, _i, _i, _i, _i,
                                    _vp, _i64, _i64, _i, _i, _vp]),
    "nawsod_fc_bwd_w_stacks": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _i64, _i64, _vp, _i64, _i, _vp]),
    "nawsod_fc_bias_grad": (_i, [_vp, _i64, _i64, _i, _i, _i, _i, _vp, _i64, _i, _vp]),
    "nawsod_convert_f32_to_bf16": (_i, [_vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "nawsod_mil_workspace_bytes": (_i64, [_i, _i, _i]),
    "nawsod_mil_head_fwd_bwd": (_i, [_vp] * 4 + [_i64] + [_vp] * 3 + [_i, _i, _i, _i] + [_vp] * 11 + [_i64, _vp, _vp]),
    "nawsod_roi_iou": (_i, [_vp, _i, _vp, _vp]),
    "nawsod_cross_entropy_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "nawsod_cross_entropy_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "nawsod_round_to_tf32": (_i, [_vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "nawsod_split_tf32": (_i, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp]),
    "nawsod_sgd_update": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _i, _i, _i64, _vp, _i, _vp]),
    "nawsod_sgd_update_reduce": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _f, _f, _f, _i, _i64, _vp, _i, _vp, _vp]),
    "nawsod_p2p_enable_peer_access": (_i, [_i]),
    "nawsod_p2p_get_mem_handle": (_i, [_vp, _c.c_char_p, _i64, _c.POINTER(_i64)]),
    "nawsod_p2p_open_mem_handle": (_i, [_c.c_char_p, _i64, _c.POINTER(_vp)]),
    "nawsod_p2p_close_mem_handle": (_i, [_vp]),
    "nawsod_p2p_copy": (_i, [_vp, _vp, _i64, _vp]),
    "nawsod_p2p_signal": (_i, [_vp, _i, _c.c_uint32, _vp]),
    "nawsod_p2p_scatter": (_i, [_vp, _vp, _i, _i64, _vp, _i, _c.c_uint32, _i, _vp]),
    "nawsod_p2p_wait": (_i, [_vp, _i, _c.c_uint32, _i64, _vp, _vp]),
    "nawsod_project_rois": (_i, [_vp, _i, _c.c_double, _c.c_double, _i, _vp, _vp, _vp, _vp]),
    "nawsod_dedup_rois": (_i, [_vp, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "nawsod_gather_rows": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "nawsod_scatter_scores": (_i, [_vp, _i64, _vp, _i, _i, _i, _vp, _vp]),
    "nawsod_scores_finalize": (_i, [_vp, _i64, _i, _vp]),
    "nawsod_nms_and_limit": (_i, [_vp, _vp, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "nawsod_min_entropy_loss_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "nawsod_min_entropy_loss_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "nawsod_sample_rois": (_i, [_vp, _i, _c.c_double, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "nawsod_image_labels": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "nawsod_bagging_mixup": (_i, [_vp, _vp, _i64, _f, _f, _vp, _vp]),
    "nawsod_set_column": (_i, [_vp, _i, _i64, _i, _f, _vp]),
    "nawsod_convert_mcg_boxes": (_i, [_vp, _i, _vp, _vp]),
    "nawsod_conv3x3_relu": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "nawsod_im2col3x3": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "nawsod_maxpool2x2": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
}

# kernels launched per C-ABI call (bench.py reports the count of OUR kernels in the timed region)
KERNELS_PER_CALL = {
    "nawsod_transpose_batched": 1, "nawsod_roi_pool_f_fwd": 1, "nawsod_roi_pool_f_bwd": 1, "nawsod_roi_feature_boost": 1,
    "nawsod_fc_fwd": 1, "nawsod_fc_fwd_gated": 1, "nawsod_fc_bwd_x": 1, "nawsod_fc_bwd_w": 1, "nawsod_fc_fwd_stacks": 1, "nawsod_fc_bwd_x_stacks": 1,
    "nawsod_fc_bwd_w_stacks": 1, "nawsod_convert_f32_to_bf16": 1,
    "nawsod_round_to_tf32": 1, "nawsod_split_tf32": 1, "nawsod_mil_head_fwd_bwd": 1, "nawsod_roi_iou": 1, "nawsod_cross_entropy_fwd": 1,
    "nawsod_cross_entropy_bwd": 1, "nawsod_sgd_update": 1, "nawsod_sgd_update_reduce": 1, "nawsod_p2p_signal": 1, "nawsod_p2p_scatter": 1,
    "nawsod_p2p_wait": 1,
    "nawsod_project_rois": 1, "nawsod_dedup_rois": 1, "nawsod_gather_rows": 1, "nawsod_scatter_scores": 1,
    "nawsod_scores_finalize": 1, "nawsod_nms_and_limit": 2, "nawsod_min_entropy_loss_fwd": 1, "nawsod_min_entropy_loss_bwd": 2,
    "nawsod_sample_rois": 1, "nawsod_image_labels": 1, "nawsod_bagging_mixup": 1, "nawsod_set_column": 1, "nawsod_convert_mcg_boxes": 1,
    "nawsod_conv3x3_relu": 1, "nawsod_im2col3x3": 1, "nawsod_maxpool2x2": 1,
}
launch_count = 0

_lib = None


def load():
    """dlopen libnawsod.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libnawsod.so is not built (%s): run `python na-fwebsod_b200/build.py` or "
                "__graft_entry__.build(); there is no fallback path" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        # NAWSOD_TUNING=key=value[,key=value...]: tuning knobs applied once at load time (measurement / test runs)
        for kv in filter(None, os.environ.get("NAWSOD_TUNING", "").split(",")):
            k, v = kv.split("=")
            if lib.nawsod_set_tuning(k.strip().encode(), int(v)) != 0:
                raise RuntimeError("NAWSOD_TUNING: %s" % lib.nawsod_last_error().decode())
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().nawsod_last_error()
        raise RuntimeError("libnawsod error %d: %s" % (rc, msg.decode() if msg else "?"))


def call(name: str, *args, extra_kernels: int = 0):
    global launch_count
    check(getattr(load(), name)(*args))
    launch_count += KERNELS_PER_CALL.get(name, 0) + extra_kernels


def set_tuning(key: str, value: int):
    call("nawsod_set_tuning", key.encode(), int(value))
