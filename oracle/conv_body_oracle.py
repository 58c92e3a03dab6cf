"""CPU oracle for the frozen VGG16 conv body (second half).  TEST INFRASTRUCTURE ONLY: imported by
``tests/`` (and nothing under ``na-fwebsod_b200/``).

Restates detectron/modeling/VGG16.py:9-58 (``add_VGG16_conv5_body_origin``): thirteen ``Conv(3x3, stride 1, pad = dilation)``
+ ``Relu`` pairs with ``MaxPool(kernel 2, pad 0)`` after groups 1-4; with ``WSL.DILATION == 2`` pool4 has stride 1 and the
conv5 group pad 2 / dilation 2 (:39-48), spatial scale 1/8; otherwise stride 2 / pad 1 / 1/16 (:49-58).  The arithmetic of
Caffe2's Conv / Relu / MaxPool (pytorch v1.3.0, not in the tree) is float32 cross-correlation with zero padding and a
floor-mode pooling window -- evaluated here with torch's CPU float32 ``conv2d`` / ``max_pool2d`` (a floating-point
kernel: the torch fp32 reference is the checker, tolerance stated in the tests).

Parity status: the OPERATOR SEQUENCE (which op on which blob with which arguments) is pinned by the reference's own
builder run on a tracing model (tests/golden/make_golden_vgg16_body.py -> tests/golden/vgg16_body.npz); the arithmetic
of the Caffe2 built-ins is a restatement (the reference has no tests for them).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

GROUPS = [
    [("conv1_1", 3, 64), ("conv1_2", 64, 64)],
    [("conv2_1", 64, 128), ("conv2_2", 128, 128)],
    [("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256)],
    [("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512)],
    [("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)],
]


def synth_params(seed=0):
    """He-scaled Gaussian weights (activations keep an O(1) scale through thirteen layers) and small biases, regenerated
    from the seed wherever they are needed (14.7 M parameters are not stored in the golden file)."""
    rng = np.random.default_rng(seed)
    params = {}
    for group in GROUPS:
        for name, cin, cout in group:
            params[name + "_w"] = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin))).astype(F32)
            params[name + "_b"] = (rng.standard_normal(cout) * 0.05).astype(F32)
    return params


def param_checksum(params):
    return np.array([float(np.asarray(params[k], np.float64).sum()) for k in sorted(params)])


def op_sequence(dilation=2):
    """(type, in, out, args) in the reference builder's order (VGG16.py:10-56; Relu in place)."""
    seq, blob = [], "data"
    for gi, group in enumerate(GROUPS):
        last = gi == 4
        d = 2 if (last and dilation == 2) else 1
        for name, cin, cout in group:
            seq.append(("Conv", blob, name, dict(dim_in=cin, dim_out=cout, kernel=3, pad=d, stride=1, dilation=d)))
            seq.append(("Relu", name, name, {}))
            blob = name
        if not last:
            seq.append(("MaxPool", blob, "pool%d" % (gi + 1), dict(kernel=2, pad=0, stride=1 if (gi == 3 and dilation == 2) else 2)))
            blob = "pool%d" % (gi + 1)
    return seq


def run_op(kind, x, args, w=None, b=None):
    """One Caffe2 operator on a float32 NCHW array (torch CPU float32)."""
    import torch
    import torch.nn.functional as Fn
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=F32))
    if kind == "Conv":
        y = Fn.conv2d(t, torch.from_numpy(np.ascontiguousarray(w, dtype=F32)), torch.from_numpy(np.ascontiguousarray(b, dtype=F32)),
                      stride=args.get("stride", 1), padding=args.get("pad", 0), dilation=args.get("dilation", 1))
    elif kind == "Relu":
        y = torch.relu(t)
    elif kind == "MaxPool":
        y = Fn.max_pool2d(t, kernel_size=args["kernel"], stride=args["stride"], padding=args.get("pad", 0), ceil_mode=False)
    else:
        raise ValueError(kind)
    return y.numpy()


def conv5_body(data, params, dilation=2, round_bf16=False, keep=()):
    """data [N,3,H,W] float32 -> (conv5_3 [N,512,h,w] float32, 512, spatial_scale, kept blobs).  ``round_bf16`` evaluates
    the same function on what the bf16 product path stores: input, weights and every layer output rounded to bf16
    (accumulation stays float32, as on the tensor cores)."""
    import torch
    rb = (lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F32)).to(torch.bfloat16).float().numpy()) if round_bf16 else (lambda a: a)
    ws = {"data": rb(np.asarray(data, F32))}
    kept = {}
    for kind, src, dst, args in op_sequence(dilation):
        if kind == "Conv":
            ws[dst] = run_op(kind, ws[src], args, rb(params[dst + "_w"]), params[dst + "_b"])
        elif kind == "Relu":
            ws[dst] = rb(run_op(kind, ws[src], args))          # the product stores the post-ReLU activation in bf16
        else:
            ws[dst] = run_op(kind, ws[src], args)
        if dst in keep:
            kept[dst] = ws[dst]
    return ws["conv5_3"], 512, (1.0 / 8.0 if dilation == 2 else 1.0 / 16.0), kept
