"""The frozen VGG16 conv body in channels-last bf16 (na-fwebsod_b200/csrc/conv_body.cu + conv_body.py):
3x3 convolutions as implicit GEMMs on the tcgen05 tensor cores (shifted 4-D TMA boxes of the map, no patch matrix), the
patch-matrix form for conv1_1, the 2x2 max-pools.

Parity bars: the patch matrix and the max-pool move / compare bf16 values -> bit-exact against NumPy on the same bf16
inputs; the implicit GEMM is bit-identical to the patch matrix times the same weights on the FC GEMM (same products, same
accumulation order); a convolution and the whole body against the oracle evaluated on the same bf16-rounded inputs and
weights (oracle.conv_body_oracle, torch CPU float32): relative L2 <= 1e-2 per convolution (the bf16 bar of north_star); the
whole thirteen-layer body <= 1.5e-2 against the oracle on the same bf16 inputs and <= 2e-2 against the float32 run of the
reference's builder (tests/golden/vgg16_body.npz) -- the measured noise floor of bf16 storage through thirteen layers is
5e-3 / 7e-3 (see the comment in test_body_vs_reference_builder_run)."""
import os

import numpy as np
import pytest
import torch

from oracle import conv_body_oracle as CB

pytestmark = pytest.mark.gpu


def _ops():
    from nafwebsod_b200 import ops
    return ops


def _bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("shape,dil", [((2, 6, 7, 8), 1), ((1, 13, 9, 64), 2), ((2, 37, 50, 128), 1), ((1, 5, 5, 512), 2)])
def test_im2col_bit_exact(shape, dil):
    ops = _ops()
    N, H, W, C = shape
    x = _bf16(np.random.default_rng(H).standard_normal(shape))
    cols = ops.Im2Col3x3(x.cuda(), dilation=dil).float().cpu().numpy()
    xf = x.float().numpy()
    pad = np.pad(xf, ((0, 0), (dil, dil), (dil, dil), (0, 0)))
    want = np.stack([pad[:, kh * dil:kh * dil + H, kw * dil:kw * dil + W, :] for kh in range(3) for kw in range(3)], axis=3)
    assert np.array_equal(cols, want.reshape(N * H * W, 9 * C))


@pytest.mark.parametrize("shape,stride", [((2, 6, 7, 8), 2), ((1, 13, 9, 64), 1), ((2, 38, 50, 512), 2), ((1, 2, 2, 8), 2), ((1, 75, 101, 256), 1)])
def test_maxpool_bit_exact(shape, stride):
    ops = _ops()
    x = _bf16(np.random.default_rng(shape[1]).standard_normal(shape))
    y = ops.MaxPool2x2(x.cuda(), stride=stride).float().cpu().numpy()
    want = CB.run_op("MaxPool", x.float().numpy().transpose(0, 3, 1, 2), dict(kernel=2, pad=0, stride=stride)).transpose(0, 2, 3, 1)
    assert y.shape == want.shape and np.array_equal(y, want)


def test_errors():
    ops = _ops()
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.Im2Col3x3(torch.zeros((1, 4, 4, 3), dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(RuntimeError, match="dilation"):
        ops.Im2Col3x3(torch.zeros((1, 4, 4, 8), dtype=torch.bfloat16, device="cuda"), dilation=3)
    with pytest.raises(RuntimeError, match="stride"):
        ops.MaxPool2x2(torch.zeros((1, 4, 4, 8), dtype=torch.bfloat16, device="cuda"), stride=3)
    with pytest.raises(RuntimeError):
        ops.Im2Col3x3(torch.zeros((1, 4, 4, 8), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("shape,cout,dil", [((1, 19, 23, 64), 128, 1), ((2, 8, 16, 64), 64, 1), ((1, 37, 50, 512), 512, 2),
                                            ((1, 60, 80, 256), 512, 1), ((2, 9, 17, 128), 256, 2), ((1, 3, 5, 64), 32, 1)])
def test_implicit_gemm_equals_patch_matrix_gemm(shape, cout, dil):
    """Whole and ragged 8 x 16 pixel tiles, every tile width (BN = 64 / 128 / 256), both dilations, two images."""
    ops = _ops()
    N, H, W, cin = shape
    rng = np.random.default_rng(H * W + cout)
    x = _bf16(np.maximum(rng.standard_normal(shape), 0)).cuda()
    wm = _bf16(rng.standard_normal((cout, 9 * cin)) * np.sqrt(2.0 / (9 * cin))).cuda()
    b = torch.from_numpy((rng.standard_normal(cout) * 0.05).astype(np.float32)).cuda()
    for relu in (True, False):
        y0 = ops.Conv3x3Relu(x, wm, b, dilation=dil, relu=relu, implicit=False)
        y1 = ops.Conv3x3Relu(x, wm, b, dilation=dil, relu=relu, implicit=True)
        assert y1.shape == y0.shape and torch.equal(y0.view(torch.int16), y1.view(torch.int16))
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.Conv3x3Relu(x[..., :8].contiguous(), wm[:, :72].contiguous(), b, implicit=True)


@pytest.mark.parametrize("cin,cout,dil", [(64, 128, 1), (512, 512, 2), (8, 64, 1)])
def test_conv3x3_relu_vs_oracle(cin, cout, dil):
    ops = _ops()
    rng = np.random.default_rng(cin + dil)
    x = _bf16(np.maximum(rng.standard_normal((1, cin, 19, 23)), 0))                     # post-ReLU input, NCHW
    w = _bf16(rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (9 * cin)))
    b = (rng.standard_normal(cout) * 0.05).astype(np.float32)
    want = CB.run_op("Relu", CB.run_op("Conv", x.float().numpy(), dict(pad=dil, dilation=dil), w.float().numpy(), b), {})
    wm = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous().cuda()
    y = ops.Conv3x3Relu(x.permute(0, 2, 3, 1).contiguous().cuda(), wm, torch.from_numpy(b).cuda(), dilation=dil)
    assert rel_l2(y.float().cpu().numpy().transpose(0, 3, 1, 2), want) <= 1e-2        # only the bf16 rounding of the output


@pytest.mark.parametrize("tag,dil", [("d2", 2), ("d1", 1)])
def test_body_vs_reference_builder_run(golden_dir, tag, dil):
    from nafwebsod_b200.conv_body import VGG16ConvBody, add_VGG16_conv5_body_origin
    g = np.load(os.path.join(golden_dir, "vgg16_body.npz"))
    params = CB.synth_params(int(g["seed"]))
    body = VGG16ConvBody(dilation=dil)
    body.load_reference_params(params)
    body.feed_image(torch.from_numpy(g["data"]).cuda())
    y, dim, scale = add_VGG16_conv5_body_origin(body)
    torch.cuda.synchronize()
    assert (dim, scale) == (int(g[tag + "_dim_out"]), float(g[tag + "_spatial_scale"]))
    got = y.float().cpu().numpy().transpose(0, 3, 1, 2)
    same_inputs, _, _, _ = CB.conv5_body(g["data"], params, dil, round_bf16=True)
    assert got.shape == g[tag + "_conv5_3"].shape
    # Thirteen layers of bf16 storage amplify fp32 summation-order differences: two CPU evaluations of the SAME bf16 function
    # that differ only in the order conv2d adds its products are 5e-3 apart at conv5_3, and bf16 storage
    # itself costs 7e-3 against float32 (tests/test_conv_body_oracle.py).  Single layers are held to 1e-2 above.
    assert rel_l2(got, same_inputs) <= 1.5e-2                    # the same function on the same bf16 inputs
    assert rel_l2(got, g[tag + "_conv5_3"]) <= 2e-2             # end to end against the float32 run of the reference's builder


def test_body_feeds_the_head_without_a_layout_change(golden_dir):
    """conv5_3 comes out channels-last bf16 -- exactly what WeblyHeadModel.FeedBlobs(x_layout='NHWC') and RoIPoolF consume."""
    ops = _ops()
    from nafwebsod_b200.conv_body import VGG16ConvBody
    g = np.load(os.path.join(golden_dir, "vgg16_body.npz"))
    body = VGG16ConvBody(dilation=2)
    body.load_reference_params(CB.synth_params(int(g["seed"])))
    body.feed_image(torch.from_numpy(g["data"]).cuda())
    y, dim, scale = body.run()
    rois = torch.tensor([[0, 0, 0, 47, 31], [0, 8, 4, 40, 30]], dtype=torch.float32, device="cuda")
    Y, A = ops.RoIPoolF(y, rois, spatial_scale=scale, x_layout="NHWC", y_layout="NHWC")
    assert tuple(Y.shape) == (2, 7, 7, dim) and Y.dtype == torch.bfloat16
    assert float(Y.float().max()) == float(y[0, :4, :6].float().max())       # the first RoI covers the whole map
