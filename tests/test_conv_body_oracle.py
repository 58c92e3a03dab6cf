"""CPU tests of the frozen VGG16 conv body: the oracle and the product's layer list against what
the reference's own builder emits (tests/golden/vgg16_body.npz, made by tests/golden/make_golden_vgg16_body.py from
detectron/modeling/VGG16.py:9-58 with the shipped flickr_voc config)."""
import os

import numpy as np
import pytest

from oracle import conv_body_oracle as CB


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "vgg16_body.npz"))


def _normalise(seq):
    """(type, in, out, sorted args) with the reference's implicit defaults spelled out (no `dilation` argument = 1)."""
    out = []
    for kind, src, dst, args in seq:
        a = dict(args)
        if kind == "Conv":
            a.setdefault("dilation", 1)
        out.append((kind, src, dst, tuple(sorted((k, int(v)) for k, v in a.items()))))
    return out


def _parse(trace):
    seq = []
    for t in trace:
        t = str(t)
        kind, rest = t.split("(", 1)
        src, rest = rest.split(")->(", 1)
        dst, rest = rest.split(")", 1)
        args = dict(kv.split("=") for kv in rest.split())
        if kind == "StopGradient":          # pool2 under TRAIN.FREEZE_AT == 2: no forward effect
            assert (src, dst) == ("pool2", "pool2")
            continue
        seq.append((kind, src, dst, args))
    return _normalise(seq)


@pytest.mark.parametrize("tag,dil", [("d2", 2), ("d1", 1)])
def test_operator_sequence_is_the_reference_builders(gold, tag, dil):
    import torch  # noqa: F401  (the product module imports torch at module level)
    from nafwebsod_b200 import conv_body
    want = _parse(gold[tag + "_trace"])
    assert len(want) == 30                                        # 13 Conv + 13 Relu + 4 MaxPool
    assert _normalise(CB.op_sequence(dil)) == want
    assert _normalise(conv_body.body_ops(dil)) == want
    assert conv_body.spatial_scale(dil) == (int(gold[tag + "_dim_out"]), float(gold[tag + "_spatial_scale"]))


@pytest.mark.parametrize("tag,dil", [("d2", 2), ("d1", 1)])
def test_oracle_reproduces_the_builder_run(gold, tag, dil):
    params = CB.synth_params(int(gold["seed"]))
    assert np.array_equal(CB.param_checksum(params), gold["param_checksum"])
    y, dim, scale, kept = CB.conv5_body(gold["data"], params, dil, keep=("conv3_3", "pool4"))
    assert (dim, scale) == (512, float(gold[tag + "_spatial_scale"]))
    assert np.array_equal(kept["conv3_3"], gold[tag + "_conv3_3"])
    assert np.array_equal(kept["pool4"], gold[tag + "_pool4"])
    assert np.array_equal(y, gold[tag + "_conv5_3"])
    # pool4 of the dilated body keeps the 1/8 grid minus one row / column (kernel 2, stride 1, pad 0)
    h8, w8 = gold["data"].shape[2] // 8, gold["data"].shape[3] // 8
    assert y.shape[2:] == ((h8 - 1, w8 - 1) if dil == 2 else (h8 // 2, w8 // 2))


def test_bf16_storage_stays_inside_the_bf16_bar(gold):
    """The product keeps every activation and weight in bf16 (float32 accumulation).  Emulated on the CPU, thirteen layers
    of that cost 7e-3 relative L2 on conv5_3 against the float32 builder run -- inside the 1e-2 bar of the bf16 path."""
    params = CB.synth_params(int(gold["seed"]))
    y, _, _, _ = CB.conv5_body(gold["data"], params, 2, round_bf16=True)
    want = gold["d2_conv5_3"].astype(np.float64)
    rel = np.linalg.norm(y.astype(np.float64) - want) / np.linalg.norm(want)
    assert 1e-4 < rel <= 1e-2, rel


def test_weight_permutation_matches_the_patch_order():
    """conv_body.VGG16ConvBody stores W as [Cout, (kh, kw, c)]: a patch matrix in that K-order times W^T must equal the
    convolution (checked in float32 on the CPU with an explicit NumPy im2col, incl. dilation 2 and the 3 -> 8 plane pad)."""
    import torch
    import torch.nn.functional as Fn
    rng = np.random.default_rng(1)
    for cin, cout, dil in ((3, 5, 1), (8, 4, 2)):
        x = rng.standard_normal((2, cin, 6, 7)).astype(np.float32)
        w = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32)
        cp = (cin + 7) // 8 * 8
        wm = np.zeros((cout, 3, 3, cp), np.float32)
        wm[..., :cin] = w.transpose(0, 2, 3, 1)
        xcl = np.zeros((2, 6, 7, cp), np.float32)
        xcl[..., :cin] = x.transpose(0, 2, 3, 1)
        pad = np.pad(xcl, ((0, 0), (dil, dil), (dil, dil), (0, 0)))
        cols = np.stack([pad[:, kh * dil:kh * dil + 6, kw * dil:kw * dil + 7, :] for kh in range(3) for kw in range(3)], axis=3)
        y = cols.reshape(2 * 6 * 7, 9 * cp) @ wm.reshape(cout, 9 * cp).T
        want = Fn.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=dil, dilation=dil).numpy().transpose(0, 2, 3, 1)
        np.testing.assert_allclose(y.reshape(2, 6, 7, cout), want, rtol=1e-4, atol=1e-4)


def test_product_weight_layout_on_the_host(gold):
    """VGG16ConvBody.load_reference_params (pure tensor code: runs on the CPU): GEMM operand [Cout, (kh, kw, c)] in bf16 with
    conv1_1's three planes zero-padded to eight, biases float32; wrong shapes are refused like the reference's
    initialize_gpu_from_weights_file (utils/net_wsl.py:105-111)."""
    import torch
    from nafwebsod_b200.conv_body import VGG16ConvBody
    params = CB.synth_params(int(gold["seed"]))
    body = VGG16ConvBody(dilation=2, device="cpu")
    body.load_reference_params(params)
    assert sorted(body.w) == sorted(k[:-2] for k in params if k.endswith("_w")) and len(body.w) == 13
    for name, cin in (("conv1_1", 3), ("conv3_2", 256), ("conv5_3", 512)):
        w = params[name + "_w"]
        cout, cp = w.shape[0], (cin + 7) // 8 * 8
        got = body.w[name]
        assert got.dtype == torch.bfloat16 and tuple(got.shape) == (cout, 9 * cp) and got.is_contiguous()
        want = np.zeros((cout, 3, 3, cp), np.float32)
        want[..., :cin] = w.transpose(0, 2, 3, 1)
        assert torch.equal(got.float(), torch.from_numpy(want.reshape(cout, 9 * cp)).to(torch.bfloat16).float())
        assert body.b[name].dtype == torch.float32 and np.array_equal(body.b[name].numpy(), params[name + "_b"])
    bad = dict(params)
    bad["conv2_1_w"] = params["conv2_1_w"][:, :32]
    with pytest.raises(RuntimeError, match="conv2_1"):
        VGG16ConvBody(dilation=1, device="cpu").load_reference_params(bad)
    with pytest.raises(RuntimeError, match="DILATION"):
        VGG16ConvBody(dilation=3, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        body.feed_image(torch.zeros(1, 3, 8, 8))


@pytest.mark.parametrize("tag,dil", [("d2", 2), ("d1", 1)])
def test_product_body_plumbing_on_stand_in_kernels(gold, monkeypatch, tag, dil):
    """VGG16ConvBody.run() with the two device entry points replaced -- in this test only -- by torch CPU stand-ins that
    consume exactly what the kernels consume (channels-last bf16 map, the [Cout, (kh,kw,c)] bf16 operand, float32 bias):
    the layer sequence, the dilation / stride arguments, the padded first layer and the return values must reproduce the
    oracle evaluated on the same bf16-rounded inputs (what the GPU test asserts of the real kernels)."""
    import torch
    import torch.nn.functional as Fn
    from nafwebsod_b200 import ops
    from nafwebsod_b200.conv_body import VGG16ConvBody, add_VGG16_conv5_body_origin

    seen_implicit = []

    def conv(X, Wmat, b, *, dilation=1, relu=True, cols=None, implicit=None):
        seen_implicit.append((X.shape[3], implicit))
        cout, cp = Wmat.shape[0], X.shape[3]
        w = Wmat.float().reshape(cout, 3, 3, cp).permute(0, 3, 1, 2).contiguous()
        y = Fn.conv2d(X.float().permute(0, 3, 1, 2).contiguous(), w, b, padding=dilation, dilation=dilation)
        return (torch.relu(y) if relu else y).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)

    def pool(X, *, stride=2):
        return Fn.max_pool2d(X.float().permute(0, 3, 1, 2).contiguous(), 2, stride).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    monkeypatch.setattr(ops, "Conv3x3Relu", conv)
    monkeypatch.setattr(ops, "MaxPool2x2", pool)
    params = CB.synth_params(int(gold["seed"]))
    body = VGG16ConvBody(dilation=dil, device="cpu")
    body.load_reference_params(params)
    x = torch.zeros((1,) + gold["data"].shape[2:] + (8,), dtype=torch.bfloat16)        # what feed_image builds on the device
    x[..., :3] = torch.from_numpy(gold["data"]).permute(0, 2, 3, 1).to(torch.bfloat16)
    body.blobs["data"] = x
    y, dim, scale = add_VGG16_conv5_body_origin(body)
    assert (dim, scale) == (512, float(gold[tag + "_spatial_scale"])) and y.dtype == torch.bfloat16 and body.blobs["conv5_3"] is y
    want, _, _, _ = CB.conv5_body(gold["data"], params, dil, round_bf16=True)
    got = y.float().numpy().transpose(0, 3, 1, 2)
    assert got.shape == want.shape
    rel = np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64))
    assert rel <= 1e-6, rel            # same torch functions at the same rounding points: equal up to the padded first layer's sum order
    # the host-side choice of the convolution kernel: default = the library decides (None); never the implicit GEMM for the
    # padded 8-plane first layer when forced on
    assert all(imp is None for _, imp in seen_implicit)
    del seen_implicit[:]
    body.implicit = True
    body.run()
    assert [imp for _, imp in seen_implicit] == [cin % 64 == 0 for cin, _ in seen_implicit] and not seen_implicit[0][1]
    body.implicit = None
    _, _, _ = body.run(keep=("pool4",))
    assert tuple(body.blobs["pool4"].shape[1:3]) == tuple(gold[tag + "_pool4"].shape[2:])
