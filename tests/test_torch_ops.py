"""CPU tests of na-fwebsod_b200/torch_ops.py: the reference's operators registered as ``torch.ops.nawsod.*``.

What can be checked without a GPU: the schemas (the Caffe2 blob order), the shape functions on meta tensors, that a CPU
tensor is refused (no fallback), and the AUTOGRAD WIRING -- for that the ``ops`` entry points the custom ops call are
replaced, in this test only, by stand-ins that evaluate the CPU oracle (tests may use the oracle as the checker): a
backward pass through ``torch.ops.nawsod.*`` must hand every saved blob to the right gradient operator and return its
outputs in the right slots.  The kernels themselves are covered by the GPU suite (tests/test_gpu_zzz_torch_ops.py)."""
import numpy as np
import pytest
import torch

from oracle import nawsod_oracle as O
from oracle import test_time_oracle as TT


@pytest.fixture(scope="module")
def tops():
    from nafwebsod_b200 import torch_ops
    return torch_ops


def test_schemas_follow_the_caffe2_blob_order(tops):
    want = {
        "RoIPoolF": "(Tensor X, Tensor rois, SymInt pooled_h, SymInt pooled_w, float spatial_scale) -> (Tensor, Tensor)",
        "RoIPoolFGradient": "(Tensor X, Tensor rois, Tensor argmax, Tensor dY) -> Tensor",
        "RoIFeatureBoost": "(Tensor X, Tensor S) -> Tensor",
        "FC": "(Tensor X, Tensor W, Tensor b) -> Tensor",
        "FCGradient": "(Tensor X, Tensor W, Tensor dY) -> (Tensor, Tensor, Tensor)",
        "RoIIoU": "(Tensor rois) -> Tensor",
        "CrossEntropyWithLogits": "(Tensor X, Tensor L, bool is_mean) -> Tensor",
        "WeightedCrossEntropyWithLogits": "(Tensor X, Tensor L, Tensor W, bool is_mean) -> Tensor",
        "WeightedCrossEntropyWithLogitsGradient": "(Tensor X, Tensor L, Tensor W, Tensor dY, bool is_mean) -> Tensor",
        "MinEntropyLoss": "(Tensor X, Tensor L) -> Tensor",
        "ACMWeightDecayMomentumSGDUpdate": "(Tensor g, Tensor(a1!) m, Tensor lr, Tensor(a3!) p, Tensor(a4!)? acc, float momentum, "
                                           "SymInt iter_size, SymInt gpu_num, float lr_mult, float weight_decay, SymInt iter_count) -> ()",
    }
    for name, sig in want.items():
        schema = str(getattr(torch.ops.nawsod, name).default._schema)
        assert schema == "nawsod::" + name + sig, schema


def test_shape_functions_on_meta_tensors(tops):
    m = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device="meta")
    Y, A = torch.ops.nawsod.RoIPoolF(m(2, 512, 38, 50), m(300, 5), 7, 7, 1.0 / 16)
    assert tuple(Y.shape) == tuple(A.shape) == (300, 512, 7, 7) and Y.dtype == torch.float32 and A.dtype == torch.int32
    assert tuple(torch.ops.nawsod.RoIPoolFGradient(m(2, 512, 38, 50), m(300, 5), A, Y).shape) == (2, 512, 38, 50)
    assert tuple(torch.ops.nawsod.RoIFeatureBoost(Y, m(300, 1)).shape) == (300, 512, 7, 7)
    y = torch.ops.nawsod.FC(m(300, 25088, dt=torch.bfloat16), m(4096, 25088, dt=torch.bfloat16), m(4096))
    assert tuple(y.shape) == (300, 4096) and y.dtype == torch.bfloat16
    dW, db, dX = torch.ops.nawsod.FCGradient(m(300, 25088, dt=torch.bfloat16), m(4096, 25088, dt=torch.bfloat16), y)
    assert (tuple(dW.shape), dW.dtype, tuple(db.shape), tuple(dX.shape)) == ((4096, 25088), torch.float32, (4096,), (300, 25088))
    assert tuple(torch.ops.nawsod.RoIIoU(m(300, 5)).shape) == (300, 300)
    assert torch.ops.nawsod.WeightedCrossEntropyWithLogits(m(1, 20), m(1, 20), m(1, 20), True).shape == ()
    assert torch.ops.nawsod.MinEntropyLoss(m(300, 20), m(1, 20)).shape == ()


def test_cpu_tensors_are_refused(tops):
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.nawsod.RoIIoU(torch.zeros(4, 5))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.nawsod.RoIPoolF(torch.zeros(1, 8, 6, 6), torch.zeros(2, 5), 7, 7, 0.0625)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.nawsod.FC(torch.zeros(4, 8), torch.zeros(3, 8), torch.zeros(3))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture
def oracle_backed_ops(monkeypatch, tops):
    """TEST-ONLY stand-ins for the ctypes entry points (same signatures, CPU oracle arithmetic)."""
    from nafwebsod_b200 import ops
    n = lambda t: t.detach().numpy()

    def roi_pool_f(X, rois, *, pooled_h=7, pooled_w=7, spatial_scale=1 / 16, **kw):
        Y, A = O.roi_pool_f(n(X), n(rois), spatial_scale, pooled_h, pooled_w)
        return _t(Y), _t(A)
    monkeypatch.setattr(ops, "RoIPoolF", roi_pool_f)
    monkeypatch.setattr(ops, "RoIPoolFGradient", lambda X, rois, A, dY, layout="NCHW": _t(O.roi_pool_f_grad(tuple(X.shape), n(rois), n(A), n(dY))))
    monkeypatch.setattr(ops, "RoIFeatureBoost", lambda X, S, out=None: _t(O.roi_feature_boost(n(X), n(S))))
    monkeypatch.setattr(ops, "FC", lambda X, W, b=None, **kw: _t(O.fc(n(X), n(W), n(b))))
    monkeypatch.setattr(ops, "FCGradientW", lambda dY, X, **kw: tuple(_t(a) for a in O.fc_grad(n(X), np.zeros((dY.shape[1], X.shape[1]), np.float32), n(dY))[:2]))
    monkeypatch.setattr(ops, "FCGradientX", lambda dY, W, **kw: _t(O.fc_grad(np.zeros((dY.shape[0], W.shape[1]), np.float32), n(W), n(dY))[2]))
    monkeypatch.setattr(ops, "CrossEntropyWithLogits", lambda X, L, is_mean=False: torch.tensor(O.cross_entropy_with_logits(n(X), n(L), None, is_mean)))
    monkeypatch.setattr(ops, "CrossEntropyWithLogitsGradient", lambda X, L, dY, is_mean=False: _t(O.cross_entropy_with_logits_grad(n(X), n(L), n(dY).reshape(-1)[0], None, is_mean)))
    monkeypatch.setattr(ops, "WeightedCrossEntropyWithLogits", lambda X, L, W, is_mean=False: torch.tensor(O.cross_entropy_with_logits(n(X), n(L), n(W), is_mean)))
    monkeypatch.setattr(ops, "WeightedCrossEntropyWithLogitsGradient", lambda X, L, W, dY, is_mean=False: _t(O.cross_entropy_with_logits_grad(n(X), n(L), n(dY).reshape(-1)[0], n(W), is_mean)))
    monkeypatch.setattr(ops, "MinEntropyLoss", lambda X, L: torch.tensor(np.float32(TT.min_entropy_loss(n(X), n(L))[0])))
    monkeypatch.setattr(ops, "MinEntropyLossGradient", lambda X, L, dY: _t(TT.min_entropy_loss_grad(n(X), n(L), np.float32(n(dY).reshape(-1)[0]))))

    def sgd(g, m, lr, p, acc, *, momentum, iter_size, gpu_num, lr_mult, weight_decay, iter_count):
        m2, p2, _, _ = O.acm_sgd_update(n(g), n(m).copy(), n(lr).reshape(-1)[0], n(p).copy(), np.zeros_like(n(g)) if acc is None else n(acc).copy(), momentum=momentum,
                                        weight_decay=weight_decay, lr_mult=lr_mult, iter_size=iter_size, gpu_num=gpu_num, iter_count=iter_count)
        m.copy_(_t(m2)); p.copy_(_t(p2))
    monkeypatch.setattr(ops, "ACMWeightDecayMomentumSGDUpdate", sgd)
    return ops


def test_autograd_routes_through_the_reference_gradient_operators(oracle_backed_ops):
    """conv5 -> RoIPoolF -> RoIFeatureBoost -> FC -> (column sums as a stand-in head) -> WeightedCrossEntropyWithLogits,
    backward through torch.ops.nawsod.*, against the same chain evaluated operator by operator with the oracle."""
    rng = np.random.default_rng(0)
    X = O.synth_conv5(1, 8, 10, 12, seed=3)
    rois = O.synth_rois(12, 160, 192, seed=4)
    S = (rng.random((12, 1)) + 1).astype(np.float32)
    W = (rng.standard_normal((5, 8 * 49)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(5) * 0.1).astype(np.float32)
    L = np.zeros((1, 5), np.float32); L[0, 2] = 1
    Wc = rng.random((1, 5)).astype(np.float32)
    tX, tW, tb = _t(X).requires_grad_(), _t(W).requires_grad_(), _t(b).requires_grad_()
    Y, A = torch.ops.nawsod.RoIPoolF(tX, _t(rois), 7, 7, 1.0 / 16)
    assert not A.requires_grad and A.dtype == torch.int32
    Yb = torch.ops.nawsod.RoIFeatureBoost(Y, _t(S))
    fc = torch.ops.nawsod.FC(Yb.reshape(12, -1), tW, tb)
    prob = torch.sigmoid(fc).mean(dim=0, keepdim=True)                       # any differentiable [1, C] in (0, 1)
    loss = torch.ops.nawsod.WeightedCrossEntropyWithLogits(prob, _t(L), _t(Wc), True)
    loss.backward()
    # the same chain with the oracle's operators
    Yo, Ao = O.roi_pool_f(X, rois, 1.0 / 16)
    Ybo = O.roi_feature_boost(Yo, S)
    fco = O.fc(Ybo.reshape(12, -1), W, b)
    sg = 1.0 / (1.0 + np.exp(-fco.astype(np.float64)))
    po = sg.mean(axis=0, keepdims=True).astype(np.float32)
    assert np.allclose(loss.item(), O.cross_entropy_with_logits(po, L, Wc, True), rtol=1e-6)
    dprob = O.cross_entropy_with_logits_grad(po, L, np.float32(1.0), Wc, True)
    dfc = (dprob.astype(np.float64) / 12 * sg * (1 - sg)).astype(np.float32)
    dW, db, dYb = O.fc_grad(Ybo.reshape(12, -1), W, dfc)
    dY = O.roi_feature_boost_grad(dYb.reshape(Yo.shape), S)
    dX = O.roi_pool_f_grad(X.shape, rois, Ao, dY)
    np.testing.assert_allclose(tW.grad.numpy(), dW, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(tb.grad.numpy(), db, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(tX.grad.numpy(), dX, rtol=2e-4, atol=1e-7)
    assert np.array_equal(tX.grad.numpy() == 0, dX == 0)                      # gradient lands exactly on the argmax cells


def test_losses_and_update_route_their_arguments(oracle_backed_ops):
    rng = np.random.default_rng(1)
    P = rng.random((1, 6)).astype(np.float32) * 0.9 + 0.05
    L = (rng.random((1, 6)) < 0.4).astype(np.float32)
    for is_mean in (False, True):
        tP = _t(P).requires_grad_()
        (torch.ops.nawsod.CrossEntropyWithLogits(tP, _t(L), is_mean) * 3.0).backward()
        np.testing.assert_allclose(tP.grad.numpy(), O.cross_entropy_with_logits_grad(P, L, np.float32(3.0), None, is_mean), rtol=1e-6)
    Xp = rng.random((9, 6)).astype(np.float32)
    tXp = _t(Xp).requires_grad_()
    (torch.ops.nawsod.MinEntropyLoss(tXp, _t(L)) * 0.5).backward()
    np.testing.assert_allclose(tXp.grad.numpy(), TT.min_entropy_loss_grad(Xp, L, np.float32(0.5)), rtol=1e-6)
    g, m, p = (rng.standard_normal(40).astype(np.float32) for _ in range(3))
    tm, tp = _t(m.copy()), _t(p.copy())
    out = torch.ops.nawsod.ACMWeightDecayMomentumSGDUpdate(_t(g), tm, torch.tensor([1e-2]), tp, None, 0.9, 1, 2, 2.0, 5e-4, 3)
    assert out is None
    m2, p2, _, _ = O.acm_sgd_update(g, m.copy(), np.float32(1e-2), p.copy(), np.zeros_like(g), momentum=0.9, weight_decay=5e-4, lr_mult=2.0,
                                    iter_size=1, gpu_num=2, iter_count=3)
    assert np.array_equal(tm.numpy(), m2) and np.array_equal(tp.numpy(), p2) and not np.array_equal(p2, p) and np.isfinite(p2).all()


def test_opcheck_on_the_oracle_stand_ins(oracle_backed_ops):
    """torch.library.opcheck (schema incl. the in-place annotations of the SGD op, autograd registration, fake-tensor shape
    functions against real outputs) -- runnable on the CPU because the stand-ins accept CPU tensors."""
    from torch.library import opcheck
    rng = np.random.default_rng(2)
    utils = ("test_schema", "test_autograd_registration", "test_faketensor")
    X = _t(O.synth_conv5(1, 8, 10, 12, seed=3)).requires_grad_()
    rois = _t(O.synth_rois(6, 160, 192, seed=4))
    opcheck(torch.ops.nawsod.RoIPoolF.default, (X, rois, 7, 7, 1.0 / 16), test_utils=utils)
    Y = torch.rand(6, 8, 7, 7, requires_grad=True)
    opcheck(torch.ops.nawsod.RoIFeatureBoost.default, (Y, torch.rand(6, 1) + 1), test_utils=utils)
    opcheck(torch.ops.nawsod.FC.default, (torch.rand(6, 392, requires_grad=True), torch.rand(5, 392, requires_grad=True),
                                          torch.rand(5, requires_grad=True)), test_utils=utils)
    P = (torch.rand(1, 5) * 0.9 + 0.05).requires_grad_()
    Lh = (torch.rand(1, 5) < 0.4).float()
    opcheck(torch.ops.nawsod.WeightedCrossEntropyWithLogits.default, (P, Lh, torch.rand(1, 5), True), test_utils=utils)
    opcheck(torch.ops.nawsod.CrossEntropyWithLogits.default, (P, Lh, False), test_utils=utils)
    opcheck(torch.ops.nawsod.MinEntropyLoss.default, (torch.rand(9, 5, requires_grad=True), Lh), test_utils=utils)
    g, m, p = (_t(rng.standard_normal(16).astype(np.float32)) for _ in range(3))
    opcheck(torch.ops.nawsod.ACMWeightDecayMomentumSGDUpdate.default, (g, m, torch.tensor([1e-2]), p, None, 0.9, 1, 1, 1.0, 5e-4, 1),
            test_utils=("test_schema", "test_faketensor"))
