"""GPU test of na-fwebsod_b200/torch_ops.py (``torch.ops.nawsod.*``: the reference's operators as PyTorch custom ops with
autograd, each one a single call through the C ABI).  Added after the round's last GPU call -- verified kernels behind new
host code -- so the file sorts last like tests/test_gpu_zzz_reference_vectors.py.  CPU counterpart (schemas, shape
functions, autograd wiring on oracle stand-ins): tests/test_torch_ops.py.

Chain: conv5 -> RoIPoolF -> RoIFeatureBoost -> (bf16) FC -> sigmoid / RoI mean -> WeightedCrossEntropyWithLogits, forward
and backward through autograd, against the same chain evaluated operator by operator with the CPU oracle.  Bars: RoIPoolF
/ boost bit-exact, loss rel <= 1e-2 and gradients rel-L2 <= 2e-2 (bf16 FC operands; the oracle gets the same bf16-rounded
operands), dX non-zero exactly on the argmax cells."""
import numpy as np
import pytest
import torch

from oracle import nawsod_oracle as O
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def test_reference_operator_chain_forward_and_backward():
    from nafwebsod_b200 import torch_ops  # noqa: F401  (registers torch.ops.nawsod.*)
    rng = np.random.default_rng(0)
    R, Cc, C = 64, 16, 8
    X = O.synth_conv5(1, Cc, 12, 16, seed=3)
    rois = O.synth_rois(R, 192, 256, seed=4)
    S = (rng.random((R, 1)) + 1).astype(np.float32)
    W = _bf(rng.standard_normal((C, Cc * 49)) * 0.05)
    b = (rng.standard_normal(C) * 0.1).astype(np.float32)
    L = np.zeros((1, C), np.float32); L[0, 2] = 1
    Wc = rng.random((1, C)).astype(np.float32)

    tX = dev(X).requires_grad_()
    tW = dev(W).to(torch.bfloat16).requires_grad_()
    tb = dev(b).requires_grad_()
    Y, A = torch.ops.nawsod.RoIPoolF(tX, dev(rois), 7, 7, 1.0 / 16)
    Yb = torch.ops.nawsod.RoIFeatureBoost(Y, dev(S))
    fc = torch.ops.nawsod.FC(Yb.reshape(R, -1).to(torch.bfloat16), tW, tb)
    assert fc.dtype == torch.bfloat16 and not A.requires_grad
    prob = torch.sigmoid(fc.float()).mean(dim=0, keepdim=True)
    loss = torch.ops.nawsod.WeightedCrossEntropyWithLogits(prob, dev(L), dev(Wc), True)
    loss.backward()
    torch.cuda.synchronize()

    Yo, Ao = CO.roi_pool_f(X, rois, 1.0 / 16)
    assert np.array_equal(Y.detach().cpu().numpy(), Yo) and np.array_equal(A.cpu().numpy(), Ao)
    Ybo = O.roi_feature_boost(Yo, S)
    assert np.array_equal(Yb.detach().cpu().numpy(), Ybo)
    feat = _bf(Ybo.reshape(R, -1))
    fco = _bf(O.fc(feat, W, b))                                     # the product stores the FC output in bf16
    assert rel_l2(fc.detach().float().cpu().numpy(), fco) <= 1e-2
    sg = 1.0 / (1.0 + np.exp(-fco.astype(np.float64)))
    po = sg.mean(axis=0, keepdims=True).astype(np.float32)
    want_loss = float(O.cross_entropy_with_logits(po, L, Wc, True))
    assert abs(loss.item() - want_loss) <= 1e-2 * abs(want_loss)
    dprob = O.cross_entropy_with_logits_grad(po, L, np.float32(1.0), Wc, True)
    dfc = _bf((dprob.astype(np.float64) / R * sg * (1 - sg)).astype(np.float32))      # autograd hands FCGradient a bf16 dY
    dW, db, dfeat = O.fc_grad(feat, W, dfc)
    dY = O.roi_feature_boost_grad(_bf(dfeat).reshape(Yo.shape), S)                    # dX of FCGradient comes back in bf16
    dX = O.roi_pool_f_grad(X.shape, rois, Ao, dY)
    assert rel_l2(tW.grad.float().cpu().numpy(), dW) <= 2e-2
    assert rel_l2(tb.grad.cpu().numpy(), db) <= 2e-2
    got_dX = tX.grad.cpu().numpy()
    assert rel_l2(got_dX, dX) <= 2e-2
    assert np.array_equal(got_dX == 0, dX == 0)


def test_standalone_ops_match_the_verified_wrappers():
    """Each custom op returns exactly what the ops.* wrapper the rest of the GPU suite verifies returns."""
    from nafwebsod_b200 import ops, torch_ops  # noqa: F401
    rng = np.random.default_rng(1)
    rois = dev(O.synth_rois(40, 320, 400, seed=2))
    assert torch.equal(torch.ops.nawsod.RoIIoU(rois), ops.RoIIoU(rois))
    P = dev(rng.random((1, 20)).astype(np.float32) * 0.9 + 0.05).requires_grad_()
    Lh = dev((rng.random((1, 20)) < 0.3).astype(np.float32))
    y = torch.ops.nawsod.CrossEntropyWithLogits(P, Lh, True)
    assert torch.equal(y.detach(), ops.CrossEntropyWithLogits(P.detach(), Lh, is_mean=True))
    (y * 2.0).backward()
    assert torch.equal(P.grad, ops.CrossEntropyWithLogitsGradient(P.detach(), Lh, torch.tensor([2.0], device="cuda"), is_mean=True))
    Xp = dev(rng.random((50, 20)).astype(np.float32)).requires_grad_()
    me = torch.ops.nawsod.MinEntropyLoss(Xp, Lh)
    me.backward()
    np.testing.assert_allclose(Xp.grad.cpu().numpy(),
                               ops.MinEntropyLossGradient(Xp.detach(), Lh, torch.tensor([1.0], device="cuda")).cpu().numpy(), rtol=1e-6)
    n = 4096
    g, m0, p0 = (dev(rng.standard_normal(n).astype(np.float32)) for _ in range(3))
    lr = torch.tensor([1e-2], device="cuda")
    ma, pa, mb, pb = m0.clone(), p0.clone(), m0.clone(), p0.clone()
    torch.ops.nawsod.ACMWeightDecayMomentumSGDUpdate(g, ma, lr, pa, None, 0.9, 1, 2, 2.0, 5e-4, 3)
    ops.ACMWeightDecayMomentumSGDUpdate(g, mb, lr, pb, None, momentum=0.9, iter_size=1, gpu_num=2, lr_mult=2.0, weight_decay=5e-4, iter_count=3)
    assert torch.equal(ma, mb) and torch.equal(pa, pb) and not torch.equal(pa, p0)
