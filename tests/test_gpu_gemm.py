"""GPU parity tests of the tcgen05 FC GEMMs (through the C ABI) against fp32 references."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2   # north_star: rel <= 1e-2 on the bf16 path
TF32_TOL = 1e-3   # rel <= 1e-3 on the fp32/TF32 path


def _ops():
    from nafwebsod_b200 import ops
    return ops


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def _data(M, N, K, dtype, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(M, K, device="cuda", generator=g).clamp_min(0) * 0.5          # post-ReLU-like activations
    W = torch.randn(N, K, device="cuda", generator=g) * 0.02
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    mask = (torch.rand(M, N, device="cuda", generator=g) < 0.5).to(torch.uint8)
    return X.to(dtype), W.to(dtype), b, mask


SHAPES = [(128, 256, 64), (128, 256, 512), (300, 520, 200), (4000, 4096, 1024), (777, 40, 4096), (2000, 8192, 1568)]


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, BF16_TOL), (torch.float32, TF32_TOL)])
@pytest.mark.parametrize("shape", SHAPES)
def test_fc_forward(shape, dtype, tol):
    ops = _ops()
    M, N, K = shape
    if dtype == torch.float32 and K % 4:
        pytest.skip("ld must be a multiple of 16 bytes")
    X, W, b, mask = _data(M, N, K, dtype)
    ref = X.float() @ W.float().t() + b
    Y = ops.FC(X, W, b, out_dtype=torch.float32)
    assert _rel(Y, ref) <= tol
    Y = ops.FC(X, W, b, relu=True, dropout_mask=mask)
    ref2 = torch.relu(ref) * mask.float() * 2
    assert Y.dtype == dtype and _rel(Y.float(), ref2) <= tol
    Y = ops.FC(X, W, None, relu=True, out_dtype=torch.float32)
    assert _rel(Y, torch.relu(X.float() @ W.float().t())) <= tol


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, BF16_TOL), (torch.float32, TF32_TOL)])
@pytest.mark.parametrize("shape", [(128, 256, 256), (300, 520, 200), (4000, 4096, 4096), (777, 40, 4096), (1000, 8192, 512)])
def test_fc_backward(shape, dtype, tol):
    ops = _ops()
    M, N, K = shape
    X, W, b, mask = _data(M, N, K, dtype, seed=1)
    g = torch.Generator(device="cuda").manual_seed(2)
    dY = (torch.randn(M, N, device="cuda", generator=g) * 0.1).to(dtype)
    # dX with the ReLU / dropout gradient of the layer below (X plays the role of its own post-dropout activation)
    dX = ops.FCGradientX(dY, W, act_below=X, dropout=True, out_dtype=torch.float32)
    ref = (dY.float() @ W.float()) * 2 * (X.float() > 0)
    assert _rel(dX, ref) <= tol
    dX = ops.FCGradientX(dY, W, out_dtype=torch.float32)
    assert _rel(dX, dY.float() @ W.float()) <= tol
    dW, db = ops.FCGradientW(dY, X)
    assert _rel(dW, dY.float().t() @ X.float()) <= tol
    assert _rel(db, dY.float().sum(0)) <= 1e-4
    dW2, db2 = ops.FCGradientW(dY, X, dW=dW.clone(), db=db.clone(), accumulate=True)
    assert _rel(dW2, 2 * (dY.float().t() @ X.float())) <= tol and _rel(db2, 2 * dY.float().sum(0)) <= 1e-4


def test_fc_column_slices_and_linearity():
    """Operands addressed as column slices of wider buffers (the fused two-stack layout), and a
    size-independent property at full size: FC is linear in X."""
    ops = _ops()
    M, K, N = 4000, 4096, 4096
    g = torch.Generator(device="cuda").manual_seed(5)
    wide = (torch.randn(M, 2 * K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    out = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.FC(wide[:, K:], W, None, out=out[:, N:])
    ref = wide[:, K:].float() @ W.float().t()
    assert _rel(out[:, N:].float(), ref) <= BF16_TOL and (out[:, :N] == 0).all()
    a = ops.FC(wide[:, :K], W, None, out_dtype=torch.float32)
    b = ops.FC(wide[:, K:], W, None, out_dtype=torch.float32)
    s = ops.FC((wide[:, :K].float() + wide[:, K:].float()).to(torch.bfloat16), W, None, out_dtype=torch.float32)
    assert _rel(s, a + b) <= 2 * BF16_TOL


def test_fc_full_size_fc6():
    """fc6 at BASELINE config 2 size (4000 RoIs, both stacks fused: N = 8192, K = 25088), bf16."""
    ops = _ops()
    M, N, K = 4000, 8192, 25088
    g = torch.Generator(device="cuda").manual_seed(7)
    X = (torch.rand(M, K, device="cuda", generator=g) * (torch.rand(M, K, device="cuda", generator=g) < 0.5)).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.01).to(torch.bfloat16)
    b = torch.zeros(N, device="cuda")
    Y = ops.FC(X, W, b, relu=True, out_dtype=torch.float32)
    rows = torch.randint(0, M, (64,), device="cuda")
    ref = torch.relu(X[rows].float() @ W.float().t())
    assert _rel(Y[rows], ref) <= BF16_TOL
    dY = (torch.randn(M, N, device="cuda", generator=g) * 0.01).to(torch.bfloat16)
    dW, _ = ops.FCGradientW(dY, X, want_db=False)
    cols = torch.randint(0, K, (256,), device="cuda")
    refW = dY.float().t() @ X[:, cols].float()
    assert _rel(dW[:, cols], refW) <= BF16_TOL


def test_to_bf16_and_errors():
    ops = _ops()
    x = torch.randn(100, 40, device="cuda")
    assert torch.equal(ops.to_bf16(x), x.to(torch.bfloat16))
    wide = torch.randn(100, 80, device="cuda")
    assert torch.equal(ops.to_bf16(wide[:, 40:]), wide[:, 40:].to(torch.bfloat16))
    with pytest.raises(RuntimeError):
        ops.FC(torch.zeros(8, 16, device="cuda", dtype=torch.bfloat16), torch.zeros(8, 24, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):     # K*2 bytes not a multiple of 16
        ops.FC(torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16), torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,iter_count,keep", [((256, 512, 256), 0, True), ((300, 520, 200), 3, True), ((1000, 768, 1568), 2, False),
                                                   ((777, 40, 4096), 1, True), ((2000, 1024, 3136), 5, False)])
def test_fc_backward_w_fused_sgd_is_bit_exact(shape, iter_count, keep, dtype):
    """FCGradient's dW consumed by ACMWeightDecayMomentumSGDUpdate inside the GEMM epilogue == the weight-gradient
    GEMM followed by the stand-alone update kernel: the same fp32 operations in the same order on the same
    gradient, so momentum, master parameter and GEMM-operand shadow agree bit for bit (also on the scalar edge
    tiles of shapes that are not multiples of the tile, and on the first call, which ignores the momentum buffer)."""
    ops = _ops()
    M, N, K = shape
    X, W, b, mask = _data(M, N, K, dtype, seed=3)
    g = torch.Generator(device="cuda").manual_seed(4)
    dY = (torch.randn(M, N, device="cuda", generator=g) * 0.1).to(dtype)
    p0 = W.float().contiguous()
    m0 = torch.randn(N, K, device="cuda", generator=g) * 1e-3
    lr = torch.tensor([3e-3], device="cuda")
    hyper = dict(momentum=0.9, weight_decay=5e-4, lr_mult=1.0, gpu_num=1, iter_count=iter_count)
    # reference schedule: GEMM -> gradient buffer -> update kernel
    dW, db = ops.FCGradientW(dY, X)
    p_ref, m_ref, s_ref = p0.clone(), m0.clone(), torch.empty(N, K, dtype=dtype, device="cuda")
    ops.ACMWeightDecayMomentumSGDUpdate(dW.view(-1), m_ref.view(-1), lr, p_ref.view(-1), None, p_shadow=s_ref.view(-1), **hyper)
    # fused
    p, m, s = p0.clone(), m0.clone(), torch.empty(N, K, dtype=dtype, device="cuda")
    dW2 = torch.full((N, K), float("nan"), device="cuda") if keep else None
    db2 = torch.empty(N, device="cuda")
    ops.FCGradientW_SGDUpdate(dY, X, m, lr, p, dW=dW2, db=db2, p_shadow=s, **hyper)
    torch.cuda.synchronize()
    assert torch.equal(p, p_ref) and torch.equal(m, m_ref)
    assert torch.equal(s.view(torch.int16 if dtype == torch.bfloat16 else torch.int32),
                       s_ref.view(torch.int16 if dtype == torch.bfloat16 else torch.int32))
    assert not torch.equal(p, p0)
    if keep:
        assert torch.equal(dW2, dW)
    assert _rel(db2, db) <= 1e-5
    # row panels of a wider parameter block (how the head calls it): leading dimension > K is not needed, row slices are
    if N >= 512:
        p, m, s = p0.clone(), m0.clone(), torch.empty(N, K, dtype=dtype, device="cuda")
        h = (N // 2) // 64 * 64                    # panel boundary: a TMA operand base must stay 16-byte aligned
        for r0, r1 in ((0, h), (h, N)):
            ops.FCGradientW_SGDUpdate(dY[:, r0:r1], X, m[r0:r1], lr, p[r0:r1], p_shadow=s[r0:r1], **hyper)
        assert torch.equal(p, p_ref) and torch.equal(m, m_ref)


def test_fc_backward_w_fused_sgd_errors():
    ops = _ops()
    X, W, b, mask = _data(128, 64, 128, torch.bfloat16)
    dY = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    lr = torch.tensor([1e-3], device="cuda")
    p, m = torch.zeros(64, 128, device="cuda"), torch.zeros(64, 128, device="cuda")
    with pytest.raises(RuntimeError):
        ops.FCGradientW_SGDUpdate(dY, X, m[:32], lr, p)                       # shape mismatch
    with pytest.raises(RuntimeError):
        ops.FCGradientW_SGDUpdate(dY, X, m, lr, p, accumulate=True)           # accumulate without a gradient buffer
    with pytest.raises(RuntimeError):
        ops.FCGradientW_SGDUpdate(dY, X, m, torch.zeros(2, device="cuda"), p)
