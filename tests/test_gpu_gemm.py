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


def _stacked(buf, S):
    """[R, S*H] buffer -> [S, R, H] strided view of its column blocks (heads.WeblyHeadModel._stacked)."""
    R = buf.shape[0]
    return buf.view(R, S, buf.shape[1] // S).permute(1, 0, 2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(96, 128, 128), (96, 128, 12), (300, 256, 40), (2000, 1024, 512)])
def test_fc_stacked_launch_equals_per_stack_launches(shape, dtype):
    """The head's clean / noisy stacks run as ONE launch over 3-d tensor maps (nawsod_fc_*_stacks): every stacked
    op must be bit-identical to the same op launched once per stack on the 2-d column blocks -- forward (bias,
    ReLU, injected dropout mask), dX (ReLU/dropout gradient of the layer below) and dW / db, operands given as
    column-block views of [R, S*H] activation buffers exactly like heads.py passes them."""
    ops = _ops()
    R, H, N = shape
    S = 2
    g = torch.Generator(device="cuda").manual_seed(5)
    Np = (N + 15) // 16 * 16          # padded row pitch of the narrow (fc8-like) operands
    X = torch.randn(R, S * H, device="cuda", generator=g).to(dtype)
    W = (torch.randn(S, N, H, device="cuda", generator=g) * 0.1).to(dtype)
    b = torch.randn(S, Np, device="cuda", generator=g)[:, :N]
    mask = (torch.rand(S, R, Np, device="cuda", generator=g) < 0.5).to(torch.uint8)[:, :, :N]
    Y = torch.full((R, S * Np), float("nan"), device="cuda", dtype=dtype)
    Ys = _stacked(Y, S)[:, :, :N]
    ops.FC(_stacked(X, S), W, b, relu=True, dropout_mask=mask, out=Ys)
    for s in range(S):
        ref = ops.FC(X[:, s * H:(s + 1) * H], W[s], b[s].contiguous(), relu=True, dropout_mask=mask[s])
        assert torch.equal(Ys[s], ref), ("fwd", s)
    dY = torch.randn(S, R, Np, device="cuda", generator=g).to(dtype)[:, :, :N]
    dX = torch.full((R, S * H), float("nan"), device="cuda", dtype=dtype)
    ops.FCGradientX(dY, W, act_below=_stacked(X, S), out=_stacked(dX, S))
    for s in range(S):
        ref = ops.FCGradientX(dY[s], W[s], act_below=X[:, s * H:(s + 1) * H])
        assert torch.equal(dX[:, s * H:(s + 1) * H], ref), ("bwd_x", s)
    dW = torch.full((S, N, H), float("nan"), device="cuda")
    db = torch.zeros(S, Np, device="cuda")[:, :N]
    ops.FCGradientW(dY, _stacked(X, S), dW=dW, db=db)
    for s in range(S):
        rW, rb = ops.FCGradientW(dY[s], X[:, s * H:(s + 1) * H])
        assert torch.equal(dW[s], rW), ("bwd_w", s)
        assert torch.allclose(db[s], rb, rtol=1e-4, atol=1e-4), ("db", s)      # column sums use float atomics


def test_split_tf32_is_an_exact_two_term_expansion():
    """nawsod_split_tf32: hi = nearest TF32 (bit-equal to nawsod_round_to_tf32, in place allowed), lo = nearest TF32 of the
    exact remainder: both carry 13 zero low bits, and hi + lo recovers src to 2^-21 relative."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randn(257, 1003, device="cuda", generator=g) * torch.exp(4 * torch.randn(257, 1003, device="cuda", generator=g))
    src[0, :4] = torch.tensor([0.0, -0.0, 1.0, -3.5], device="cuda")
    want_hi = ops.round_to_tf32(src)
    buf = src.clone()
    hi, lo = ops.split_tf32(buf, hi=buf)                       # in place
    assert hi.data_ptr() == buf.data_ptr()
    assert torch.equal(hi.view(torch.int32), want_hi.view(torch.int32) & ~0x1FFF)     # the TF32 value round_to_tf32 stores
    for part in (hi, lo):
        assert int((part.view(torch.int32) & 0x1FFF).abs().max()) == 0
    err = (src.double() - hi.double() - lo.double()).abs()
    assert bool((err <= src.double().abs() * 2.0 ** -21).all())
    view = torch.zeros(16, 64, device="cuda")                   # column slices (leading dimensions) on every operand
    hi2, lo2 = ops.split_tf32(src[:16, 8:40], hi=view[:, :32], lo=view[:, 32:])
    assert torch.equal(view[:, :32], hi[:16, 8:40]) and torch.equal(view[:, 32:], lo[:16, 8:40])
    with pytest.raises(RuntimeError):
        ops.split_tf32(src, hi=src, lo=src)


@pytest.mark.parametrize("shape", [(300, 520, 200), (777, 40, 4096), (2000, 1024, 1568)])
def test_fc_split_operand_three_pass_reaches_fp32(shape):
    """The fp32 path: operands as (high, low) TF32 pairs, three accumulating tensor-core passes per product (FC,
    FCGradientX, FCGradientW incl. bias / ReLU / dropout epilogues, which run once on the summed product).  Against float64:
    <= 1e-5 of the output scale, against ~3e-4 for the single TF32 pass on the same data (both printed).  What is left is not
    the operand split (2^-21) but the tensor core's accumulator: a chain of n kind::tf32 MMAs loses ~n * 2^-25 of the running sum
    (measured unchunked: 4e-6 at K = 1568, 9e-6 at K = 4096), which is why ops.X3_CHUNK cuts the dominant pass into chains of
    128 MMAs that the epilogue adds up in round-to-nearest fp32."""
    ops = _ops()
    M, N, K = shape
    X, W, b, mask = _data(M, N, K, torch.float32, seed=3)
    dY = torch.randn(M, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)) * 0.1
    Xh, Xl = ops.split_tf32(X)
    Wh, Wl = ops.split_tf32(W)
    dYh, dYl = ops.split_tf32(dY)
    X64, W64, dY64 = X.double(), W.double(), dY.double()
    # forward with the whole epilogue
    ref = torch.relu(X64 @ W64.T + b.double()) * 2.0 * mask.double()
    y3 = ops.FC(Xh, Wh, b, relu=True, dropout_mask=mask, X_lo=Xl, W_lo=Wl)
    y1 = ops.FC(Xh, Wh, b, relu=True, dropout_mask=mask)
    e3, e1 = _rel(y3, ref), _rel(y1, ref)
    # dX gated by an activation pattern
    act = (torch.rand(M, K, device="cuda") < 0.5).float()
    refx = (dY64 @ W64) * 2.0 * act.double()
    dx3 = ops.FCGradientX(dYh, Wh, act_below=act, dropout=True, dY_lo=dYl, W_lo=Wl)
    dx1 = ops.FCGradientX(dYh, Wh, act_below=act, dropout=True)
    # dW, db (accumulating into a prior value)
    prior = torch.randn(N, K, device="cuda") * 0.01
    dW3, db3 = ops.FCGradientW(dYh, Xh, dW=prior.clone(), accumulate=True, db=torch.ones(N, device="cuda"), dY_lo=dYl, X_lo=Xl)
    dW1, _ = ops.FCGradientW(dYh, Xh)
    refw, refb = dY64.T @ X64 + prior.double(), dY64.sum(0) + 1.0
    print("three-pass vs one-pass TF32 (max error / output scale): fwd %.1e / %.1e, dX %.1e / %.1e, dW %.1e / %.1e" % (
        e3, e1, _rel(dx3, refx), _rel(dx1, refx), _rel(dW3, refw), _rel(dW1 + prior, refw)))
    assert e3 <= 1e-5 and _rel(dx3, refx) <= 1e-5 and _rel(dW3, refw) <= 1e-5 and _rel(db3, refb) <= 1e-5
    assert e1 >= 10 * e3                                          # the single pass is the TF32 path, not this one
    with pytest.raises(RuntimeError):
        ops.FC(Xh, Wh, b, X_lo=Xl)                               # both low parts or neither
    with pytest.raises(RuntimeError):
        ops.FC(Xh.bfloat16(), Wh.bfloat16(), out=torch.empty(M, N, device="cuda", dtype=torch.bfloat16), accumulate=True)


def _with_pair(on, fn):
    from nafwebsod_b200 import _lib
    _lib.set_tuning("gemm_pair", 1 if on else 0)
    try:
        return fn()
    finally:
        _lib.set_tuning("gemm_pair", 1)           # the built-in default


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(256, 256, 64), (300, 520, 200), (1000, 304, 512), (4000, 4096, 1024), (2000, 8192, 1568)])
def test_fc_cta_pair_equals_single_cta(shape, dtype):
    """The CTA-pair form of the GEMM (tcgen05 cta_group::2: a (2,1,1) cluster owns a 256 x 256 tile, each CTA stages its 128
    rows of A and half of the B tile, the even CTA issues M = 256 MMAs for both; tuning knob gemm_pair) walks K in the same
    order with the same instruction shape per output element, so forward (bias + ReLU + dropout mask), dX (gated) and dW / db
    -- K-major and MN-major operands, ragged M / N / K edges, stacked launches -- must equal the one-CTA kernel bit for bit."""
    ops = _ops()
    M, N, K = shape
    X, W, b, mask = _data(M, N, K, dtype, seed=7)
    g = torch.Generator(device="cuda").manual_seed(11)
    dY = (torch.randn(M, N, device="cuda", generator=g) * 0.1).to(dtype)
    act = (torch.rand(M, K, device="cuda", generator=g) < 0.5).to(dtype)

    def run():
        y = ops.FC(X, W, b, relu=True, dropout_mask=mask, out_dtype=torch.float32)
        dx = ops.FCGradientX(dY, W, act_below=act, dropout=True, out_dtype=torch.float32)
        dw, db = ops.FCGradientW(dY, X)
        # two stacks in one launch (3-d operands)
        Xs, Ws = torch.stack([X, X.flip(0)]), torch.stack([W, W.flip(0)])
        ys = ops.FC(Xs, Ws, torch.stack([b, b]), relu=True, out_dtype=torch.float32)
        torch.cuda.synchronize()
        return y, dx, dw, db, ys
    one = _with_pair(False, run)
    two = _with_pair(True, run)
    for name, a, c in zip(("fwd", "dX", "dW", "db", "stacked fwd"), one, two):
        if name == "db":      # column sums by atomically accumulated partials (colsum_kernel, not the GEMM): equal up to fp32 order
            assert _rel(a, c) <= 1e-5
        else:
            assert torch.equal(a, c), (name, (a - c).abs().max().item())
    ref = torch.relu(X.float() @ W.float().T + b) * 2.0 * mask.float()
    assert _rel(two[0], ref) <= (BF16_TOL if dtype == torch.bfloat16 else TF32_TOL)


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(256, 256, 64), (300, 520, 200), (1000, 304, 512), (4000, 1024, 4096), (2000, 40, 4096)])
def test_fc_tma_store_epilogue_equals_register_stores(shape, dtype, pair):
    """Plain fp32 outputs (the weight gradients; the partial passes of the three-pass fp32 path) may leave through the TMA:
    every epilogue warp stages its 32 x 32 chunk in 128B-swizzled shared memory and one lane issues a bulk tensor store, or a
    bulk reduce-add when accumulating (tuning knob gemm_tma_store).  Same values, same fp32 adds: bit-identical to the
    register-store epilogue -- dW plain, accumulated onto a prior value, two stacks per launch, ragged edges (the tensor map
    clips), on the one-CTA and the CTA-pair kernel."""
    from nafwebsod_b200 import _lib
    ops = _ops()
    M, N, K = shape
    X, W, b, mask = _data(M, N, K, dtype, seed=13)
    g = torch.Generator(device="cuda").manual_seed(17)
    dY = (torch.randn(M, N, device="cuda", generator=g) * 0.1).to(dtype)
    prior = torch.randn(N, K, device="cuda", generator=g) * 0.01

    def run():
        dw, _ = ops.FCGradientW(dY, X, want_db=False)
        acc = prior.clone()
        ops.FCGradientW(dY, X, dW=acc, want_db=False, accumulate=True)
        dws = torch.full((2, N, K), float("nan"), device="cuda")
        ops.FCGradientW(torch.stack([dY, dY.flip(0)]), torch.stack([X, X.flip(0)]), dW=dws, want_db=False)
        y = ops.FC(X, W, out_dtype=torch.float32)                   # forward without bias / activation: same epilogue
        torch.cuda.synchronize()
        return dw, acc, dws, y
    res = []
    try:
        _lib.set_tuning("gemm_pair", pair)
        for tma in (0, 1):
            _lib.set_tuning("gemm_tma_store", tma)
            res.append(run())
    finally:
        _lib.set_tuning("gemm_tma_store", 1)      # the built-in defaults
        _lib.set_tuning("gemm_pair", 1)
    for name, a, c in zip(("dW", "dW accumulated", "stacked dW", "plain fwd"), res[0], res[1]):
        assert torch.equal(a, c), (name, (a - c).abs().max().item())
    assert _rel(res[1][0], dY.float().T @ X.float()) <= (BF16_TOL if dtype == torch.bfloat16 else TF32_TOL)
