"""EXPERIMENTAL kernels of na-fwebsod_b200/csrc/gemm_fused.cu (fc6 weight gradient fused with the SGD update, or with the
scatter to the owner ranks).  They are opt-in in the product (NAWSOD_FUSED_SGD=1 / NAWSOD_P2P_FUSED_SCATTER=1) and were
written after this round's GPU budget was spent, so these tests only run when NAWSOD_EXPERIMENTAL=1
(tools/gpu_round2_first.sh sets it): a kernel that has never executed must not be able to turn the regular suite red.
Parity bar: bit-identical to the verified stand-alone kernels (tcgen05 GEMM -> ACMWeightDecayMomentumSGDUpdate)."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NAWSOD_EXPERIMENTAL") != "1", reason="experimental kernels: set NAWSOD_EXPERIMENTAL=1")]


def _ops():
    from nafwebsod_b200 import ops
    return ops


def _operands(M, N, K, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dY = (torch.randn(M, N, device="cuda", generator=g) * 0.05).to(dtype)
    X = torch.rand(M, K, device="cuda", generator=g).to(dtype)
    if dtype == torch.float32:
        ops = _ops()
        dY, X = ops.round_to_tf32(dY), ops.round_to_tf32(X)
    return dY, X


def _state(N, K, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    p = torch.randn(N, K, device="cuda", generator=g) * 0.01
    m = torch.randn(N, K, device="cuda", generator=g) * 0.001
    shadow = torch.zeros(N, K, device="cuda", dtype=dtype)
    return m, p, shadow


# (RoIs, rows of W, columns of W): whole tiles, ragged rows / columns (TMA zero fill + guarded epilogue), one tile only
SHAPES = [(512, 256, 512), (300, 200, 520), (4000, 384, 1024), (64, 128, 256)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("iter_count", [0, 3])
@pytest.mark.parametrize("keep_grad", [False, True])
def test_fused_sgd_equals_gemm_then_update(shape, dtype, iter_count, keep_grad):
    ops = _ops()
    M, N, K = shape
    dY, X = _operands(M, N, K, dtype, seed=M + N)
    lr = torch.tensor([1e-3], device="cuda")
    kw = dict(momentum=0.9, gpu_num=1, lr_mult=1.0, weight_decay=5e-4, iter_count=iter_count)
    # verified path: GEMM, then the stand-alone update
    m0, p0, s0 = _state(N, K, dtype, seed=K)
    dW0, db0 = ops.FCGradientW(dY, X)
    ops.ACMWeightDecayMomentumSGDUpdate(dW0.clone(), m0, lr, p0, None, p_shadow=s0, **kw)
    # fused
    m1, p1, s1 = _state(N, K, dtype, seed=K)
    dW1 = torch.full((N, K), float("nan"), device="cuda") if keep_grad else None
    db1 = torch.empty(N, device="cuda")
    ops.FCGradientWSGD(dY, X, m1, lr, p1, s1, dW=dW1, db=db1, **kw)
    torch.cuda.synchronize()
    assert torch.equal(m1, m0) and torch.equal(p1, p0), "momentum / parameter differ from GEMM + stand-alone update"
    assert torch.equal(s1.view(torch.int16 if dtype == torch.bfloat16 else torch.int32),
                       s0.view(torch.int16 if dtype == torch.bfloat16 else torch.int32)), "operand shadow differs"
    assert torch.equal(db1, db0)
    if keep_grad:
        assert torch.equal(dW1, dW0)


def test_fused_sgd_accumulates_into_the_gradient_buffer():
    ops = _ops()
    M, N, K = 256, 128, 512
    dY, X = _operands(M, N, K, torch.bfloat16, seed=5)
    lr = torch.tensor([1e-3], device="cuda")
    base = torch.randn(N, K, device="cuda")
    m0, p0, s0 = _state(N, K, torch.bfloat16, seed=9)
    g0 = base.clone()
    ops.FCGradientW(dY, X, dW=g0, want_db=False, accumulate=True)
    ops.ACMWeightDecayMomentumSGDUpdate(g0.clone(), m0, lr, p0, None, p_shadow=s0, momentum=0.9, weight_decay=5e-4, iter_count=2)
    m1, p1, s1 = _state(N, K, torch.bfloat16, seed=9)
    g1 = base.clone()
    ops.FCGradientWSGD(dY, X, m1, lr, p1, s1, dW=g1, accumulate=True, momentum=0.9, weight_decay=5e-4, iter_count=2)
    assert torch.equal(g1, g0) and torch.equal(m1, m0) and torch.equal(p1, p0) and torch.equal(s1.view(torch.int16), s0.view(torch.int16))


def test_fused_sgd_rejects_what_it_cannot_do():
    ops = _ops()
    dY, X = _operands(64, 128, 256, torch.bfloat16, seed=1)
    lr = torch.tensor([1e-3], device="cuda")
    m, p, s = _state(128, 256, torch.bfloat16, seed=1)
    with pytest.raises(RuntimeError):
        ops.FCGradientWSGD(dY, X, m, lr, p, s.float())                  # shadow must have the operands' type
    with pytest.raises(RuntimeError):
        ops.FCGradientWSGD(dY, X, m, lr, p, s, accumulate=True)         # accumulate needs dW
    with pytest.raises(RuntimeError):
        ops.FCGradientWSGD(dY, X, m[:64], lr, p, s)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("owners,rows_per_owner,N", [(2, 128, 256), (4, 128, 512), (8, 256, 2048), (3, 128, 300)])
def test_scatter_epilogue_lands_every_row_at_its_owner(dtype, owners, rows_per_owner, N):
    """One GPU, W separate destination buffers standing for the owners' staging areas (on a multi-GPU box they are
    peer-mapped; the kernel only sees addresses)."""
    ops = _ops()
    M, K = 320, 520
    dY, X = _operands(M, N, K, dtype, seed=N)
    dW, db = ops.FCGradientW(dY, X)
    bufs = [torch.full((rows_per_owner, K), float("nan"), device="cuda") for _ in range(owners)]
    db1 = torch.empty(N, device="cuda")
    ops.FCGradientWScatter(dY, X, [b.data_ptr() for b in bufs], rows_per_owner, K, db=db1)
    torch.cuda.synchronize()
    for k, b in enumerate(bufs):
        lo, hi = k * rows_per_owner, min(N, (k + 1) * rows_per_owner)
        assert torch.equal(b[:hi - lo], dW[lo:hi]), "owner %d" % k
        assert torch.isnan(b[hi - lo:]).all()                            # rows beyond N are never written
    assert torch.equal(db1, db)
    with pytest.raises(RuntimeError):
        ops.FCGradientWScatter(dY, X, [b.data_ptr() for b in bufs], 100, K)     # an output tile must have one owner


def test_head_step_with_fused_sgd_equals_the_pipelined_step(monkeypatch):
    """Training steps of the whole head on one GPU, fused fc6 update vs the default (GEMM per panel, stand-alone update on
    the side stream).  After ONE step the weights, their momenta and the operand shadow are bit-identical (the weight
    gradients are deterministic); the biases agree up to the summation order of their atomically accumulated column sums,
    which is also why a second step (whose forward sees those biases) is compared with the tolerance of
    test_gpu_head.py::test_pipelined_update_matches_plain_schedule."""
    from nafwebsod_b200.dp import DataParallelHead
    from nafwebsod_b200.heads import WeblyHeadModel
    from oracle import nawsod_oracle as O            # inputs only (the checker's synthetic data)

    def run(fused, steps):
        monkeypatch.setenv("NAWSOD_FUSED_SGD", "1" if fused else "0")
        model = WeblyHeadModel(21, 64, 7, 512, noise=True, dtype=torch.bfloat16)
        g = torch.Generator(device="cuda").manual_seed(2)
        model.flat_param[:model.n_weights].normal_(0.0, 0.01, generator=g)
        model.sync_shadow()
        model.UpdateWorkspaceLr(1e-2)
        dp = DataParallelHead(model, fc6_panels=4, sync="auto")
        assert (dp._fused_mode() == "sgd") == fused
        X = torch.from_numpy(O.synth_conv5(2, 64, 20, 25, seed=0)).cuda()
        rois = np.concatenate([O.synth_rois(150, 320, 400, b, seed=1 + b) for b in range(2)])
        obn = (np.random.default_rng(2).random(300) + 1).astype(np.float32)
        L = np.zeros((2, 20), np.float32); L[0, 3] = 1; L[1, 7] = 1
        model.FeedBlobs(X, torch.from_numpy(rois).cuda(), torch.from_numpy(obn).cuda(), torch.from_numpy(L).cuda(),
                        torch.tensor([0, 150, 300], dtype=torch.int32, device="cuda"), x_layout="NCHW")
        for it in range(steps):
            dp.step(dropout_seed=it + 1)
        dp.flush()
        torch.cuda.synchronize()
        return model.flat_param.clone(), model.flat_mom.clone(), model.flat_lp.clone(), model.n_weights

    (pa, ma, la, nw), (pb, mb, lb, _) = run(False, 1), run(True, 1)
    assert torch.equal(pa[:nw], pb[:nw]) and torch.equal(ma[:nw], mb[:nw])
    assert torch.equal(la[:nw].view(torch.int16), lb[:nw].view(torch.int16))
    upd = ma.abs().max().item()
    assert upd > 0 and (pa - pb).abs().max().item() <= 1e-5 * upd and (ma - mb).abs().max().item() <= 1e-5 * upd
    (pa, ma, la, nw), (pb, mb, lb, _) = run(False, 2), run(True, 2)
    upd = ma.abs().max().item()
    assert (pa - pb).abs().max().item() <= 1e-3 * upd and (ma - mb).abs().max().item() <= 1e-3 * upd
    assert (la.float() - lb.float()).abs().max().item() <= 2e-2 * la.float().abs().max().item()


def test_bias_gradients_on_a_side_stream(monkeypatch):
    """NAWSOD_BIAS_SIDE_STREAM=1: the column sums run beside the weight-gradient GEMMs; same gradients up to the atomics'
    summation order (the sums are float atomicAdd over row blocks in both schedules)."""
    from nafwebsod_b200.heads import WeblyHeadModel
    from oracle import nawsod_oracle as O            # inputs only

    def run(side):
        monkeypatch.setenv("NAWSOD_BIAS_SIDE_STREAM", "1" if side else "0")
        model = WeblyHeadModel(21, 64, 7, 512, noise=True, dtype=torch.bfloat16)
        g = torch.Generator(device="cuda").manual_seed(2)
        model.flat_param[:model.n_weights].normal_(0.0, 0.01, generator=g)
        model.sync_shadow()
        X = torch.from_numpy(O.synth_conv5(1, 64, 20, 25, seed=0)).cuda()
        rois = torch.from_numpy(O.synth_rois(300, 320, 400, seed=1)).cuda()
        obn = torch.from_numpy((np.random.default_rng(2).random(300) + 1).astype(np.float32)).cuda()
        L = torch.zeros(1, 20, device="cuda"); L[0, 3] = 1
        model.FeedBlobs(X, rois, obn, L, x_layout="NCHW")
        model.RunTrainStep(dropout_seed=1, fc6_panels=4)
        torch.cuda.synchronize()
        return model.flat_grad.clone(), model.n_weights

    (a, nw), (b, _) = run(False), run(True)                          # weights come from the same GEMMs: identical
    assert torch.equal(a[:nw], b[:nw])
    scale = a[nw:].abs().max().item()
    assert scale > 0 and (a[nw:] - b[nw:]).abs().max().item() <= 1e-5 * scale


@pytest.mark.parametrize("knobs", [{"pool_skip_idle": 1}, {"pool_prefetch_roi": 1}, {"pool_skip_idle": 1, "pool_prefetch_roi": 1},
                                   {"pool_lean": 1}])
@pytest.mark.parametrize("argmax", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_pool_experimental_variants_are_bit_identical(dtype, argmax, knobs):
    """pool_skip_idle = 1: CTAs whose chunk holds no RoI of their image return before staging the map; pool_prefetch_roi = 1:
    a warp claims and loads its next RoI before working on the current one; pool_lean = 1: both plus the seven bin rows
    unrolled (rows3_bins).  Two images, RoIs grouped by image (the loader's order) and, as a torture case, interleaved, plus
    the config-3 RoI mixture (tiny, clipped, outside, inverted boxes -> the empty-bin path): values and argmax equal the
    default kernel."""
    import nafwebsod_b200 as pkg
    from oracle import nawsod_oracle as O            # inputs only
    ops = _ops()
    X = torch.from_numpy(O.synth_conv5(2, 512, 38, 50, seed=0)).cuda()
    Xcl = ops.to_channels_last(X, dtype)
    grouped = np.concatenate([O.synth_rois(1000, 608, 800, b, seed=1 + b) for b in range(2)])
    inter = grouped[np.random.default_rng(3).permutation(2000)]
    obn = torch.from_numpy((np.random.default_rng(9).random(2000) + 1).astype(np.float32)).cuda()
    rng = np.random.default_rng(4)
    mixed = grouped.copy()
    kind = rng.integers(0, 6, 2000)
    mixed[kind == 0, 3:] = mixed[kind == 0, 1:3] + rng.integers(0, 30, (int((kind == 0).sum()), 2))      # tiny
    mixed[kind == 1, 1:] = (-60, -35, 200, 150)                                                        # overhanging
    mixed[kind == 2, 1:] = (900, 700, 1000, 800)                                                       # outside the map
    mixed[kind == 3, 1:] = mixed[kind == 3][:, [3, 4, 1, 2]]                                           # inverted
    for rois in (grouped, inter, mixed.astype(np.float32)):
        r = torch.from_numpy(np.ascontiguousarray(rois)).cuda()
        kw = dict(spatial_scale=1 / 16, is_test=not argmax, boost=obn, x_layout="NHWC", y_layout="NHWC", out_dtype=dtype)
        Y0, A0 = ops.RoIPoolF(Xcl, r, **kw)
        for k, v in knobs.items():
            pkg.set_tuning(k, v)
        try:
            Y1, A1 = ops.RoIPoolF(Xcl, r, **kw)
        finally:
            for k in knobs:
                pkg.set_tuning(k, 0)
        assert torch.equal(Y0.view(torch.int16 if dtype == torch.bfloat16 else torch.int32),
                           Y1.view(torch.int16 if dtype == torch.bfloat16 else torch.int32))
        assert (A0 is None and A1 is None) or torch.equal(A0, A1)
