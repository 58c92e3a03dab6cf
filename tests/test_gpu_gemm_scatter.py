"""na-fwebsod_b200/csrc/gemm_scatter.cu: the fc6 weight-gradient GEMM whose epilogue stores every tile straight into the
buffer of the rank that owns those rows (the send leg of the data-parallel reduce-scatter that replaces the reference's
NCCLAllreduce, detectron/modeling/optimizer_wsl.py:52-72).  Parity bar: dW bit-identical to the verified stand-alone
tcgen05 GEMM (same main loop, same accumulation order); db -- float atomicAdd column sums in both -- to rounding noise.
One GPU suffices: the kernel only sees destination addresses (tests/test_gpu_zzzz_dp_2gpu.py runs it over peer-mapped
memory between GPUs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("owners,rows_per_owner,N", [(2, 128, 256), (4, 128, 512), (8, 256, 2048), (3, 128, 300)])
def test_scatter_epilogue_lands_every_row_at_its_owner(dtype, owners, rows_per_owner, N):
    """W separate destination buffers stand for the owners' staging areas."""
    ops = _ops()
    M, K = 320, 520
    dY, X = _operands(M, N, K, dtype, seed=N)
    if N % 8:                                             # a row pitch of 16 bytes for the plain GEMM's operands
        dYp = torch.zeros(M, (N + 7) // 8 * 8, device="cuda", dtype=dtype)
        dYp[:, :N] = dY
        dY = dYp[:, :N]
    dW, db = ops.FCGradientW(dY, X)
    bufs = [torch.full((rows_per_owner, K), float("nan"), device="cuda") for _ in range(owners)]
    db1 = torch.empty(N, device="cuda")
    ops.FCGradientWScatter(dY, X, [b.data_ptr() for b in bufs], rows_per_owner, K, db=db1)
    torch.cuda.synchronize()
    for k, b in enumerate(bufs):
        lo, hi = k * rows_per_owner, min(N, (k + 1) * rows_per_owner)
        assert torch.equal(b[:hi - lo], dW[lo:hi]), "owner %d" % k
        assert torch.isnan(b[hi - lo:]).all()                            # rows beyond N are never written
    scale = db.abs().max().item()
    assert (db1 - db).abs().max().item() <= 1e-5 * scale                 # atomics: summation order only
    with pytest.raises(RuntimeError):
        ops.FCGradientWScatter(dY, X, [b.data_ptr() for b in bufs], 100, K)     # an output tile must have one owner


def test_scatter_full_size_fc6_panel():
    """One fc6 row panel of config 2 (4000 RoIs, 2048 of the 8192 stacked rows, K = 25088) over 8 owners."""
    ops = _ops()
    M, N, K = 4000, 2048, 25088
    dY, X = _operands(M, N, K, torch.bfloat16, seed=1)
    dW, _ = ops.FCGradientW(dY, X, want_db=False)
    bufs = [torch.empty(N // 8, K, device="cuda") for _ in range(8)]
    ops.FCGradientWScatter(dY, X, [b.data_ptr() for b in bufs], N // 8, K)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(bufs), dW)
