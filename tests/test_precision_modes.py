"""Host logic of the head's three precisions (no GPU): which (dtype, precision) pairs exist, what the three-pass fp32 path
allocates, and that it refuses the data-parallel schedules that would leave the low parts of the parameters stale."""
import pytest
import torch


def _model(**kw):
    from nafwebsod_b200.heads import WeblyHeadModel
    return WeblyHeadModel(5, 8, 7, 16, device="cpu", **kw)


def test_precision_follows_the_storage_type():
    assert _model(dtype=torch.bfloat16).precision == "bf16"
    m = _model(dtype=torch.float32)
    assert m.precision == "tf32" and m.tf32 and not m.x3 and m.flat_lo is None and m.wl == {}
    m = _model(dtype=torch.float32, precision="fp32")
    assert m.precision == "fp32" and m.tf32 and m.x3
    # the low parts mirror the operand shadow: one flat float32 buffer, the same views
    assert m.flat_lo.dtype == torch.float32 and m.flat_lo.numel() == m.flat_lp.numel() == m.n_total
    assert set(m.wl) == set(m.w) and all(m.wl[k].shape == m.w[k].shape for k in m.w)
    for bad in (dict(dtype=torch.bfloat16, precision="tf32"), dict(dtype=torch.bfloat16, precision="fp32"),
                dict(dtype=torch.float32, precision="bf16"), dict(dtype=torch.float32, precision="fp64"),
                dict(dtype=torch.float16)):
        with pytest.raises(RuntimeError):
            _model(**bad)


def test_three_pass_path_needs_replicated_masters_across_ranks(monkeypatch):
    """The sharded and peer schedules send only the operand (high-part) shadow of a slice to the other ranks; the fp32 path
    re-splits the fp32 MASTERS every step, so across ranks it runs on the reference's all-reduce schedule only."""
    import torch.distributed as dist
    from nafwebsod_b200 import dp
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 2)
    monkeypatch.setattr(dist, "get_rank", lambda group=None: 0)
    m = _model(dtype=torch.float32, precision="fp32")
    for sync in ("sharded", "p2p", "auto"):
        with pytest.raises(RuntimeError, match="allreduce"):
            dp.DataParallelHead(m, sync=sync)


def test_split_operand_products_need_both_low_parts():
    from nafwebsod_b200 import ops
    a = torch.zeros(4, 8)
    for fn, kw in ((ops.FC, dict(X_lo=a)), (ops.FC, dict(W_lo=a)), (ops.FCGradientX, dict(dY_lo=a)), (ops.FCGradientW, dict(X_lo=a))):
        with pytest.raises(RuntimeError, match="both"):
            fn(a, a, **kw)
    with pytest.raises(RuntimeError):                      # and there is no CPU path behind them
        ops.split_tf32(a)
