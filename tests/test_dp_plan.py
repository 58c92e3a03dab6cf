"""Host arithmetic of the data-parallel exchange that needs neither a GPU nor a process group: the bucket plan of the
bench's parameter layout splits into 16-byte aligned per-rank slices for 2, 4 and 8 ranks (so `sync=auto` can select the
peer path there), the fused-scatter precondition, and where P2PExchange.owner_ptrs() points."""
import pytest

from nafwebsod_b200 import dp


def bench_layout(C=20, S=2, H=4096, D=25088):
    """heads.WeblyHeadModel._alloc_params for the bench model, in elements."""
    r = lambda v: (v + 63) // 64 * 64
    n_w6 = S * H * D
    off = r(n_w6)
    off = r(off + S * H * H)
    n_weights = r(off + S * 2 * C * H)
    C2p = (2 * C + 7) // 8 * 8
    n_total = n_weights
    for n in (S * H, S * H, S * C2p):
        n_total = r(n_total + n)
    return n_w6, S * H, D, n_weights, n_total


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("C", [20, 80])
def test_bench_buckets_split_into_aligned_slices(world, C):
    n_w6, rows, cols, n_weights, n_total = bench_layout(C=C)
    plan = dp.bucket_plan(n_w6, rows, cols, n_weights, n_total, panels=4)
    assert plan[0][0] == 0 and sum(n for _, n, _ in plan) == n_total
    covered = 0
    for off, n, tag in plan:
        assert off == covered or tag != "fc6_panel"
        covered = off + n
        assert dp.slices_aligned(n, world), (tag, n, world)
        so, sn = dp.rank_slice(off, n, world, world - 1)
        assert so + sn == off + n and (so * 4) % 16 == 0 and (so * 2) % 16 == 0
    assert covered == n_total
    # the GEMM-fused scatter needs whole 128-row tiles per owner in every fc6 panel
    panel_rows = [n // cols for _, n, tag in plan if tag == "fc6_panel"]
    assert sum(panel_rows) == rows and all(r % (128 * world) == 0 for r in panel_rows)


def test_owner_ptrs_address_the_owners_slot_for_this_rank():
    class FakeFlat:
        def data_ptr(self):
            return 1 << 20

    ex = dp.P2PExchange.__new__(dp.P2PExchange)
    ex.world, ex.rank, ex.flat = 4, 2, FakeFlat()
    ex.peer_stage = [0x10000000 * (k + 1) for k in range(4)]
    offset, length = 4096, 4 * 1000
    n = length // 4
    ptrs = ex.owner_ptrs(offset, length)
    assert ptrs[2] == (1 << 20) + 4 * (offset + 2 * n)                      # own slice: the local gradient buffer
    for k in (0, 1, 3):                                                     # owner k's staging, slot of rank 2
        assert ptrs[k] == ex.peer_stage[k] + 4 * (offset + 2 * n)
