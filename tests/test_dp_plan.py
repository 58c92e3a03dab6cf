"""Host arithmetic of the data-parallel exchange that needs neither a GPU nor a process group: the bucket plan of the
bench's parameter layout splits into 16-byte aligned per-rank slices for 2, 4 and 8 ranks (so `sync=auto` can select the
peer path there), and the split of the bias block into the bucket fc6 needs and the rest."""
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
    panel_rows = [n // cols for _, n, tag in plan if tag == "fc6_panel"]
    assert sum(panel_rows) == rows


def test_bias_block_splits_into_what_fc6_needs_and_the_rest():
    n_w6, rows, cols, n_weights, n_total = bench_layout()
    plan = dp.bucket_plan(n_w6, rows, cols, n_weights, n_total, panels=4, n_bias_fc6=8192)
    tags = [t for _, _, t in plan]
    assert tags == ["fc6_panel"] * 4 + ["small_weights", "biases_fc6", "biases"]
    assert sum(n for _, n, _ in plan) == n_total and plan[-2][:2] == (n_weights, 8192) and plan[-1][0] == n_weights + 8192
    for world in (2, 4, 8):                      # both bias buckets are replicated, whatever their alignment
        assert not dp.bucket_is_sliced(plan[-1][1], "biases", world) and not dp.bucket_is_sliced(8192, "biases_fc6", world)
        assert dp.bucket_is_sliced(plan[0][1], "fc6_panel", world)
