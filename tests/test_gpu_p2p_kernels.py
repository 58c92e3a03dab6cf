"""Single-GPU checks of the peer-exchange kernels of na-fwebsod_b200/csrc/p2p.cu (the kernels only see addresses; between
GPUs they run in tests/test_gpu_zzzz_dp_*.py): the SM-driven scatter moves exactly the requested bytes
to every destination and then publishes the sequence number; the wait kernel returns once the flags carry it and reports
a time-out through the status word; the owner's update kernel skips its work once that word is set."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from nafwebsod_b200 import ops
    return ops


@pytest.mark.parametrize("npeers,nbytes", [(1, 16), (3, 8192), (7, 8192 * 5 + 4080), (2, 3 * 1024 * 1024 + 16), (7, 25690112 // 8)])
def test_scatter_moves_the_bytes_then_publishes(npeers, nbytes):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(nbytes % 9973)
    srcs = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device="cuda", generator=g) for _ in range(npeers)]
    pad = 256
    dsts = [torch.full((nbytes + 2 * pad,), 0xAB, dtype=torch.uint8, device="cuda") for _ in range(npeers)]
    flags = torch.zeros(npeers + 1, dtype=torch.int32, device="cuda")
    fptr = [flags.data_ptr() + 4 * i for i in range(npeers + 1)]
    for seq in (1, 2):                           # twice: the completion counter of the slot must reset itself
        ops.p2p_scatter([s.data_ptr() for s in srcs], [d.data_ptr() + pad for d in dsts], nbytes, fptr, seq, 5)
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.p2p_wait(flags, seq, 2000, status)
        torch.cuda.synchronize()
        assert int(status.item()) == 0 and bool((flags == seq).all())
        for s, d in zip(srcs, dsts):
            assert torch.equal(d[pad:pad + nbytes], s)
            assert bool((d[:pad] == 0xAB).all()) and bool((d[pad + nbytes:] == 0xAB).all())      # nothing outside the range
        for s in srcs:
            s.add_(1)


def test_flag_only_launch_and_wait_timeout():
    ops = _ops()
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.p2p_scatter([], [], 0, [flags.data_ptr() + 4 * i for i in range(3)], 7, 9)     # no bytes, flags only (a replicated bucket's operand leg)
    torch.cuda.synchronize()
    assert flags.tolist() == [7, 7, 7, 0]
    ops.p2p_wait(flags, 7, 50, status)           # the fourth flag never arrives: the watchdog fires after 50 ms
    torch.cuda.synchronize()
    assert int(status.item()) == 1


def test_update_is_skipped_once_the_watchdog_word_is_set():
    ops = _ops()
    n = 4096
    g = [torch.randn(n, device="cuda") for _ in range(3)]
    m, p = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    shadow = p.to(torch.bfloat16)
    lr = torch.tensor([1e-2], device="cuda")
    m0, p0, s0 = m.clone(), p.clone(), shadow.clone()
    word = torch.ones(1, dtype=torch.int32, device="cuda")
    ops.ACMWeightDecayMomentumSGDUpdateReduce(g, m, lr, p, gpu_num=3, weight_decay=5e-4, iter_count=2, p_shadow=shadow, abort_flag=word)
    torch.cuda.synchronize()
    assert torch.equal(m, m0) and torch.equal(p, p0) and torch.equal(shadow.view(torch.int16), s0.view(torch.int16))
    word.zero_()
    ops.ACMWeightDecayMomentumSGDUpdateReduce(g, m, lr, p, gpu_num=3, weight_decay=5e-4, iter_count=2, p_shadow=shadow, abort_flag=word)
    m1, p1 = m0.clone(), p0.clone()
    ops.ACMWeightDecayMomentumSGDUpdate((g[0] + g[1]) + g[2], m1, lr, p1, None, gpu_num=3, weight_decay=5e-4, iter_count=2)
    torch.cuda.synchronize()
    assert torch.equal(m, m1) and torch.equal(p, p1)
