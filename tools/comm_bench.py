"""Collective timings for the gradient exchange (run under torchrun, N >= 2):
    NCCL_MAX_CTAS=16 torchrun --nproc-per-node N tools/comm_bench.py [--with-gemm]
Times all_reduce / reduce_scatter (fp32) and all_gather (bf16) of one fc6 gradient panel and of the
whole gradient, alone and next to a persistent tcgen05 GEMM limited to (SMs - NAWSOD_COMM_SMS) CTAs."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nafwebsod_b200 as pkg  # noqa: E402
from nafwebsod_b200 import ops  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = torch.cuda.Stream(priority=-1)
    n_panel = 2048 * 25088
    g = torch.randn(4 * n_panel, device=dev)
    lp = torch.zeros(4 * n_panel, device=dev, dtype=torch.bfloat16)
    M, N, K = 4000, 2048, 25088
    dY = torch.randn(M, N, device=dev).to(torch.bfloat16)
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    dW = torch.empty(N, K, device=dev)
    reserve = int(os.environ.get("NAWSOD_COMM_SMS", "16"))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    def timed(fn, iters=10, side=None):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def rs(n):
        s = n // world
        dist.reduce_scatter_tensor(g[rank * s:(rank + 1) * s], g[:n])

    def ag(n):
        s = n // world
        dist.all_gather_into_tensor(lp[:n], lp[rank * s:(rank + 1) * s])

    out = []
    for label, n in (("panel 205MB", n_panel), ("all fc6 822MB", 4 * n_panel)):
        t = timed(lambda: dist.all_reduce(g[:n]))
        out.append("all_reduce fp32 %s: %.3f ms  busbw %.0f GB/s" % (label, t, 2 * (world - 1) / world * n * 4 / t / 1e6))
        t = timed(lambda: rs(n))
        out.append("reduce_scatter fp32 %s: %.3f ms  busbw %.0f GB/s" % (label, t, (world - 1) / world * n * 4 / t / 1e6))
        t = timed(lambda: ag(n))
        out.append("all_gather bf16 %s: %.3f ms  busbw %.0f GB/s" % (label, t, (world - 1) / world * n * 2 / t / 1e6))
    # GEMM alone (full grid, limited grid), then GEMM + concurrent reduce-scatter of a previous panel
    gemm = lambda: ops.FCGradientW(dY, A, dW=dW, want_db=False)
    t_full = timed(gemm)
    pkg.set_tuning("gemm_max_ctas", sms - reserve)
    t_lim = timed(gemm)
    out.append("fc6 dW panel GEMM alone: %.3f ms (148 CTAs), %.3f ms (%d CTAs)" % (t_full, t_lim, sms - reserve))

    def both():
        ev = torch.cuda.Event(); ev.record()
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            rs(n_panel)
            ag(n_panel)
        gemm()
    t = timed(both, side=comm)
    out.append("GEMM (%d CTAs) || reduce_scatter+all_gather of a panel on a side stream: %.3f ms per pair" % (sms - reserve, t))
    pkg.set_tuning("gemm_max_ctas", 0)
    t = timed(both, side=comm)
    out.append("GEMM (148 CTAs) || reduce_scatter+all_gather of a panel on a side stream: %.3f ms per pair" % t)
    if rank == 0:
        print("world %d NCCL_MAX_CTAS=%s NAWSOD_COMM_SMS=%d" % (world, os.environ.get("NCCL_MAX_CTAS"), reserve))
        print("\n".join(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
