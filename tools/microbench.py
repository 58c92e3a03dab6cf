"""Per-kernel timings on one GPU (CUDA events, L2 flushed between iterations).
    python tools/microbench.py [pool|poolbwd|mil|sgd|gemm ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nafwebsod_b200 as pkg  # noqa: E402
from nafwebsod_b200 import ops  # noqa: E402
from oracle import nawsod_oracle as O  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
_flush = None


def timeit(fn, iters=20, warmup=3, flush=True):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_pool():
    for (N, R, H, W, dt, train) in [(1, 2000, 38, 50, torch.float32, True), (2, 4000, 38, 50, torch.float32, True),
                                    (2, 4000, 38, 50, torch.bfloat16, True), (1, 2000, 38, 50, torch.bfloat16, False),
                                    (1, 4000, 75, 125, torch.float32, True), (1, 4000, 75, 125, torch.bfloat16, False)]:
        C = 512
        X = torch.from_numpy(O.synth_conv5(N, C, H, W)).cuda()
        Xcl = ops.to_channels_last(X, dt)
        rois = torch.from_numpy(np.concatenate([O.synth_rois(R // N, H * 16, W * 16, b, seed=1 + b) for b in range(N)])).cuda()
        boost = torch.rand(R, device="cuda") + 1
        es = 4 if dt == torch.float32 else 2
        alg = R * (C * 49 * es + (C * 49 * 4 if train else 0) + 20) + N * C * H * W * es
        for knobs in [dict(), dict(pool_threads=512), dict(pool_slab_bytes=64 * 1024, pool_threads=512),
                      dict(pool_slab_bytes=64 * 1024, pool_threads=1024), dict(pool_slab_bytes=32 * 1024, pool_threads=512),
                      dict(pool_chunks=5), dict(pool_chunks=14), dict(pool_force_global=1)]:
            for k, v in knobs.items():
                pkg.set_tuning(k, v)
            med, best = timeit(lambda: ops.RoIPoolF(Xcl, rois, boost=boost, is_test=not train, x_layout="NHWC", y_layout="NHWC"))
            pkg.set_tuning("pool_slab_bytes", 200 * 1024); pkg.set_tuning("pool_force_global", 0)
            pkg.set_tuning("pool_threads", 0); pkg.set_tuning("pool_chunks", 0)
            print("pool N=%d R=%d %dx%d %s train=%d %s: med %.1f us best %.1f us  %.0f GB/s (%.1f%% of %.0f)  %.2f M RoIs/s" % (
                N, R, H, W, str(dt)[6:], train, knobs, med * 1e3, best * 1e3, alg / med / 1e6, 100 * alg / med / 1e6 / PEAKS["hbm_gbs"],
                PEAKS["hbm_gbs"], R / med / 1e3), flush=True)


def bench_pool2():
    """The bench.py shapes only: default (rows2 kernel, automatic chunking), the first bin-row kernel, and a chunk sweep."""
    for (N, R, H, W, dt, train) in [(2, 4000, 38, 50, torch.bfloat16, False), (2, 4000, 38, 50, torch.bfloat16, True),
                                    (1, 2000, 38, 50, torch.float32, True), (2, 4000, 38, 50, torch.float32, False),
                                    (1, 4000, 43, 57, torch.bfloat16, False)]:
        C = 512
        X = torch.from_numpy(O.synth_conv5(N, C, H, W)).cuda()
        Xcl = ops.to_channels_last(X, dt)
        rois = torch.from_numpy(np.concatenate([O.synth_rois(R // N, H * 16, W * 16, b, seed=1 + b) for b in range(N)])).cuda()
        boost = torch.rand(R, device="cuda") + 1
        es = 4 if dt == torch.float32 else 2
        alg = R * (C * 49 * es + (C * 49 * 4 if train else 0) + 20) + N * C * H * W * es
        for knobs in [dict(), dict(pool_skip_idle=0), dict(pool_rows2=0), dict(pool_chunks=9), dict(pool_chunks=18), dict(pool_chunks=27),
                      dict(pool_chunks=36)]:
            for k, v in knobs.items():
                pkg.set_tuning(k, v)
            med, best = timeit(lambda: ops.RoIPoolF(Xcl, rois, boost=boost, is_test=not train, x_layout="NHWC", y_layout="NHWC"))
            pkg.set_tuning("pool_rows2", -1); pkg.set_tuning("pool_threads", 0); pkg.set_tuning("pool_chunks", 0)
            pkg.set_tuning("pool_skip_idle", 1)
            print("pool N=%d R=%d %dx%d %s train=%d %s: med %.1f us best %.1f us  %.0f GB/s (%.1f%% of %.0f)  %.2f M RoIs/s" % (
                N, R, H, W, str(dt)[6:], train, knobs, med * 1e3, best * 1e3, alg / med / 1e6, 100 * alg / med / 1e6 / PEAKS["hbm_gbs"],
                PEAKS["hbm_gbs"], R / med / 1e3), flush=True)


def bench_poolbwd():
    N, R, H, W, C = 2, 4000, 38, 50, 512
    X = torch.from_numpy(O.synth_conv5(N, C, H, W)).cuda()
    rois = torch.from_numpy(np.concatenate([O.synth_rois(R // N, H * 16, W * 16, b, seed=1 + b) for b in range(N)])).cuda()
    for lay in ("NHWC", "NCHW"):
        Y, A = ops.RoIPoolF(X if lay == "NCHW" else ops.to_channels_last(X), rois, x_layout=lay, y_layout=lay)
        dY = torch.randn_like(Y)
        Xl = X if lay == "NCHW" else ops.to_channels_last(X)
        med, best = timeit(lambda: ops.RoIPoolFGradient(Xl, rois, A, dY, layout=lay))
        alg = R * C * 49 * 8 + N * C * H * W * 4
        print("poolbwd %s: med %.1f us  %.0f GB/s" % (lay, med * 1e3, alg / med / 1e6), flush=True)


def bench_mil():
    for (R, C, B) in [(4000, 20, 2), (2000, 20, 1), (4000, 80, 1), (8000, 80, 2)]:
        rois = torch.from_numpy(np.concatenate([O.synth_rois(R // B, 608, 800, b, seed=b) for b in range(B)])).cuda()
        l = [torch.randn(R, C, device="cuda") for _ in range(4)]
        L = torch.zeros(B, C, device="cuda"); L[:, 3] = 1
        offs = torch.tensor([i * (R // B) for i in range(B)] + [R], dtype=torch.int32, device="cuda")
        for ent in (True, False):
            med, best = timeit(lambda: ops.mil_head(l[0], l[1], rois, offs, L, l[2], l[3], entropy=ent), flush=False)
            print("mil R=%d C=%d B=%d entropy=%d: med %.1f us best %.1f us" % (R, C, B, ent, med * 1e3, best * 1e3), flush=True)


def bench_sgd():
    n = 4096 * 25088
    p, m, g = [torch.randn(n, device="cuda") for _ in range(3)]
    sh = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    lr = torch.tensor([1e-3], device="cuda")
    for shadow in (None, sh):
        med, best = timeit(lambda: ops.ACMWeightDecayMomentumSGDUpdate(g, m, lr, p, None, weight_decay=5e-4, iter_count=3, p_shadow=shadow), iters=10)
        byt = n * (20 + (2 if shadow is not None else 0))
        print("sgd n=%d shadow=%s: med %.1f us  %.0f GB/s (%.1f%%)" % (n, shadow is not None, med * 1e3, byt / med / 1e6, 100 * byt / med / 1e6 / PEAKS["hbm_gbs"]), flush=True)




def bench_gemm():
    def run(name, M, N, K, fn, flops=None):
        med, best = timeit(fn, iters=10, warmup=3, flush=False)
        fl = flops or 2.0 * M * N * K
        print("gemm %-22s M=%d N=%d K=%d: med %.1f us best %.1f us  %.0f TFLOP/s (%.1f%% of %.0f)" % (
            name, M, N, K, med * 1e3, best * 1e3, fl / med / 1e9, 100 * fl / med / 1e9 / PEAKS["bf16_tflops"], PEAKS["bf16_tflops"]), flush=True)
    bf = torch.bfloat16
    for (M, N, K) in [(4000, 8192, 25088), (4000, 4096, 4096), (4096, 4096, 4096), (8192, 8192, 8192)]:
        X = (torch.randn(M, K, device="cuda") * 0.5).to(bf)
        W = (torch.randn(N, K, device="cuda") * 0.02).to(bf)
        b = torch.zeros(N, device="cuda")
        Y = torch.empty(M, N, device="cuda", dtype=bf)
        run("fwd bias+relu", M, N, K, lambda: ops.FC(X, W, b, relu=True, dropout=True, out=Y))
        run("torch.matmul (cuBLAS)", M, N, K, lambda: torch.matmul(X, W.t()))
        if K <= 8192:
            dA = torch.empty(M, K, device="cuda", dtype=bf)
            run("bwd_x", M, N, K, lambda: ops.FCGradientX(Y, W, act_below=X, dropout=True, out=dA))
        dW = torch.empty(N, K, device="cuda")
        run("bwd_w", M, N, K, lambda: ops.FCGradientW(Y, X, dW=dW, want_db=False))
        del X, W, Y, dW
    Xf = torch.randn(4000, 4096, device="cuda"); Wf = torch.randn(4096, 4096, device="cuda") * 0.02
    Yf = torch.empty(4000, 4096, device="cuda")
    run("fwd tf32", 4000, 4096, 4096, lambda: ops.FC(Xf, Wf, None, out=Yf))


def bench_testtime():
    """BASELINE config 5: test-time-augmented inference (5 scales x {orig, hflip} = 10 passes per image, forward only,
    bf16) with a proposal-count sweep, through test_time.im_detect_bbox_aug; then threshold + NMS + limit on the GPU next
    to the CPU oracle (the reference's Cython NMS restated in C)."""
    import time
    from nafwebsod_b200.heads import WeblyHeadModel
    from nafwebsod_b200 import test_time
    from oracle import test_time_oracle as T
    m = WeblyHeadModel(21, 512, 7, 4096, dtype=torch.bfloat16, train=False)
    g = torch.Generator(device="cuda").manual_seed(2)
    m.flat_param[:m.n_weights].normal_(0.0, 0.01, generator=g)
    m.sync_shadow()
    img_h, img_w = 375, 500                                   # a VOC-sized image
    scales = [688, 480, 576, 864, 1200]                       # TEST.SCALE, then TEST.BBOX_AUG.SCALES (yaml:40,52)
    maps = {}
    for s in scales:
        sc = s / float(min(img_h, img_w))
        h, w = int(np.ceil(img_h * sc / 16)), int(np.ceil(img_w * sc / 16))
        maps[s] = (torch.from_numpy(O.synth_conv5(1, 512, h, w, seed=s)).cuda().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), sc)
    order = [(688, True)] + [(s, f) for s in scales[1:] for f in (False, True)] + [(688, False)]
    passes = [(maps[s][0], maps[s][1], img_w if f else None) for s, f in order]
    for R in (500, 1000, 2000, 4000, 8000):
        boxes = torch.from_numpy(O.synth_rois(R, img_h, img_w, 0, seed=R)[:, 1:].copy()).cuda()
        obn = torch.rand(R, device="cuda")
        for sync in (True, False):
            med, best = timeit(lambda: test_time.im_detect_bbox_aug(m, passes, boxes, obn, sync=sync), iters=5, warmup=2, flush=False)
            print("tta R=%d sync=%d: %d passes med %.2f ms  %.2f M RoI-passes/s" % (R, sync, len(passes), med, R * len(passes) / med / 1e3), flush=True)
        scores = test_time.im_detect_bbox_aug(m, passes, boxes, obn)
        med, best = timeit(lambda: ops.nms_and_limit(scores, boxes, score_thresh=1e-9, nms_thresh=0.5, detections_per_im=100), iters=10, flush=False)
        sc_h, bx_h = scores.cpu().numpy(), boxes.cpu().numpy()
        t0 = time.perf_counter()
        T.box_results_with_nms_and_limit(sc_h, bx_h, 21, 1e-9, 0.5, 100)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        print("nms+limit R=%d K=20: GPU med %.3f ms, CPU oracle (1 core) %.1f ms" % (R, med, cpu_ms), flush=True)


def bench_convbody():
    """EXPERIMENTAL frozen VGG16 conv body (csrc/conv_body.cu + the tcgen05 GEMM): whole body and its two kernel kinds on
    flickr-sized images (short side 480 / 688; WSL.DILATION 2).  FLOPs = 2 * 9 * Cin * Cout per output pixel and layer."""
    from nafwebsod_b200.conv_body import VGG16ConvBody, body_ops
    from oracle import conv_body_oracle as CB
    body = VGG16ConvBody(dilation=2)
    body.load_reference_params(CB.synth_params(0))
    for (H, W) in [(480, 640), (688, 912)]:
        body.feed_image(torch.randn(1, 3, H, W, device="cuda") * 50)
        flops, h, w = 0, H, W
        for kind, _, _, a in body_ops(2):
            if kind == "Conv":
                flops += 2 * 9 * a["dim_in"] * a["dim_out"] * h * w
            elif kind == "MaxPool":
                h, w = (h - 2) // a["stride"] + 1, (w - 2) // a["stride"] + 1
        for implicit in (True, False):
            body.implicit = implicit
            med, best = timeit(lambda: body.run(), iters=10)
            print("convbody %dx%d -> conv5 %dx%d [%s]: med %.3f ms best %.3f ms  %.1f GFLOP  %.0f TFLOP/s (%.1f%% of %.0f)" % (
                H, W, h, w, "implicit GEMM (4-D TMA boxes)" if implicit else "patch matrix + FC GEMM", med, best, flops / 1e9, flops / med / 1e9,
                100 * flops / med / 1e9 / PEAKS["bf16_tflops"], PEAKS["bf16_tflops"]), flush=True)
        body.implicit = None
        x = body.blobs["data"]
        for name, cin, cout, hh, ww, dil in (("conv1_2", 64, 64, H, W, 1), ("conv2_2", 128, 128, H // 2, W // 2, 1), ("conv3_3", 256, 256, H // 4, W // 4, 1),
                                             ("conv4_3", 512, 512, H // 8, W // 8, 1), ("conv5_3", 512, 512, H // 8 - 1, W // 8 - 1, 2)):
            xin = torch.relu(torch.randn(1, hh, ww, cin, device="cuda")).to(torch.bfloat16)
            fl = 2.0 * 9 * cin * cout * hh * ww
            m1, _ = timeit(lambda: ops.Conv3x3Relu(xin, body.w[name], body.b[name], dilation=dil, implicit=True), iters=10)
            m0, _ = timeit(lambda: ops.Conv3x3Relu(xin, body.w[name], body.b[name], dilation=dil, implicit=False), iters=10)
            print("  %s %dx%dx%d->%d: implicit %.1f us (%.0f TFLOP/s, %.1f%%)  patch matrix %.1f us" % (
                name, hh, ww, cin, cout, m1 * 1e3, fl / m1 / 1e9, 100 * fl / m1 / 1e9 / PEAKS["bf16_tflops"], m0 * 1e3), flush=True)
        x = torch.randn(1, H // 2, W // 2, 128, device="cuda").to(torch.bfloat16)
        med, _ = timeit(lambda: ops.Im2Col3x3(x))
        nbytes = x.numel() * 2 * 10
        print("  im2col3x3 %dx%dx128: med %.1f us  %.0f GB/s (%.1f%% of %.0f)" % (H // 2, W // 2, med * 1e3, nbytes / med / 1e6,
              100 * nbytes / med / 1e6 / PEAKS["hbm_gbs"], PEAKS["hbm_gbs"]), flush=True)
        med, _ = timeit(lambda: ops.MaxPool2x2(x))
        nbytes = x.numel() * 2 * 1.25
        print("  maxpool2x2 %dx%dx128: med %.1f us  %.0f GB/s" % (H // 2, W // 2, med * 1e3, nbytes / med / 1e6), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["pool", "poolbwd", "mil", "sgd"]
    for w in which:
        globals()["bench_" + w]()
