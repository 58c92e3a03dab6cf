#!/usr/bin/env python
"""Headline benchmark: RoIs/sec through the NA-fWebSOD per-proposal head, fwd + bwd (+ gradient
all-reduce + SGD update), BASELINE.json config 2 per GPU: 2 images x 2000 proposals, 20 classes,
VGG16 conv5 map 512x38x50, two-stack noise-aware head, bf16 tensor-core path.

    python bench.py --gpus N --steps K --warmup W                 # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic input.  `value` is measured with the inputs resident in HBM; `e2e` runs the same step
through the public API from pinned HOST buffers (H2D of the step's inputs and D2H of its loss
inside the timed region).  Timing: CUDA events on the launching stream, barrier + synchronize
on both sides, max over ranks.  The working set (>1.5 GB of weights, gradients and activations
per step) is far larger than the 126 MB L2, so no explicit flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


if "reference" in sys.argv[1:]:
    # The CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would throttle the BLAS / OpenMP legs nine-fold -- override it BEFORE NumPy (and its BLAS) is loaded.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(_cpu_threads())

import numpy as np  # noqa: E402

METRIC = "RoIs/sec (fwd+bwd, WSDDN head)"
UNIT = "RoIs/s"
IMAGES_PER_GPU, ROIS_PER_IMAGE, NUM_CLASSES = 2, 2000, 21
C5, H5, W5 = 512, 38, 50
CPU_SAMPLE_ROIS = 1000        # cpu_baseline leg of the GPU arm: one image x 1000 of its 2000 RoIs per step (~1 s on 16 cores)
REF_BUDGET_S = 420.0          # reference arm: wall-clock budget of the timed steps (a slow box runs fewer steps and says so)


_JSON_OUT = None


def _protect_stdout():
    """Keep stdout for the ONE JSON line: libraries that print to fd 1 (NCCL's version banner under torchrun) are
    sent to stderr instead."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def _workload(noise=True):
    """The workload both arms are quoted on (BASELINE.json configs[1], per GPU)."""
    return ("NA-fWebSOD head fwd+bwd+gradient exchange+SGD, BASELINE config 2 per GPU: %d images x %d RoIs, %d classes, conv5 %dx%dx%d, "
            "RoIPoolF 7x7 @1/16 + boost, %s fc6/fc7 4096, seeded dropout" % (
                IMAGES_PER_GPU, ROIS_PER_IMAGE, NUM_CLASSES - 1, C5, H5, W5, "two-stack (clean + noisy)" if noise else "single-stack"))


def _flops_per_roi(noise=True, C=NUM_CLASSES - 1, D=C5 * 49, H=4096):
    """fwd 2*(D*H + H*H + 2*H*C), bwd as the reference runs it (no fc6 dX)."""
    fwd = 2 * (D * H + H * H + 2 * H * C)
    bwd = 2 * (D * H) + 2 * 2 * (H * H) + 2 * 2 * (2 * H * C)
    return (fwd + bwd) * (2 if noise else 1)


def synth_conv5(n, c, h, w, seed):
    """Post-ReLU-like conv5_3 map: U[0,1) * Bernoulli(0.5)."""
    rng = np.random.default_rng(seed)
    return (rng.random((n, c, h, w), dtype=np.float32) * (rng.random((n, c, h, w)) < 0.5)).astype(np.float32)


def synth_rois(r, img_h, img_w, batch_idx, seed):
    """MCG-like integer boxes, side 16 px .. image/2, as (batch_idx, x1, y1, x2, y2)."""
    rng = np.random.default_rng(seed)
    x1 = np.floor(rng.random(r) * (img_w - 17))
    y1 = np.floor(rng.random(r) * (img_h - 17))
    bw = np.floor(16 + rng.random(r) * (img_w / 2 - 16))
    bh = np.floor(16 + rng.random(r) * (img_h / 2 - 16))
    return np.stack([np.full(r, batch_idx), x1, y1, np.minimum(x1 + bw, img_w - 1), np.minimum(y1 + bh, img_h - 1)],
                    axis=1).astype(np.float32)


def synth_inputs(images, rois_per_image, seed=0):
    X = synth_conv5(images, C5, H5, W5, seed=seed)
    rois = np.concatenate([synth_rois(rois_per_image, H5 * 16, W5 * 16, b, seed=seed + 1 + b) for b in range(images)])
    rng = np.random.default_rng(seed + 100)
    obn = (rng.random(rois.shape[0]) + 1).astype(np.float32)
    L = np.zeros((images, NUM_CLASSES - 1), np.float32)
    for b in range(images):
        L[b, rng.integers(NUM_CLASSES - 1)] = 1          # webly images are single-label (loader_wsl.py:86-93)
    offs = np.asarray([b * rois_per_image for b in range(images)] + [rois.shape[0]], np.int32)
    return X, rois, obn, L, offs


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference's algorithm, timed on this box's host cores
# ------------------------------------------------------------------------------------------------
def cpu_head_step(X, rois, obn, L, params, masks):
    """One fwd+bwd of the head for ONE image with the CPU oracle (C restatement of RoIPoolF with
    OpenMP + NumPy/BLAS for the FC stack + the MIL / noise-weight / loss restatement)."""
    from oracle import nawsod_oracle as O
    from oracle import c_oracle as CO
    Y, _ = CO.roi_pool_f(X, rois, 1.0 / 16)
    feat = O.roi_feature_boost(Y, obn).reshape(Y.shape[0], -1)
    stacks, logits = [], []
    for pre in ("", "noisy_"):
        a = O.fc_stack_forward(feat, params[pre + "fc6_w"], params[pre + "fc6_b"], params[pre + "fc7_w"], params[pre + "fc7_b"],
                               masks[pre + "drop6"], masks[pre + "drop7"])
        stacks.append(a)
        logits += [O.fc(a["drop7"], params[pre + "fc8c_w"], params[pre + "fc8c_b"]), O.fc(a["drop7"], params[pre + "fc8d_w"], params[pre + "fc8d_b"])]
    out = O.mil_head_forward_backward(logits[0], logits[1], rois, L, logits[2], logits[3])
    for pre, a, dc, dd in (("", stacks[0], out["d_fc8c"], out["d_fc8d"]), ("noisy_", stacks[1], out["d_nfc8c"], out["d_nfc8d"])):
        _, _, dxc = O.fc_grad(a["drop7"], params[pre + "fc8c_w"], dc)
        _, _, dxd = O.fc_grad(a["drop7"], params[pre + "fc8d_w"], dd)
        O.fc_stack_backward(feat, a, params[pre + "fc6_w"], params[pre + "fc7_w"], dxc + dxd, masks[pre + "drop6"], masks[pre + "drop7"])
    return float(out["loss_cls"]) + float(out["loss_cls_noise"])


def cpu_problem(rois_n, seed=0, params=None):
    from oracle import nawsod_oracle as O
    X, rois, obn, L, _ = synth_inputs(1, rois_n, seed)
    if params is None:
        params = O.synth_params(NUM_CLASSES - 1, C5 * 49, 4096, noise=True, seed=2)
    rng = np.random.default_rng(3)
    masks = {k: (rng.random((rois_n, 4096)) < 0.5).astype(np.float32) for k in ("drop6", "drop7", "noisy_drop6", "noisy_drop7")}
    return X, rois, obn, L, params, masks


def run_cpu(steps, warmup, sample_rois, images=1, budget_s=None):
    """`steps` timed passes of `images` independent one-image problems of `sample_rois` RoIs each (the reference runs one
    image per net, wsl_heads.py:214).  Returns (RoIs/s, per-step seconds list, steps actually timed)."""
    probs = []
    for b in range(images):                                  # the images share one set of parameters
        probs.append(cpu_problem(sample_rois, seed=10 * b, params=probs[0][4] if probs else None))
    for _ in range(warmup):
        for prob in probs:
            cpu_head_step(*prob)
    per = []
    for i in range(steps):
        t0 = time.perf_counter()
        for prob in probs:
            cpu_head_step(*prob)
        per.append(time.perf_counter() - t0)
        if budget_s is not None and sum(per) + per[-1] > budget_s and i + 1 < steps:
            break
    return images * sample_rois * len(per) / sum(per), per, len(per)


def cpu_forward_only(rois_n, repeats=5):
    """BASELINE config 1: head FORWARD on the CPU, 1 image x `rois_n` RoIs, clean stack only, dropout off (test net)."""
    from oracle import nawsod_oracle as O
    from oracle import c_oracle as CO
    X, rois, obn, L, params, _ = cpu_problem(rois_n)

    def fwd():
        Y, _ = CO.roi_pool_f(X, rois, 1.0 / 16)
        feat = O.roi_feature_boost(Y, obn).reshape(Y.shape[0], -1)
        a = O.fc_stack_forward(feat, params["fc6_w"], params["fc6_b"], params["fc7_w"], params["fc7_b"], None, None)
        c, d = O.fc(a["drop7"], params["fc8c_w"], params["fc8c_b"]), O.fc(a["drop7"], params["fc8d_w"], params["fc8d_b"])
        return O.test_cls_prob(O.wsl_outputs(c, d)[2])
    fwd()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fwd()
        ts.append(time.perf_counter() - t0)
    return rois_n / float(np.median(ts)), float(np.median(ts))


def reference_arm(args):
    """The reference's CPU path on this box's host cores: the SAME workload per step as the GPU arm (BASELINE config 2:
    2 images x 2000 RoIs, two-stack head, fwd+bwd, fp32), --steps / --warmup honoured, all host threads.  Caffe2 cannot be
    installed here, so the operators are the oracle port: the reference's own .cc operators where they
    compile (oracle/_ref), C/OpenMP RoIPoolF, NumPy/BLAS for the Caffe2 built-ins."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = _cpu_threads()
    rpi = args.ref_rois_per_image
    value, per, done = run_cpu(args.steps, args.warmup, rpi, images=IMAGES_PER_GPU, budget_s=REF_BUDGET_S)
    c1_value, c1_s = cpu_forward_only(rpi)
    desc = ("%d images x %d RoIs per step = the GPU arm's per-GPU workload (not a sub-sample), 512x38x50 map, 20 classes, two-stack head, "
            "fwd+bwd, fp32, %d timed steps%s; oracle port of the reference's CPU algorithm (Caffe2 itself is not installable here): "
            "C/OpenMP RoIPoolF + NumPy/BLAS, %d threads") % (
                IMAGES_PER_GPU, rpi, done, "" if done == args.steps else " (of %d requested: %.0f s budget)" % (args.steps, REF_BUDGET_S), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": float(np.mean(per)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": _workload(True), "global_rois_per_step": IMAGES_PER_GPU * rpi,
                   "parallelism": "host cores of rank 0 (the reference's CPU path does not shard; at N > 1 the other ranks exit)",
                   "sample": desc, "threads": cores,
                   "step_s": {"min": float(np.min(per)), "median": float(np.median(per)), "max": float(np.max(per))}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # BASELINE.json configs[0]: the reference's own CPU-runnable case (forward only, 1 image x 2000 RoIs, clean stack)
        "config1_cpu_forward": {"value": c1_value, "unit": "RoIs/s (fwd only)", "s_per_image": c1_s, "cores": cores,
                                "sample": "1 image x %d RoIs, median of 5" % rpi},
    }
    _emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, "/tmp/nawsod_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_start=None, t_end=None, legs=None):
        """Samples taken between the wall-clock marks t_start / t_end (the timed regions) are the ones reported;
        nvidia-smi itself is started earlier (its start-up takes longer than a short timed region).
        legs: {name: (t0, t1)} adds the median SM clock of each timed leg (a short first leg may still run at
        boost clocks while a later one is already power-capped)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        import datetime
        rows, smax = [], None
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                row = [float(f[1]), float(f[3]), f[5:9], None]
                smax = float(f[2])
            except ValueError:
                continue
            if len(f) > 9:
                try:
                    row[3] = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    pass
            rows.append(row)
        inside = [r for r in rows if r[3] is not None and t_start is not None and t_start - 0.05 <= r[3] <= t_end + 0.05]
        window = "timed regions" if inside else "whole run (no sample fell inside the timed regions)"
        sm, power, reasons = [], [], set()
        for c, p, flags, _ in (inside or rows):
            sm.append(c); power.append(p)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
               "power_w_max": max(power), "window": window}
        for name, (a, b) in (legs or {}).items():
            leg = [r[0] for r in rows if r[3] is not None and a <= r[3] <= b]
            out["sm_mhz_" + name] = float(np.median(leg)) if leg else None
        return out


def assemble_line(*, steps, warmup, world, R, S, bf16, noise, ms_total, ms_e2e, h2d_bytes, d2h_bytes, launches, clocks, kernel_ms,
                  n_panels, iso, cpu, loss, dp_info, timing=None, tf32=None):
    """The ONE JSON line of the GPU arm from the run's raw measurements (pure: no CUDA, no torch -- tests/test_bench_contract.py
    runs it on the CPU).  ms_total / ms_e2e: device time of the `steps` timed steps of the resident / end-to-end leg (max over
    ranks; the median block when the region was repeated, see `timing`); kernel_ms: mean CUDA-event duration per launch of
    fc6_fwd, fc6_bwd_w (one row panel), roi_pool_f, mil_head inside the timed steps; iso: isolated RoIPoolF timings (rank 0)
    or {}; dp_info: sync / fc6_panels / p2p_selftest / engine; timing: {"value": ..., "e2e": ...} block / per-step spread;
    tf32: {"ms_total", "steps", "kernel_ms", "n_panels"} of the fp32 (TF32 tensor path) run of the same workload, or None."""
    peaks = _peaks()
    ms_step = ms_total / steps
    value = world * R * steps / (ms_total * 1e-3)
    e2e_value = world * R * steps / (ms_e2e * 1e-3)
    es = 2 if bf16 else 4
    fc6_flops = 2.0 * R * (S * 4096) * (C5 * 49)
    t_fwd, t_bww, t_pool, t_mil = (kernel_ms.get(k) for k in ("fc6_fwd", "fc6_bwd_w", "roi_pool_f", "mil_head"))
    t_bww_total = t_bww * n_panels if t_bww else None
    pool_bytes = R * (C5 * 49 * es + 20) + IMAGES_PER_GPU * C5 * H5 * W5 * es     # no argmax: conv body frozen (StopGradient)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        # per LAUNCH, like `achieved`: the capture of a row-panel launch when the step runs the captured panel count, the
        # capture of the whole-matrix launch when it runs unpanelled, else unknown
        if n_panels == tj.get("fc6_bwd_w_panels"):
            traffic = tj.get("fc6_bwd_w_panel_dram_bytes_per_launch")
        elif n_panels == 1:
            traffic = tj.get("fc6_bwd_w_dram_bytes_per_launch")
    tensor_peak = peaks["tf_sustained"] * (1.0 if bf16 else 0.5)
    roofline = {
        "kernel": "gemm_tcgen05_kernel<256,MN,MN,pair> (fc6 weight gradient, dY^T.X, both stacks in one GEMM; CTA pairs, "
                  "tcgen05 cta_group::2, TMA-store epilogue)",
        "bound": "tensor", "achieved": fc6_flops / (t_bww_total * 1e-3) / 1e12 if t_bww_total else None, "peak": tensor_peak,
        "unit": "TFLOP/s", "traffic": traffic, "traffic_source": "ncu --set full capture of one such launch (profiles/ncu_traffic.json)",
        "launches_per_step": n_panels,
        # a panel launch reads its columns of dY and all pooled features and writes its rows of dW (fp32)
        "algorithmic_bytes_per_launch": (R * (S * 4096) * es + R * (C5 * 49) * es * n_panels + (S * 4096) * (C5 * 49) * 4) / n_panels,
        "peak_source": peaks["source"] + ("; sustained bf16" if bf16 else "; TF32 = bf16/2"),
    }
    roofline["frac"] = roofline["achieved"] / tensor_peak if roofline["achieved"] else None
    kernels = {
        "fc6_fwd": {"ms": t_fwd, "tflops": fc6_flops / (t_fwd * 1e-3) / 1e12 if t_fwd else None,
                    "frac_tensor": fc6_flops / (t_fwd * 1e-3) / 1e12 / tensor_peak if t_fwd else None},
        "fc6_bwd_w": {"ms": t_bww_total, "panels": n_panels, "tflops": roofline["achieved"], "frac_tensor": roofline["frac"]},
        "roi_pool_f": {"ms": t_pool, "gbs": pool_bytes / (t_pool * 1e-3) / 1e9 if t_pool else None,
                       "frac_hbm": pool_bytes / (t_pool * 1e-3) / 1e9 / peaks["hbm"] if t_pool else None,
                       "algorithmic_bytes": pool_bytes},
        "mil_head": {"ms": t_mil},
        "step_tensor_frac": _flops_per_roi(noise) * R / (ms_step * 1e-3) / 1e12 / tensor_peak,
    }
    if tf32:
        # the reference's precision (fp32 end to end: Caffe2 FC = sgemm) on the TF32 tensor path; peak = bf16 / 2
        t32_peak = peaks["tf_sustained"] * 0.5
        ms32 = tf32["ms_total"] / tf32["steps"]
        f32, w32 = tf32["kernel_ms"].get("fc6_fwd"), tf32["kernel_ms"].get("fc6_bwd_w")
        w32 = w32 * tf32["n_panels"] if w32 else None
        x3 = tf32.get("fp32_three_pass")
        if x3:
            ms3 = x3["ms_total"] / x3["steps"]
            kernels["fp32_step"] = {
                "note": "same workload at the reference's sgemm accuracy: fp32 storage, every GEMM operand a TF32 (high, low) pair, "
                        "three kind::tf32 passes per product (rel <= 1e-4 vs the fp32 oracle at this size), resident inputs",
                "ms_per_step": ms3, "rois_per_s": R * x3["steps"] / (x3["ms_total"] * 1e-3), "steps": x3["steps"]}
        w1 = tf32.get("wsddn_bf16")
        if w1:
            ms1 = w1["ms_total"] / w1["steps"]
            kernels["wsddn_step"] = {
                "note": "single-stack head (plain WSDDN: clean stack only, unweighted loss), same inputs, bf16, resident inputs",
                "ms_per_step": ms1, "rois_per_s": R * w1["steps"] / (w1["ms_total"] * 1e-3), "steps": w1["steps"],
                "step_tensor_frac": _flops_per_roi(False) * R / (ms1 * 1e-3) / 1e12 / tensor_peak}
        kernels["tf32_step"] = {
            "note": "same workload, fp32 storage + kind::tf32 GEMMs (operands pre-rounded to nearest TF32), resident inputs",
            "ms_per_step": ms32, "rois_per_s": R * tf32["steps"] / (tf32["ms_total"] * 1e-3), "steps": tf32["steps"],
            "peak_tflops": t32_peak,
            "fc6_fwd": {"ms": f32, "tflops": fc6_flops / (f32 * 1e-3) / 1e12 if f32 else None,
                        "frac_tensor": fc6_flops / (f32 * 1e-3) / 1e12 / t32_peak if f32 else None},
            "fc6_bwd_w": {"ms": w32, "panels": tf32["n_panels"], "tflops": fc6_flops / (w32 * 1e-3) / 1e12 if w32 else None,
                          "frac_tensor": fc6_flops / (w32 * 1e-3) / 1e12 / t32_peak if w32 else None},
            "step_tensor_frac": _flops_per_roi(noise) * R / (ms32 * 1e-3) / 1e12 / t32_peak,
        }
    if iso:
        # algorithmic bytes: Y + (argmax when the conv body trains) + rois + the map read once
        b_step = pool_bytes
        b_f32 = R * (C5 * 49 * 4 * 2 + 20) + IMAGES_PER_GPU * C5 * H5 * W5 * 4
        kernels["roi_pool_f_isolated"] = {
            "note": "RoIPoolF alone, back-to-back launches (burst HBM peak applies); in the step it shares HBM with the pipelined SGD",
            "step_config": {"ms": iso["step_config"], "gbs": b_step / (iso["step_config"] * 1e-3) / 1e9,
                            "frac_hbm": b_step / (iso["step_config"] * 1e-3) / 1e9 / peaks["hbm"], "algorithmic_bytes": b_step},
            "fp32_train_argmax": {"ms": iso["fp32_train"], "gbs": b_f32 / (iso["fp32_train"] * 1e-3) / 1e9,
                                  "frac_hbm": b_f32 / (iso["fp32_train"] * 1e-3) / 1e9 / peaks["hbm"], "algorithmic_bytes": b_f32},
        }
    exchange = {"sharded": "NCCL reduce-scatter fp32 grads + sharded SGD + all-gather bf16 operands",
                "p2p": "peer-mapped (CUDA IPC over NVSwitch) scatter of fp32 grads into the owner's staging + fused reduce/SGD on the owner "
                       "+ scatter of the bf16 operands back, ordered by flag kernels",
                "allreduce": "NCCL all-reduce fp32 grads + full SGD"}
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if bf16 else "tf32", "data": "synthetic",
        "config": {"workload": _workload(noise),
                   "global_rois_per_step": world * R, "parallelism": "dp%d (images sharded by rank; gradient exchange per step: %s)" % (
                       world, "none" if world == 1 else exchange.get(dp_info["sync"], dp_info["sync"])),
                   "l2": "working set per step (weights 0.5 GB bf16 + 0.96 GB fp32 grads + activations) >> 126 MB L2; no flush needed",
                   "fc6_panels": dp_info["fc6_panels"],
                   "p2p_selftest": dp_info["p2p_selftest"],
                   "p2p_engine": dp_info.get("engine"),
                   "fc6_update": "stand-alone SGD kernel per row panel on a side stream" if world == 1 else "on the owner rank of each slice"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": ms_e2e / steps},
        "timing": timing,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "cpu_baseline": cpu,
        "loss": loss,
    }


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import nafwebsod_b200 as pkg
    from nafwebsod_b200 import _lib
    from nafwebsod_b200.heads import WeblyHeadModel
    from nafwebsod_b200.dp import DataParallelHead

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: libnawsod has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if args.comm_sms > 0:
            # the persistent GEMMs leave --comm-sms SMs to the NCCL kernels while an exchange is in flight; cap NCCL to match
            os.environ["NAWSOD_COMM_SMS"] = str(args.comm_sms)
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.comm_sms))
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    for kv in filter(None, os.environ.get("NAWSOD_TUNING", "").split(",")):      # e.g. NAWSOD_TUNING=sgd_max_ctas=148
        k, v = kv.split("=")
        pkg.set_tuning(k, int(v))
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    noise = args.head == "na"

    # ---- model: random-init weights of the reference architecture (no checkpoints offline) ----
    model = WeblyHeadModel(NUM_CLASSES, C5, 7, 4096, noise=noise, dtype=dtype, device=dev)

    def init_parameters():
        g = torch.Generator(device=dev).manual_seed(2)
        nw = model.n_weights
        model.flat_param.zero_()
        model.flat_mom.zero_()
        model.iter_count = 0
        model.flat_param[:nw].normal_(0.0, 0.01, generator=g)          # gauss_fill(0.01); biases stay 0
        for s in range(model.S):                                        # XavierFill for fc8
            lim = float(np.sqrt(3.0 / 4096))
            model.p["W8_%d" % s].uniform_(-lim, lim, generator=g)
        model.sync_shadow()

    init_parameters()
    os.environ.setdefault("NAWSOD_P2P_TIMEOUT_MS", "10000")    # peer-exchange watchdog: a 5 ms step never waits this long
    dp = DataParallelHead(model, fc6_panels=args.fc6_panels, sync=args.dp_sync)
    dp.broadcast_parameters()
    model.UpdateWorkspaceLr(1e-3)

    def exchange_timed_out():
        """Peer exchange only: did a wait kernel's watchdog fire on any rank?  (Synchronises; same answer on every rank.)"""
        if not hasattr(dp.exchange, "status"):
            return False
        dp.flush()
        torch.cuda.synchronize()
        bad = dp.exchange.status.to(torch.int32).clone()
        if world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        return int(bad.item()) != 0

    # ---- this rank's shard of the synthetic batch (weak scaling: 2 images per GPU) ----
    X, rois, obn, L, offs = synth_inputs(IMAGES_PER_GPU, ROIS_PER_IMAGE, seed=1000 * rank)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hX, hrois, hobn, hL, hoffs = pin(X), pin(rois), pin(obn), pin(L), pin(offs)
    h2d_bytes = sum(t.numel() * t.element_size() for t in (hX, hrois, hobn, hL, hoffs))
    R = rois.shape[0]

    def feed_from_host():
        model.FeedBlobs(hX.to(dev, non_blocking=True), hrois.to(dev, non_blocking=True), hobn.to(dev, non_blocking=True),
                        hL.to(dev, non_blocking=True), hoffs.to(dev, non_blocking=True), x_layout="NCHW")

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = []                        # per timed block: host milliseconds spent enqueueing one step

    def timed(fn, steps, tail=None, step_events=None):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        h0 = time.perf_counter()
        for i in range(steps):
            fn(i)
            if step_events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                step_events.append(e)
        host_ms.append((time.perf_counter() - h0) * 1e3 / max(steps, 1))      # host time to ENQUEUE a step (no sync inside)
        dp.flush()                      # the last step's parameter exchange belongs to the timed region
        b.record()
        if tail is not None:
            tail()                      # e.g. the last device->host reads (enqueued before b was recorded)
        sync_all()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if step_events is not None:
            step_events.insert(0, a)
        return ms.item()

    def timed_blocks(run_block, steps):
        """EXACTLY `steps` steps per timed block (barrier + synchronize on both sides, max over ranks).  A block of a few ms
        per step is over in < 0.1 s, where one host hiccup dominates: the block is repeated until >= 1 s has been timed
        (at most 15 blocks, the count agreed across ranks through the max-reduced first block) and the MEDIAN block is
        reported, with the spread of the blocks and of the individual steps beside it."""
        blocks, per_step = [], []
        h_first = len(host_ms)

        def one():
            ev = []
            ms, extra = run_block(steps, ev)
            blocks.append(ms)
            per_step.extend(ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1))
            return extra
        extra = one()
        n = int(min(15, max(1, np.ceil(1000.0 / max(blocks[0], 1e-3)))))
        for _ in range(n - 1):
            extra = one()
        order = sorted(range(len(blocks)), key=lambda i: blocks[i])
        med = blocks[order[len(order) // 2]]
        stats = {"blocks": len(blocks), "steps_per_block": steps,
                 "block_ms_per_step": {"min": min(blocks) / steps, "median": med / steps, "max": max(blocks) / steps},
                 "step_ms_this_rank": {"min": float(np.min(per_step)), "median": float(np.median(per_step)), "max": float(np.max(per_step))},
                 # when this approaches ms_per_step the run is launch-bound: the GPU waits for the host
                 "host_enqueue_ms_per_step_this_rank": float(np.median(host_ms[h_first:]))}
        return med, stats, extra

    # ---- resident-input leg (value) ----
    feed_from_host()                      # inputs now live in HBM (channels-last bf16 map etc.)
    step_resident = lambda i: dp.step(dropout_seed=i + 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()               # before the warm-up: nvidia-smi needs a moment to deliver its first sample
    sync_all()                        # every rank enters the first step together (the peer exchange's watchdog counts from there)
    for i in range(args.warmup):
        step_resident(i)
        if i == 0 and args.dp_sync == "auto" and exchange_timed_out():
            # sync=auto chose the peer-mapped exchange (its dry run passed) but a peer did not deliver inside the first real
            # step: restart the run on the NCCL schedule instead of timing a broken exchange (the line records it)
            if rank == 0:
                sys.stderr.write("bench: p2p watchdog fired in the first step; falling back to --dp-sync sharded\n")
            why = "%s; watchdog fired in the first training step -> NCCL sharded" % dp.p2p_selftest
            init_parameters()
            dp = DataParallelHead(model, fc6_panels=args.fc6_panels, sync="sharded")
            dp.p2p_selftest = why
            dp.broadcast_parameters()
            step_resident(0)
    t_mark0 = time.time()
    model.profile = {}
    launches0 = _lib.launch_count
    ms_total, value_stats, _ = timed_blocks(lambda k, ev: (timed(step_resident, k, step_events=ev), None), args.steps)
    t_mark1 = time.time()
    launches = (_lib.launch_count - launches0) // value_stats["blocks"]          # per block of `steps` steps
    prof, model.profile = model.profile, None

    # ---- end-to-end leg: host buffers in, loss out, every step ----
    # The public feed path is the device-side blobs queue of nafwebsod_b200/loader.py (the reference's RoIDataLoader /
    # BlobsQueue, loader_wsl.py:215-238): step i's inputs are copied from pinned host memory on a copy stream while
    # step i-1 computes, and its loss is copied back to pinned host memory behind the step; the host blocks on the
    # loss of the previous step only.  Every step's H2D and D2H copies are issued and completed inside the timed region.
    from nafwebsod_b200.loader import BlobsQueue, LossFetcher

    # one queue / fetcher for the whole leg: their device and pinned buffers are rings allocated on first use, so the timed
    # blocks run allocation-free (the warm-up block below performs the allocations)
    queue, fetch_ring = BlobsQueue(model, capacity=2, x_layout="NCHW"), LossFetcher(lag=1)

    def run_e2e(steps, ev=None, seed0=0):
        fetch = fetch_ring
        fetch.values = []
        bytes0 = queue.h2d_bytes

        def step(i):
            if i == 0:
                queue.enqueue_blobs(hX, hrois, hobn, hL, hoffs)
            queue.dequeue_blobs()
            if i + 1 < steps:
                queue.enqueue_blobs(hX, hrois, hobn, hL, hoffs)      # prefetch: overlaps this step's kernels
            bl = dp.step(dropout_seed=seed0 + i + 1)
            fetch.push(bl["loss"])                                    # D2H read of the step's result
        ms = timed(step, steps, tail=fetch.wait_all, step_events=ev)
        assert queue.h2d_bytes - bytes0 == steps * h2d_bytes
        return ms, fetch

    run_e2e(max(2, args.warmup // 2))
    t_mark2 = time.time()
    ms_e2e, e2e_stats, fetch = timed_blocks(run_e2e, args.steps)
    losses = fetch.values
    # clocks are sampled across BOTH timed regions (resident + end-to-end) so that short runs still
    # collect samples under load
    t_mark3 = time.time()
    clocks = sampler.stop(t_mark0, t_mark3, legs={"value_leg": (t_mark0, t_mark1), "e2e_leg": (t_mark2, t_mark3)}) if rank == 0 else None
    if hasattr(dp.exchange, "check"):
        dp.exchange.check()           # peer exchange: a watchdog time-out in any wait kernel invalidates the run -- fail loudly
    d2h_bytes = int(losses[-1].numel() * losses[-1].element_size())
    assert all(bool(torch.isfinite(l).all()) for l in losses), "non-finite loss in the benchmark"

    if os.environ.get("NAWSOD_P2P_PROFILE") and getattr(dp.exchange, "profile", 0) is None:
        # one instrumented step: when did each bucket become ready / leave / arrive / get updated / get published
        dp.flush(); sync_all()
        dp.exchange.profile = []
        base = torch.cuda.Event(enable_timing=True); base.record()
        dp.step(dropout_seed=99); dp.flush()
        end = torch.cuda.Event(enable_timing=True); end.record()
        torch.cuda.synchronize()
        if rank == 0:
            tl = sorted((base.elapsed_time(e), lab, b) for lab, b, e in dp.exchange.profile)
            sys.stderr.write("p2p timeline (ms from step start; step %.3f ms): %s\n" % (
                base.elapsed_time(end), "  ".join("%.2f:%s[%d]" % t for t in tl)))
        dp.exchange.profile = None

    # ---- isolated legs (rank 0, after the timed regions): RoIPoolF launched back to back with nothing else on the GPU.
    # In the step the pool overlaps the tail of the previous step's pipelined SGD (HBM-bound), so its in-step duration
    # understates the kernel; each launch writes 196-401 MB (> 126 MB L2), so back-to-back launches stay HBM-to-HBM.
    iso = {}
    if rank == 0 and not args.no_isolated:
        from nafwebsod_b200 import ops
        dp.flush(); torch.cuda.synchronize()

        def isolated_ms(fn, iters=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters
        bl = model.blobs
        iso["step_config"] = isolated_ms(lambda: ops.RoIPoolF(
            bl["conv5"], bl["rois"], spatial_scale=1.0 / 16, is_test=True, boost=bl["obn_scores"], x_layout="NHWC", y_layout="NHWC",
            out_dtype=dtype))
        x32 = ops.to_channels_last(hX.to(dev), torch.float32)
        iso["fp32_train"] = isolated_ms(lambda: ops.RoIPoolF(
            x32, bl["rois"], spatial_scale=1.0 / 16, is_test=False, boost=bl["obn_scores"], x_layout="NHWC", y_layout="NHWC",
            out_dtype=torch.float32))
        del x32

    # ---- the reference's precision: fp32 storage on the TF32 tensor path, same workload (one GPU only; SURVEY.md 8d config 2
    # asks for both) -- reported inside the one line as kernels.tf32_step
    tf32 = None
    if world == 1 and dtype == torch.bfloat16 and not args.no_tf32:
        dp.flush(); torch.cuda.synchronize()
        main_model, main_dp = model, dp
        legs = {}
        # "tf32": one tensor-core pass per product; "fp32": operands as TF32 (high, low) pairs, three passes per product
        # (the reference's sgemm accuracy, tests/test_gpu_head.py::test_head_full_size_fp32_config2)
        # "wsddn": the single-stack head (plain WSDDN: no noisy stack, no noise-aware weights) on the bench's own bf16 path --
        # SURVEY.md 8d config 2 asks for both variants
        variants = [("tf32", torch.float32, "tf32", noise, args.steps), ("fp32", torch.float32, "fp32", noise, min(args.steps, 10))]
        if noise:
            variants.append(("wsddn", torch.bfloat16, None, False, args.steps))
        for precision, leg_dtype, leg_precision, leg_noise, nsteps in variants:
            model = WeblyHeadModel(NUM_CLASSES, C5, 7, 4096, noise=leg_noise, dtype=leg_dtype, device=dev, precision=leg_precision)
            init_parameters()
            dp = DataParallelHead(model, fc6_panels=args.fc6_panels, sync=args.dp_sync)
            model.UpdateWorkspaceLr(1e-3)
            feed_from_host()
            for i in range(3):
                step_resident(i)
            model.profile = {}
            ms32 = timed(step_resident, nsteps)
            prof32, model.profile = model.profile, None
            mean = lambda ev: sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1) if ev else None
            legs[precision] = {"ms_total": ms32, "steps": nsteps,
                               "kernel_ms": {k: mean(prof32.get(k, [])) for k in ("fc6_fwd", "fc6_bwd_w")},
                               "n_panels": max(1, len(prof32.get("fc6_bwd_w", [])) // max(nsteps, 1))}
            dp.flush(); torch.cuda.synchronize()
        tf32 = legs["tf32"]
        tf32["fp32_three_pass"] = legs["fp32"]
        tf32["wsddn_bf16"] = legs.get("wsddn")
        model, dp = main_model, main_dp

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel + per-kernel breakdown from the in-run events ----
    def avg_ms(name):
        ev = prof.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1) if ev else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = _cpu_threads()
        sample = CPU_SAMPLE_ROIS
        v, per, _ = run_cpu(5, 1, sample)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "5 steps of 1 image x %d RoIs (bounded sample of the 2 x 2000 workload), fp32 oracle port: C/OpenMP RoIPoolF + NumPy/BLAS "
                         "FC stack + MIL/loss restatement, %.2f s per step" % (sample, float(np.median(per)))}
    line = assemble_line(
        steps=args.steps, warmup=args.warmup, world=world, R=R, S=model.S, bf16=dtype == torch.bfloat16, noise=noise,
        ms_total=ms_total, ms_e2e=ms_e2e, h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes, launches=launches, clocks=clocks,
        kernel_ms={k: avg_ms(k) for k in ("fc6_fwd", "fc6_bwd_w", "roi_pool_f", "mil_head")},
        n_panels=max(1, len(prof.get("fc6_bwd_w", [])) // max(args.steps * value_stats["blocks"], 1)), iso=iso, cpu=cpu,
        loss=[float(x) for x in losses[-1].flatten().tolist()], timing={"value": value_stats, "e2e": e2e_stats}, tf32=tf32,
        dp_info={"sync": dp.sync, "fc6_panels": dp.fc6_panels, "p2p_selftest": dp.p2p_selftest,
                 "engine": ("%s/%s" % (getattr(dp.exchange, "rs_mode", "-"), dp.exchange.engine)) if hasattr(dp.exchange, "engine") else None})
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# BASELINE configs 3 and 5 (one GPU; `--config 3` / `--config 5`): not the driver's headline line, but the same harness,
# one JSON line each, so that the numbers are reproducible from a tracked command (profiles/ keeps the outputs)
# ------------------------------------------------------------------------------------------------
def _event_ms(fn, iters, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def _mixed_rois(r, img_h, img_w, seed):
    """Config 3's RoI mixture: MCG-like boxes from 16 px to the full image (window areas 1 ... ~10^4 cells)."""
    rng = np.random.default_rng(seed)
    x1 = np.floor(rng.random(r) * (img_w - 17))
    y1 = np.floor(rng.random(r) * (img_h - 17))
    side = np.exp(rng.uniform(np.log(16.0), np.log(float(max(img_h, img_w))), r))       # log-uniform sizes
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), r))
    bw, bh = np.floor(side * np.sqrt(ar)), np.floor(side / np.sqrt(ar))
    return np.stack([np.zeros(r), x1, y1, np.minimum(x1 + bw, img_w - 1), np.minimum(y1 + bh, img_h - 1)], axis=1).astype(np.float32)


def gpu_config3(args):
    """flickr_coco shape (configs/flickr_coco/na_wsddn_V-16-C5_1x.yaml): 80 classes, 4000 proposals per image, one image per
    step (the reference's 1 image / GPU), conv5 maps of the multi-scale training (short side 480 ... 1200, max side 2000)
    at 1/16 and, for the largest, at 1/8 (WSL.DILATION 2), RoI sizes from 16 px to the full image; NA two-stack head,
    bf16, fwd + bwd + SGD, resident inputs."""
    import torch
    from nafwebsod_b200 import _lib
    from nafwebsod_b200.heads import WeblyHeadModel
    from nafwebsod_b200.dp import DataParallelHead
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peaks = _peaks()
    ncls, R = 81, 4000
    model = WeblyHeadModel(ncls, C5, 7, 4096, noise=True, dtype=torch.bfloat16, device=dev)
    g = torch.Generator(device=dev).manual_seed(2)
    model.flat_param[:model.n_weights].normal_(0.0, 0.01, generator=g)
    model.sync_shadow()
    model.UpdateWorkspaceLr(1e-3)
    dp = DataParallelHead(model, fc6_panels=args.fc6_panels)
    cases = []
    for (h, w, stride, what) in ((30, 40, 16, "short side 480"), (43, 57, 16, "short side 688"), (54, 72, 16, "short side 864"),
                                 (75, 125, 16, "short side 1200, max side 2000"), (150, 250, 8, "short side 1200, max side 2000 at 1/8 (WSL.DILATION 2)")):
        X = torch.from_numpy(synth_conv5(1, C5, h, w, seed=h)).to(dev)
        rois = torch.from_numpy(_mixed_rois(R, h * stride, w * stride, seed=w)).to(dev)
        obn = (torch.rand(R, device=dev) + 1)
        L = torch.zeros(1, ncls - 1, device=dev); L[0, 17] = 1
        model.spatial_scale = 1.0 / stride
        model.FeedBlobs(X, rois, obn, L, x_layout="NCHW")
        model.profile = {}
        med, best = _event_ms(lambda: dp.step(), args.steps, warmup=args.warmup)
        dp.flush(); torch.cuda.synchronize()
        prof, model.profile = model.profile, None
        mean = lambda k: float(np.mean([a.elapsed_time(b) for a, b in prof.get(k, [])])) if prof.get(k) else None
        pool_bytes = R * (C5 * 49 * 2 + 20) + C5 * h * w * 2
        t_pool = mean("roi_pool_f")
        cases.append({"map": "%dx%d @1/%d" % (h, w, stride), "what": what, "ms_per_step": med, "rois_per_s": R / (med * 1e-3),
                      "step_tensor_frac": _flops_per_roi(True, C=ncls - 1) * R / (med * 1e-3) / 1e12 / peaks["tf_sustained"],
                      "roi_pool_f": {"ms": t_pool, "gbs": pool_bytes / (t_pool * 1e-3) / 1e9 if t_pool else None,
                                     "frac_hbm": pool_bytes / (t_pool * 1e-3) / 1e9 / peaks["hbm"] if t_pool else None},
                      "mil_head_ms": mean("mil_head"), "fc6_fwd_ms": mean("fc6_fwd")})
    worst = min(c["rois_per_s"] for c in cases)
    _emit({"metric": METRIC, "value": worst, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": max(c["ms_per_step"] for c in cases), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "BASELINE config 3 (flickr_coco shape): 1 image x 4000 RoIs per step, 80 classes, NA two-stack head fwd+bwd+SGD, "
                                  "conv5 maps of the multi-scale schedule, RoIs 16 px ... full image; `value` = the slowest map",
                      "l2": "working set per step >> 126 MB L2"},
           "cases": cases})
    return 0


def gpu_config5(args):
    """Test-time-augmented inference (core/test_wsl.py:181-281): forward only, 5 scales x {orig, hflip} = 10 passes per
    image through test_time.im_detect_bbox_aug (projection, dedup, head forward, inverse scatter, averaging on the GPU),
    then threshold + NMS + limit; proposal-count sweep 500 ... 8000 (+ the shipped TEST.PROPOSAL_LIMIT 9999), bf16."""
    import torch
    from nafwebsod_b200 import _lib, ops, test_time
    from nafwebsod_b200.heads import WeblyHeadModel
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    m = WeblyHeadModel(NUM_CLASSES, C5, 7, 4096, dtype=torch.bfloat16, train=False, device=dev)
    g = torch.Generator(device=dev).manual_seed(2)
    m.flat_param[:m.n_weights].normal_(0.0, 0.01, generator=g)
    m.sync_shadow()
    img_h, img_w = 375, 500                                   # a VOC-sized image
    scales = [688, 480, 576, 864, 1200]                       # TEST.SCALE, then TEST.BBOX_AUG.SCALES
    maps = {}
    for sc_px in scales:
        sc = sc_px / float(min(img_h, img_w))
        h, w = int(np.ceil(img_h * sc / 16)), int(np.ceil(img_w * sc / 16))
        maps[sc_px] = (torch.from_numpy(synth_conv5(1, C5, h, w, seed=sc_px)).to(dev).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), sc)
    order = [(688, True)] + [(sc_px, f) for sc_px in scales[1:] for f in (False, True)] + [(688, False)]
    passes = [(maps[sc_px][0], maps[sc_px][1], img_w if f else None) for sc_px, f in order]
    cases = []
    for R in (500, 1000, 2000, 4000, 8000, 9999):
        boxes = torch.from_numpy(synth_rois(R, img_h, img_w, 0, seed=R)[:, 1:].copy()).to(dev)
        obn = torch.rand(R, device=dev)
        med, _ = _event_ms(lambda: test_time.im_detect_bbox_aug(m, passes, boxes, obn, sync=False), max(5, args.steps // 2))
        scores = test_time.im_detect_bbox_aug(m, passes, boxes, obn, sync=False)
        nms_ms, _ = _event_ms(lambda: ops.nms_and_limit(scores, boxes, score_thresh=1e-9, nms_thresh=0.5, detections_per_im=100), 10)
        cases.append({"rois": R, "passes": len(passes), "ms_per_image": med, "roi_passes_per_s": R * len(passes) / (med * 1e-3),
                      "images_per_s": 1e3 / (med + nms_ms), "nms_and_limit_ms": nms_ms})
    ref = [c for c in cases if c["rois"] == 2000][0]
    _emit({"metric": "RoI-passes/sec (forward only, 10-pass test-time augmentation)", "value": ref["roi_passes_per_s"], "unit": "RoI-passes/s",
           "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": ref["ms_per_image"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "BASELINE config 5: test-time-augmented inference, 5 scales x {orig, hflip} on a 375x500 image, 20 classes, "
                                  "DEDUP_BOXES 1/16, no host round trip per pass; `value` at 2000 RoIs per image; sweep in `cases`"},
           "cases": cases})
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--head", default="na", choices=["na", "wsddn"])
    ap.add_argument("--fc6-panels", type=int, default=4)
    ap.add_argument("--comm-sms", type=int, default=0, help="N>1: SMs the GEMMs leave to NCCL during the exchange (0 = no reservation)")
    ap.add_argument("--dp-sync", default="auto", choices=["auto", "sharded", "p2p", "allreduce"],
                    help="N>1 gradient exchange: auto = p2p when the ranks can map each other's memory, else NCCL sharded; allreduce = the reference's schedule")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5],
                    help="BASELINE.json configuration: 2 = the headline (default; 4 = the same under --gpus N), 3 = flickr_coco shape, 5 = TTA inference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-isolated", action="store_true", help="skip the isolated RoIPoolF timing after the timed regions")
    ap.add_argument("--ref-rois-per-image", type=int, default=ROIS_PER_IMAGE, help=argparse.SUPPRESS)   # tests shrink the CPU arm
    ap.add_argument("--no-tf32", action="store_true", help="skip the fp32 / TF32 run of the same workload (kernels.tf32_step)")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)          # timing rule: at least three warm-up steps on the device
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl == "ours" and world == 1 and args.gpus > 1):      # the torchrun re-launch keeps the parent's stdout
        _protect_stdout()
    if args.impl == "reference":
        return reference_arm(args)
    if args.config == 3:
        return gpu_config3(args)
    if args.config == 5:
        return gpu_config5(args)
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun on one node
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise RuntimeError("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
