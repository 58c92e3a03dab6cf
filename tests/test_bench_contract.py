"""bench.py contract on CPU: the reference arm (the oracle port timed on the host cores) prints exactly ONE JSON
line on stdout with the keys the driver reads; ranks other than 0 print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                           "--warmup", "1", "--ref-rois-per-image", "250"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run({"OMP_NUM_THREADS": "1"})          # what torchrun exports to its workers: the arm must override it
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "RoIs/sec (fwd+bwd, WSDDN head)" and d["unit"] == "RoIs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "RoIs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] == 2 and d["warmup"] == 1                               # --steps / --warmup honoured
    assert d["config"]["threads"] == len(os.sched_getaffinity(0)) == cb["cores"]
    assert d["config1_cpu_forward"]["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_line_assembly():
    """bench.assemble_line is everything the GPU arm does AFTER measuring: run it on the round's recorded measurements
    (profiles/r1k_bench_n1.json) and check the contract keys, the arithmetic and that the line serialises."""
    sys.path.insert(0, ROOT)
    import bench
    rec = json.load(open(os.path.join(ROOT, "profiles", "r1k_bench_n1.json")))
    steps, R = rec["steps"], 4000
    k = rec["kernels"]
    for world, sync, selftest in ((1, "local", None), (8, "p2p", "ok"), (4, "sharded", "a wait kernel timed out after 3000 ms")):
        line = bench.assemble_line(
            steps=steps, warmup=rec["warmup"], world=world, R=R, S=2, bf16=True, noise=True, ms_total=rec["ms_per_step"] * steps,
            ms_e2e=rec["e2e"]["ms_per_step"] * steps, h2d_bytes=rec["e2e"]["h2d_bytes_per_step"], d2h_bytes=rec["e2e"]["d2h_bytes_per_step"],
            launches=rec["gpu_launches"], clocks=rec["clocks"],
            kernel_ms={"fc6_fwd": k["fc6_fwd"]["ms"], "fc6_bwd_w": k["fc6_bwd_w"]["ms"] / 4, "roi_pool_f": k["roi_pool_f"]["ms"],
                       "mil_head": k["mil_head"]["ms"]},
            n_panels=4, iso={"step_config": 0.0809, "fp32_train": 0.238} if world == 1 else {}, cpu=rec["cpu_baseline"] if world == 1 else None,
            loss=rec["loss"], dp_info={"sync": sync, "fc6_panels": 4, "p2p_selftest": selftest, "engine": "ce" if sync == "p2p" else None})
        d = json.loads(json.dumps(line))
        assert d["metric"] == "RoIs/sec (fwd+bwd, WSDDN head)" and d["unit"] == "RoIs/s" and d["n_gpus"] == world
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "bf16"
        assert abs(d["value"] - world * rec["value"]) <= 1e-6 * world * rec["value"]            # whole-job aggregate
        assert abs(d["e2e"]["value"] - world * rec["e2e"]["value"]) <= 1e-6 * world * rec["e2e"]["value"]
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        rf = d["roofline"]
        assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
        assert abs(rf["achieved"] - rec["roofline"]["achieved"]) <= 1e-6 * rec["roofline"]["achieved"]
        # per launch: the ncu capture of a row-panel launch (0.400 GB) against its algorithmic bytes (0.423 GB)
        assert rf["launches_per_step"] == 4 and 0.39e9 < rf["traffic"] < 0.41e9
        assert abs(rf["algorithmic_bytes_per_launch"] - (4000 * 8192 * 2 / 4 + 4000 * 25088 * 2 + 8192 * 25088 * 4 / 4)) < 1
        assert "workload" in d["config"] and "model" not in d["config"] and d["config"]["p2p_selftest"] == selftest
        assert ("none" in d["config"]["parallelism"]) == (world == 1)
        assert ("roi_pool_f_isolated" in d["kernels"]) == (world == 1) and (d["cpu_baseline"] is None) == (world != 1)
    # unpanelled launch: the other capture; an uncaptured panel count: unknown
    common = dict(steps=2, warmup=3, world=1, R=R, S=2, bf16=True, noise=True, ms_total=8.0, ms_e2e=9.0, h2d_bytes=1, d2h_bytes=1, launches=1,
                  clocks=None, kernel_ms={"fc6_bwd_w": 1.2}, iso={}, cpu=None, loss=[0.0],
                  dp_info={"sync": "local", "fc6_panels": 1, "p2p_selftest": None, "engine": None})
    assert bench.assemble_line(n_panels=1, **common)["roofline"]["traffic"] > 3e9
    assert bench.assemble_line(n_panels=2, **common)["roofline"]["traffic"] is None
    # the fp32 / TF32 run of the same workload rides in the same line
    t = bench.assemble_line(n_panels=4, tf32={"ms_total": 160.0, "steps": 20, "kernel_ms": {"fc6_fwd": 2.2, "fc6_bwd_w": 0.8}, "n_panels": 4},
                            **common)["kernels"]["tf32_step"]
    assert t["ms_per_step"] == 8.0 and abs(t["rois_per_s"] - 500000.0) < 1e-6 and 0 < t["fc6_fwd"]["frac_tensor"] < 1.5
