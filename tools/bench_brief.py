"""One-line digest of a bench.py JSON line (used by the tools/gpu_round*.sh scripts)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print("%s: unreadable (%s)" % (path, e))
        continue
    k = d.get("kernels", {})
    t = d.get("timing") or {}
    print("%s: N=%s %.0f RoIs/s  %.3f ms/step  e2e %.3f ms  dW %s ms (%.2f)  fwd %s ms  pool %s ms  selftest=%s  blocks=%s host-enqueue %s ms  e2e-step min/med/max=%s" % (
        path.split("/")[-1], d.get("n_gpus"), d.get("value", 0), d.get("ms_per_step", 0), d.get("e2e", {}).get("ms_per_step", 0),
        ("%.3f" % k["fc6_bwd_w"]["ms"]) if k.get("fc6_bwd_w", {}).get("ms") else "?", k.get("fc6_bwd_w", {}).get("frac_tensor") or 0,
        ("%.3f" % k["fc6_fwd"]["ms"]) if k.get("fc6_fwd", {}).get("ms") else "?",
        ("%.3f" % k["roi_pool_f"]["ms"]) if k.get("roi_pool_f", {}).get("ms") else "?",
        d.get("config", {}).get("p2p_selftest"), (t.get("value") or {}).get("blocks"),
        round((t.get("value") or {}).get("host_enqueue_ms_per_step_this_rank") or 0, 3),
        {a: round(b, 3) for a, b in ((t.get("e2e") or {}).get("step_ms_this_rank") or {}).items()}))
