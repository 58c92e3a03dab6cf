#!/usr/bin/env bash
# gpurun call for the three-pass fp32 path (about 5 minutes of box time): the whole GPU suite (the GEMM epilogue changed: accumulate
# before bias / activation), with the printed error tables of the full-size TF32 and fp32 head tests, the default bench line
# (kernels.tf32_step and kernels.fp32_step inside it), smoke.
#   gpurun --timeout 420 -- 'bash tools/gpu_round_fp32.sh r2q'
set -u
TAG="${1:-r2q}"
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
el "pytest -m gpu"
timeout 260 python -m pytest tests -m gpu -q --timeout 150 --durations=8 -p no:cacheprovider -rP > $OUT/${TAG}_pytest_gpu_full.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu_full.log
grep -E "relative errors|unconditioned|three-pass vs one-pass|passed|failed|^FAILED|^ERROR|Error" $OUT/${TAG}_pytest_gpu_full.log | cut -c1-1500 > $OUT/${TAG}_pytest_gpu.log
echo "pytest exit (see ${TAG}_pytest_gpu_full.log)" >> $OUT/${TAG}_pytest_gpu.log
cat $OUT/${TAG}_pytest_gpu.log
el "bench"
timeout 150 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
python tools/bench_brief.py $OUT/${TAG}_bench_n1.json | cut -c1-400
python - <<P
import json
d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
k = d.get("kernels", {})
print("tf32_step", {a: k.get("tf32_step", {}).get(a) for a in ("ms_per_step", "rois_per_s", "step_tensor_frac")})
print("fp32_step", k.get("fp32_step"))
P
tail -n 5 $OUT/${TAG}_bench_n1.err
el "smoke"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
tail -n 3 $OUT/${TAG}_smoke.log
el "done"
