#!/usr/bin/env bash
# One-GPU evidence call: the whole GPU suite, smoke, the default bench line, the ncu launch list of the bench command and one
# full ncu capture of the GEMMs.      gpurun --timeout 600 -- 'bash tools/gpu_round_final.sh r2u'
set -u
TAG="${1:-r2u}"; OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
el "pytest -m gpu"
timeout 300 python -m pytest tests -m gpu -q --timeout 150 --durations=8 -p no:cacheprovider -rP > $OUT/${TAG}_pytest_gpu_full.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu_full.log
grep -E "relative errors|unconditioned|three-pass vs one-pass|passed|failed|^FAILED|^ERROR" $OUT/${TAG}_pytest_gpu_full.log | cut -c1-1500 > $OUT/${TAG}_pytest_gpu.log
tail -n 4 $OUT/${TAG}_pytest_gpu.log | cut -c1-400
el "smoke"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
tail -n 2 $OUT/${TAG}_smoke.log
el "bench"
timeout 200 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
python tools/bench_brief.py $OUT/${TAG}_bench_n1.json | cut -c1-400
python - <<P
import json
d = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
k = d.get("kernels", {})
print("tf32_step", {a: k.get("tf32_step", {}).get(a) for a in ("ms_per_step", "rois_per_s", "step_tensor_frac")})
print("fp32_step", {a: k.get("fp32_step", {}).get(a) for a in ("ms_per_step", "rois_per_s")})
print("pool isolated", k.get("roi_pool_f_isolated"))
print("clocks", d.get("clocks"))
P
tail -n 3 $OUT/${TAG}_bench_n1.err
if [ "${2:-}" != "noncu" ]; then
el "ncu launch list"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated --no-tf32 > $OUT/${TAG}_ncu_launches_bench.log 2>&1
el "ncu full: GEMMs of one step"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 14 -c 10 -o $OUT/${TAG}_ncu_gemm -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated --no-tf32 > $OUT/${TAG}_ncu_gemm.log 2>&1
fi
el "done"
