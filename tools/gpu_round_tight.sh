#!/usr/bin/env bash
# Tight gpurun call (about 6 minutes of box time) for a nearly spent GPU budget, most important stage first:
# GPU parity tests, the default bench line, smoke, then -- if time is left --
# the ncu launch list of the bench command.  Everything lands in gpurun_out/.
#   gpurun --timeout 420 -- 'bash tools/gpu_round_tight.sh r1h'
set -u
TAG="${1:-r1}"
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
el "pytest"
timeout 200 python -m pytest tests -m gpu -q --timeout 100 --durations=10 -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -n 12 $OUT/${TAG}_pytest_gpu.log
el "bench"
timeout 120 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
cut -c1-2600 $OUT/${TAG}_bench_n1.json
tail -n 5 $OUT/${TAG}_bench_n1.err
el "smoke"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
tail -n 3 $OUT/${TAG}_smoke.log
el "microbench pool2"
timeout 40 python tools/microbench.py pool2 > $OUT/${TAG}_microbench_pool.log 2>&1; echo "microbench exit $?"
grep -v "pool_chunks\|pool_rows2" $OUT/${TAG}_microbench_pool.log | cut -c1-160
el "ncu launch list (bench command)"
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated > $OUT/${TAG}_ncu_launches_bench.log 2>&1
el "done"
