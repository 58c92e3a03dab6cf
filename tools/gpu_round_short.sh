#!/usr/bin/env bash
# Short gpurun call (about 9 minutes of box time): GPU parity tests, smoke, both bench arms, the ncu launch
# list of the bench command, one `ncu --set full` capture of the step's kernels, per-kernel microbenchmarks.
# Everything lands in gpurun_out/.  Each stage has its own timeout so that one hang cannot eat the box.
#   gpurun --timeout 660 -- 'bash tools/gpu_round_short.sh r1d'
set -u
TAG="${1:-r1}"
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
el "pytest"
timeout 330 python -m pytest tests -m gpu -x -q --timeout 200 --durations=15 -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -n 6 $OUT/${TAG}_pytest_gpu.log
el "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
el "bench"
timeout 200 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
cut -c1-1800 $OUT/${TAG}_bench_n1.json
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cut -c1-300 $OUT/${TAG}_bench_ref.json
el "ncu launch list (bench command)"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches_bench.log 2>&1
el "ncu full (one step)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|roi_pool|mil_head|sgd_kernel" -s 20 -c 22 -f \
    -o $OUT/${TAG}_ncu_step python tools/ncu_step.py 3 > $OUT/${TAG}_ncu_step.log 2>&1
el "microbench"
timeout 120 python tools/microbench.py pool2 mil testtime > $OUT/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
tail -n 30 $OUT/${TAG}_microbench.log
if [ "${2:-}" = "more" ]; then
  el "variants"
  timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dtype tf32 > $OUT/${TAG}_bench_tf32.json 2> $OUT/${TAG}_bench_tf32.err
  timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --head wsddn > $OUT/${TAG}_bench_wsddn.json 2> $OUT/${TAG}_bench_wsddn.err
fi
el "done"
ls -la $OUT | head -40
