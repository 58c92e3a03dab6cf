#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench (both arms + variants), per-kernel microbenchmarks and
# the ncu captures.  Everything lands in gpurun_out/ (merged back by gpurun).  Each stage has its own timeout
# so that one hang cannot eat the box.
#   gpurun --timeout 1100 -- 'bash tools/gpu_round.sh r1c'
set -u
TAG="${1:-r1}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest" ; date +%s
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --durations=20 -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -n 15 $OUT/${TAG}_pytest_gpu.log
echo "== smoke" ; date +%s
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
echo "== bench" ; date +%s
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
cat $OUT/${TAG}_bench_n1.json | cut -c1-1500
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
for v in "--dtype tf32" "--fc6-panels 1"; do
  n=$(echo $v | tr -d ' -')
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $v > $OUT/${TAG}_bench_${n}.json 2> $OUT/${TAG}_bench_${n}.err
done
NAWSOD_LOCAL_PIPELINE=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_nopipeline.json 2> $OUT/${TAG}_bench_nopipeline.err
NAWSOD_TUNING=pool_rowcache=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_norowcache.json 2> $OUT/${TAG}_bench_norowcache.err
echo "== microbench" ; date +%s
timeout 300 python tools/microbench.py pool2 mil testtime > $OUT/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
tail -n 40 $OUT/${TAG}_microbench.log
echo "== ncu" ; date +%s
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_ncu_launches.csv python tools/ncu_step.py 3 > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"roi_pool|mil_head|sgd_kernel" -c 8 -f -o $OUT/${TAG}_ncu_small python tools/ncu_step.py 2 > $OUT/${TAG}_ncu_small.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"roi_pool" -c 4 -f -o $OUT/${TAG}_ncu_pool python tools/ncu_pool.py > $OUT/${TAG}_ncu_pool.log 2>&1
date +%s
ls -la $OUT | head -50
