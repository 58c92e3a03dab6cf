#!/usr/bin/env bash
# CTA-pair GEMM (tcgen05 cta_group::2): parity against the one-CTA kernel under a hard time-out (a first run of new barrier
# code must not be able to hang the box), then the bench with the knob off / on.
#   gpurun --timeout 400 -- 'bash tools/gpu_round_pair.sh r2t'
set -u
TAG="${1:-r2t}"; OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
el "pair parity"
timeout -k 5 120 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider --timeout 60 -k "cta_pair" > $OUT/${TAG}_pytest_pair.log 2>&1
rc=$?; echo "pytest exit $rc" | tee -a $OUT/${TAG}_pytest_pair.log; tail -n 15 $OUT/${TAG}_pytest_pair.log | cut -c1-300
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,utilization.gpu,memory.used --format=csv; el "stopping: the pair kernel is not correct yet"; exit 0; fi
for pair in 0 1; do
  el "bench gemm_pair=$pair"
  NAWSOD_TUNING=gemm_pair=$pair timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-tf32 > $OUT/${TAG}_bench_n1_pair${pair}.json 2> $OUT/${TAG}_bench_n1_pair${pair}.err
  echo "exit $?"; python tools/bench_brief.py $OUT/${TAG}_bench_n1_pair${pair}.json | cut -c1-330; tail -n 3 $OUT/${TAG}_bench_n1_pair${pair}.err
done
el "full gemm + head tests with the pair kernel as default knob"
NAWSOD_TUNING=gemm_pair=1 timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_head.py -m gpu -q -p no:cacheprovider --timeout 150 -k "not cta_pair" > $OUT/${TAG}_pytest_pair_default.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_pair_default.log; tail -n 6 $OUT/${TAG}_pytest_pair_default.log | cut -c1-300
el "done"
