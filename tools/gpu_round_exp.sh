#!/usr/bin/env bash
# Experiments of one call: TMA-store epilogue (parity, bench off / on).  (Call r2w also measured a CTA-pair form of the convolution:
# no faster, removed -- profiles/r2w_microbench_convbody_pair*.log.)
#   gpurun --timeout 400 -- 'bash tools/gpu_round_exp.sh r2w'
set -u
TAG="${1:-r2w}"; OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
el "parity: TMA-store epilogue, pair convolution"
timeout -k 5 150 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv_body.py -m gpu -q -p no:cacheprovider --timeout 60 -k "tma_store or cta_pair" > $OUT/${TAG}_pytest_exp.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_exp.log; tail -n 12 $OUT/${TAG}_pytest_exp.log | cut -c1-300
for tma in 0 1; do
  el "bench gemm_tma_store=$tma"
  NAWSOD_TUNING=gemm_tma_store=$tma timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-tf32 --no-isolated > $OUT/${TAG}_bench_n1_tma${tma}.json 2> $OUT/${TAG}_bench_n1_tma${tma}.err
  echo "exit $?"; python tools/bench_brief.py $OUT/${TAG}_bench_n1_tma${tma}.json | cut -c1-330; tail -n 3 $OUT/${TAG}_bench_n1_tma${tma}.err
done
el "done"
