#!/usr/bin/env bash
# RoIPoolF iteration on one GPU: the op's parity tests, the microbenchmark on the bench shapes, optionally an ncu capture.
#   gpurun --timeout 300 -- 'bash tools/gpu_round_pool.sh r2d [ncu]'
set -u
TAG="${1:-r2}"; OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "roi_pool or RoIPool or pool" > $OUT/${TAG}_pytest_pool.log 2>&1
echo "pytest exit $?"; tail -n 6 $OUT/${TAG}_pytest_pool.log
timeout 120 python tools/microbench.py pool2 > $OUT/${TAG}_microbench_pool.log 2>&1; echo "exit $?"; cut -c1-170 $OUT/${TAG}_microbench_pool.log
if [ "${2:-}" = "ncu" ]; then
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rows2" -s 1 -c 1 -f -o $OUT/${TAG}_ncu_pool python tools/ncu_pool.py > $OUT/${TAG}_ncu_pool.log 2>&1
  echo "ncu exit $?"
fi
