#!/usr/bin/env bash
# Short gpurun call for the RoIPoolF kernels only: parity tests of every forward variant, the pool microbenchmark
# (old bin-row kernel vs rows2) and one `ncu --set full` capture of the rows2 kernel (fp32-train and bf16-infer).
#   gpurun --timeout 200 -- 'bash tools/gpu_round_pool.sh r1j'
set -u
TAG="${1:-r1}"
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
el "pytest roi_pool"
timeout 100 python -m pytest tests/test_gpu_ops.py -m gpu -q -k roi_pool --timeout 60 -p no:cacheprovider > $OUT/${TAG}_pytest_pool.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_pool.log
tail -n 8 $OUT/${TAG}_pytest_pool.log
el "microbench pool2"
timeout 60 python tools/microbench.py pool2 > $OUT/${TAG}_microbench_pool.log 2>&1; echo "microbench exit $?"
cat $OUT/${TAG}_microbench_pool.log | cut -c1-200
el "ncu full rows2"
NAWSOD_TUNING=pool_rows2=1 timeout 70 ncu --set full --clock-control none --import-source on -k regex:"roi_pool" -c 4 -f -o $OUT/${TAG}_ncu_pool \
    python tools/ncu_pool.py > $OUT/${TAG}_ncu_pool.log 2>&1; echo "ncu exit $?"
el "done"
