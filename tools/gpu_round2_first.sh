#!/usr/bin/env bash
# FIRST GPU call of round 2: everything that was written after round 1's GPU budget ran out, measured in one go.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r2a'
# 1. the regular GPU suite (incl. the new UpdateReduce / reference-kernel tests), 2. the experimental fused kernels
# (NAWSOD_EXPERIMENTAL=1), 3. bench.py default vs NAWSOD_FUSED_SGD=1 vs sgd_max_ctas sweep, 4. ncu launch list and one full
# capture of the fused kernel.  Each stage has its own timeout; a hang in the experimental kernel cannot eat the box.
set -u
TAG="${1:-r2a}"
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
el "pytest (regular suite)"
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log; tail -n 5 $OUT/${TAG}_pytest_gpu.log
el "pytest (experimental fused kernels)"
NAWSOD_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental_fused.py -m gpu -q --timeout 120 -p no:cacheprovider \
    > $OUT/${TAG}_pytest_fused.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_fused.log; tail -n 25 $OUT/${TAG}_pytest_fused.log
el "pytest (experimental conv body: im2col / max-pool kernels + the VGG16 body on the tcgen05 GEMM)"
NAWSOD_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental_conv_body.py -m gpu -q --timeout 120 -p no:cacheprovider \
    > $OUT/${TAG}_pytest_conv_body.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_conv_body.log; tail -n 15 $OUT/${TAG}_pytest_conv_body.log
el "bench default"
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "exit $?"
cut -c1-400 $OUT/${TAG}_bench_n1.json
for v in "--dtype tf32" "--head wsddn"; do      # fp32 (TF32) path and the single-stack WSDDN head
  n=$(echo $v | tr -d ' -'); el "bench $v"
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-isolated $v > $OUT/${TAG}_bench_n1_$n.json 2> $OUT/${TAG}_bench_n1_$n.err; echo "exit $?"
  cut -c1-200 $OUT/${TAG}_bench_n1_$n.json
done
el "bench fused SGD"
NAWSOD_FUSED_SGD=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-isolated \
    > $OUT/${TAG}_bench_n1_fused_sgd.json 2> $OUT/${TAG}_bench_n1_fused_sgd.err; echo "exit $?"
cut -c1-400 $OUT/${TAG}_bench_n1_fused_sgd.json; tail -n 3 $OUT/${TAG}_bench_n1_fused_sgd.err
el "bench bias gradients on a side stream"
NAWSOD_BIAS_SIDE_STREAM=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-isolated \
    > $OUT/${TAG}_bench_n1_bias_side.json 2> $OUT/${TAG}_bench_n1_bias_side.err; echo "exit $?"
cut -c1-200 $OUT/${TAG}_bench_n1_bias_side.json
for c in 148 296 592; do
  el "bench sgd_max_ctas=$c"
  NAWSOD_TUNING=sgd_max_ctas=$c timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-isolated \
      > $OUT/${TAG}_bench_n1_sgdctas$c.json 2> $OUT/${TAG}_bench_n1_sgdctas$c.err; echo "exit $?"
  cut -c1-200 $OUT/${TAG}_bench_n1_sgdctas$c.json
done
el "pool microbenchmark (incl. the pool_skip_idle / pool_prefetch_roi / pool_lean variants)"
timeout 200 python tools/microbench.py pool2 > $OUT/${TAG}_microbench_pool.log 2>&1; echo "exit $?"; grep "skip_idle\|prefetch\|lean\|{}:" $OUT/${TAG}_microbench_pool.log | cut -c1-160
el "conv body microbenchmark (experimental)"
timeout 200 python tools/microbench.py convbody > $OUT/${TAG}_microbench_convbody.log 2>&1; echo "exit $?"; cut -c1-200 $OUT/${TAG}_microbench_convbody.log
el "ncu launch list (fused)"
NAWSOD_FUSED_SGD=1 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_ncu_launches_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated \
    > $OUT/${TAG}_ncu_launches_fused.log 2>&1
el "ncu full (fused kernel)"
NAWSOD_FUSED_SGD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_dw_fused" -s 8 -c 2 -f \
    -o $OUT/${TAG}_ncu_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated > $OUT/${TAG}_ncu_fused.log 2>&1
el "done"; ls -la $OUT | tail -n 20
