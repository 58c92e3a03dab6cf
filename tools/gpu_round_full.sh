#!/usr/bin/env bash
# One-GPU validation call: the whole GPU suite, smoke(), the headline bench (with its TF32 run), configs 3 and 5, panel-count
# variants, the conv-body microbenchmark, the launch list of the bench and a full ncu capture of the dominant kernel.
#   gpurun --timeout 1200 -- 'bash tools/gpu_round_full.sh r2p'
set -u
TAG="${1:-r2}"; OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
el "pytest -m gpu"
timeout 700 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -s > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log; grep -a "passed\|failed\|TF32 config-2\|unconditioned" $OUT/${TAG}_pytest_gpu.log | tail -n 8 | cut -c1-900
el "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "exit $?"; tail -n 2 $OUT/${TAG}_smoke.log
el "bench default"
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "exit $?"
python tools/bench_brief.py $OUT/${TAG}_bench_n1.json
for pn in 2 8; do
  el "bench --fc6-panels $pn"
  timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-isolated --no-tf32 --fc6-panels $pn > $OUT/${TAG}_bench_n1_p$pn.json 2>/dev/null
  python tools/bench_brief.py $OUT/${TAG}_bench_n1_p$pn.json
done
el "bench --config 3"
timeout 200 python bench.py --config 3 --steps 10 --warmup 3 > $OUT/${TAG}_bench_config3.json 2> $OUT/${TAG}_bench_config3.err; echo "exit $?"; cut -c1-300 $OUT/${TAG}_bench_config3.json
el "bench --config 5"
timeout 300 python bench.py --config 5 --steps 10 > $OUT/${TAG}_bench_config5.json 2> $OUT/${TAG}_bench_config5.err; echo "exit $?"; cut -c1-300 $OUT/${TAG}_bench_config5.json
el "bench --impl reference (2 steps)"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "exit $?"; cut -c1-250 $OUT/${TAG}_bench_ref.json
el "conv body microbenchmark"
timeout 200 python tools/microbench.py convbody > $OUT/${TAG}_microbench_convbody.log 2>&1; echo "exit $?"; cut -c1-200 $OUT/${TAG}_microbench_convbody.log
el "ncu launch list"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated --no-tf32 > $OUT/${TAG}_ncu_launches_bench.log 2>&1
el "ncu full: fc6 dW panel GEMM + implicit conv"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel" -s 40 -c 6 -f -o $OUT/${TAG}_ncu_gemm \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-isolated --no-tf32 > $OUT/${TAG}_ncu_gemm.log 2>&1
el "done"
