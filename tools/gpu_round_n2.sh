#!/usr/bin/env bash
# Multi-GPU gpurun call (N = $2 GPUs, default 2; NAWSOD_EXPERIMENTAL=1 adds the GEMM-fused scatter variant): the 2-GPU parity test of the gradient exchange and bench.py under
# torchrun for each exchange schedule.    gpurun --gpus 2 --timeout 300 -- 'bash tools/gpu_round_n2.sh r1e 2'
set -u
TAG="${1:-r1}"; N="${2:-2}"
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
el "pytest dp"
timeout 150 python -m pytest tests/test_gpu_zzzz_dp_2gpu.py -m gpu -x -q --timeout 140 -p no:cacheprovider > $OUT/${TAG}_pytest_dp.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_dp.log; tail -n 3 $OUT/${TAG}_pytest_dp.log
for sync in auto sharded allreduce; do
  el "bench N=$N sync=$sync"
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 --dp-sync $sync > $OUT/${TAG}_bench_n${N}_${sync}.json 2> $OUT/${TAG}_bench_n${N}_${sync}.err
  echo "exit $?"; cut -c1-420 $OUT/${TAG}_bench_n${N}_${sync}.json; tail -n 3 $OUT/${TAG}_bench_n${N}_${sync}.err
done
el "done"
