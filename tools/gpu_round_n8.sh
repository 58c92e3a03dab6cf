#!/usr/bin/env bash
# One bench.py run on N GPUs (default 8) with the peer-exchange timeline of one instrumented step on stderr.
#   gpurun --gpus 8 --timeout 200 -- 'bash tools/gpu_round_n8.sh r1h 8'
set -u
TAG="${1:-r1}"; N="${2:-8}"
OUT=gpurun_out; mkdir -p $OUT
NAWSOD_P2P_PROFILE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
echo "exit $?"; cut -c1-300 $OUT/${TAG}_bench_n${N}.json; grep -v "^\*\*\*\|OMP_NUM" $OUT/${TAG}_bench_n${N}.err | tail -n 6 | cut -c1-3000
