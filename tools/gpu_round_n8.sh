#!/usr/bin/env bash
# Multi-GPU bench matrix on N GPUs (default 8): sync=auto (peer path after its self-test, outcome in config.p2p_selftest,
# with the peer-exchange timeline of one instrumented step on stderr), NCCL sharded, the reference's all-reduce schedule,
# and -- experimental -- the peer path with the GEMM-fused scatter.  Each run has its own timeout.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round_n8.sh r2n8 8'
set -u
TAG="${1:-r2}"; N="${2:-8}"
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
run() {   # name, extra env (as VAR=VALUE words), extra bench args
  local name="$1" envs="$2"; shift 2
  el "bench N=$N $name"
  env $envs timeout 220 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  echo "exit $?"; cut -c1-300 $OUT/${TAG}_bench_n${N}_${name}.json
  grep -o '"p2p_selftest": [^,]*' $OUT/${TAG}_bench_n${N}_${name}.json
  grep -v "^\*\*\*\|OMP_NUM" $OUT/${TAG}_bench_n${N}_${name}.err | tail -n 4 | cut -c1-2000
}
run auto "NAWSOD_P2P_PROFILE=1"
run sharded "NAWSOD_X=0" --dp-sync sharded
run allreduce "NAWSOD_X=0" --dp-sync allreduce
run fused_scatter "NAWSOD_P2P_FUSED_SCATTER=1 NAWSOD_P2P_PROFILE=1"
el "done"
