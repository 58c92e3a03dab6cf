#!/usr/bin/env bash
# Multi-GPU call on N GPUs (default 8): parity of every exchange schedule against the oracle's reference schedule at world N,
# then the bench matrix over the peer exchange's engines / knobs and the NCCL schedule, each with the peer-exchange timeline of
# one instrumented step.     gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_round_n8.sh r2c 8'
set -u
TAG="${1:-r2}"; N="${2:-8}"; shift 2 || true
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
el "pytest dp (world $N vs the oracle's reference schedule)"
env ${NAWSOD_DP_TEST_ENV:-NAWSOD_X=0} timeout 240 python -m pytest tests/test_gpu_zzzz_dp_oracle_schedule.py -m gpu -q -s --timeout 220 -p no:cacheprovider > $OUT/${TAG}_pytest_dp_n${N}.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_dp_n${N}.log; grep -a "worst relative\|passed\|failed\|Error" $OUT/${TAG}_pytest_dp_n${N}.log | tail -n 12
run() {   # name, extra env (VAR=VALUE words), extra bench args
  local name="$1" envs="$2"; shift 2
  el "bench N=$N $name"
  env $envs timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  echo "exit $?"; python tools/bench_brief.py $OUT/${TAG}_bench_n${N}_${name}.json
  grep -a "p2p timeline" $OUT/${TAG}_bench_n${N}_${name}.err | cut -c1-1800
}
for v in "$@"; do
  case $v in
    pull_sm)   run pull_sm "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=sm NAWSOD_P2P_PROFILE=1" ;;
    pull_sm_nogate) run pull_sm_nogate "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=sm NAWSOD_P2P_GATED_FC6=0 NAWSOD_P2P_PROFILE=1" ;;
    pull_ce_p8) run pull_ce_p8 "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=ce NAWSOD_P2P_PROFILE=1" --fc6-panels 8 ;;
    pull_sm_p8) run pull_sm_p8 "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=sm NAWSOD_P2P_PROFILE=1" --fc6-panels 8 ;;
    pull_sm32) run pull_sm32 "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=sm NAWSOD_TUNING=p2p_ctas=32 NAWSOD_P2P_PROFILE=1" ;;
    pull_sm32_p8) run pull_sm32_p8 "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=sm NAWSOD_TUNING=p2p_ctas=32 NAWSOD_P2P_PROFILE=1" --fc6-panels 8 ;;
    pull_ce)   run pull_ce "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=ce NAWSOD_P2P_PROFILE=1" ;;
    pull_sm_sgd296) run pull_sm_sgd296 "NAWSOD_P2P_RS=pull NAWSOD_P2P_ENGINE=sm NAWSOD_TUNING=p2p_ctas=32,sgd_max_ctas=296 NAWSOD_P2P_PROFILE=1" ;;
    sm)        run sm "NAWSOD_P2P_RS=push NAWSOD_P2P_ENGINE=sm NAWSOD_P2P_PROFILE=1" ;;
    ce)        run ce "NAWSOD_P2P_RS=push NAWSOD_P2P_ENGINE=ce NAWSOD_P2P_PROFILE=1" ;;
    ce7)       run ce7 "NAWSOD_P2P_ENGINE=ce NAWSOD_P2P_COPY_STREAMS=7 NAWSOD_P2P_PROFILE=1" ;;
    sm32)      run sm32 "NAWSOD_P2P_RS=push NAWSOD_P2P_ENGINE=sm NAWSOD_TUNING=p2p_ctas=32 NAWSOD_P2P_PROFILE=1" ;;
    sm74)      run sm74 "NAWSOD_P2P_ENGINE=sm NAWSOD_TUNING=p2p_ctas=74 NAWSOD_P2P_PROFILE=1" ;;
    ce_p8)     run ce_p8 "NAWSOD_P2P_ENGINE=ce NAWSOD_P2P_PROFILE=1" --fc6-panels 8 ;;
    sm_p8)     run sm_p8 "NAWSOD_P2P_ENGINE=sm NAWSOD_P2P_PROFILE=1" --fc6-panels 8 ;;
    default)   run default "NAWSOD_P2P_PROFILE=1" ;;
    sharded)   run sharded "NAWSOD_X=0" --dp-sync sharded ;;
    allreduce) run allreduce "NAWSOD_X=0" --dp-sync allreduce ;;
  esac
done
el "done"
