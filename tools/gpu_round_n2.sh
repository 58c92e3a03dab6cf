#!/usr/bin/env bash
# Multi-GPU gpurun call on N = $2 GPUs (default 2): the parity tests of the gradient exchange (schedules against each other AND
# against the oracle's reference schedule) and bench.py under torchrun for each exchange schedule / engine.
#   gpurun --gpus 2 --timeout 420 -- 'bash tools/gpu_round_n2.sh r2b 2'
set -u
TAG="${1:-r2}"; N="${2:-2}"; shift 2 || true
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
el "pytest dp (schedules vs each other, vs the oracle's reference schedule)"
timeout 240 python -m pytest tests/test_gpu_zzzz_dp_2gpu.py tests/test_gpu_zzzz_dp_oracle_schedule.py -m gpu -q -s --timeout 200 -p no:cacheprovider > $OUT/${TAG}_pytest_dp.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_dp.log; grep -a "worst relative\|passed\|failed\|Error" $OUT/${TAG}_pytest_dp.log | tail -n 12
run() {   # name, extra env (VAR=VALUE words), extra bench args
  local name="$1" envs="$2"; shift 2
  el "bench N=$N $name"
  env $envs timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/${TAG}_bench_n${N}_${name}.json 2> $OUT/${TAG}_bench_n${N}_${name}.err
  echo "exit $?"; python tools/bench_brief.py $OUT/${TAG}_bench_n${N}_${name}.json
  grep -a "p2p timeline" $OUT/${TAG}_bench_n${N}_${name}.err | cut -c1-1500
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
    push_sm)   run push_sm "NAWSOD_P2P_RS=push NAWSOD_P2P_ENGINE=sm NAWSOD_P2P_PROFILE=1" ;;
    push_ce)   run push_ce "NAWSOD_P2P_RS=push NAWSOD_P2P_ENGINE=ce NAWSOD_P2P_PROFILE=1" ;;
    sharded)   run sharded "NAWSOD_X=0" --dp-sync sharded ;;
    allreduce) run allreduce "NAWSOD_X=0" --dp-sync allreduce ;;
  esac
done
el "done"
