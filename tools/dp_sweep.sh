#!/bin/bash
# N-GPU data-parallel A/B runs of bench.py (train numbers only)
# usage: tools/dp_sweep.sh NGPU OUT_PREFIX [labels...]
N=${1:-2}; OUT=${2:-gpurun_out/dp_sweep}; shift 2
port=29600
run() {  # label, env...
  label=$1; shift
  port=$((port+1))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 8 --warmup 3 --no-roofline > ${OUT}_${label}.json 2> ${OUT}_${label}.err
  python - "$label" "${OUT}_${label}.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]).read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],2), "cells/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for label in "$@"; do
  case $label in
    multicast) run multicast A=1 ;;
    p2p_stores) run p2p_stores CELLCOMM_B200_MULTICAST=0 ;;
    skip_update) run skip_update CELLCOMM_B200_DP_SKIP_UPDATE=1 ;;
    nccl) run nccl CELLCOMM_B200_PEER_OPT=0 ;;
    sync) run sync CELLCOMM_B200_ASYNC_OPT=0 ;;
    graph) run graph A=1 ;;
    eager) run eager CELLCOMM_B200_DP_GRAPH=0 ;;
    nccl_allreduce) run nccl_allreduce CELLCOMM_B200_PEER_ALLREDUCE=0 ;;
  esac
done
