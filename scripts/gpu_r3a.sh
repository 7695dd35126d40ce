#!/bin/bash
# round 2 (2 GPUs): which of the two searches slows the exchange-carrying pass at N = 2 (65.9 ms against 48.5 before)?
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
port=29570
for combo in "2 1" "4 0" "0 1"; do
set -- $combo
port=$((port+1))
QVM_B200_PLAN_SEARCH=$1 QVM_B200_ROUND_SEARCH=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --no-e2e --no-configs > gpurun_out/r3a_n2_plan$1_round$2.json 2> gpurun_out/r3a_n2_plan$1_round$2.err; echo "rc=$?"
tail -c 200 gpurun_out/r3a_n2_plan$1_round$2.json
done
