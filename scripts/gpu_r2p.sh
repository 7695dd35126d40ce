#!/bin/bash
# round 2, sixteenth GPU call (8 GPUs): order in which a fused push visits its destinations, A/B at N=8
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
port=29520
for mode in 0 1 2; do
port=$((port+1))
QVM_B200_PUSH_ORDER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --no-e2e --no-configs > gpurun_out/r2p_bench_q33_n8_order$mode.json 2> gpurun_out/r2p_bench_q33_n8_order$mode.err; echo "bench n8 order $mode rc=$?"
tail -c 400 gpurun_out/r2p_bench_q33_n8_order$mode.json
done
