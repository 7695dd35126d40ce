#!/bin/bash
# round 2, last 2-GPU GPU call (2 GPUs): the bench at N = 2 and N = 1 on the same box
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 > gpurun_out/r3b_bench_q33_n2.json 2> gpurun_out/r3b_bench_q33_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/r3b_bench_q33_n2.json
echo skip n1

