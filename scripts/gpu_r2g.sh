#!/bin/bash
# round 2, seventh GPU call (2 GPUs): the extra-leg code path of bench.py (34 qubits on 2 GPUs: 128 GiB shards, no
# second buffer -> in-place exchange) and the refreshed ahead-of-time cache
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --config5-world 2 --config5-qubits 34 > gpurun_out/r2g_bench_q33_n2_c5.json 2> gpurun_out/r2g_bench_q33_n2_c5.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r2g_bench_q33_n2_c5.json
tail -5 gpurun_out/r2g_bench_q33_n2_c5.err
