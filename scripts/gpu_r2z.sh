#!/bin/bash
# round 2, last 8-GPU GPU call (8 GPUs): multi-process sharded tests on 2 / 4 / 8 GPUs, then the bench at N = 8 (with the
# 36-qubit leg) and 4
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2z_gpus.txt
QVM_TEST_BIG=1 timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -v -p no:cacheprovider > gpurun_out/r2z_pytest_sharded_2_4_8.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest_sharded_2_4_8.txt
tail -6 gpurun_out/r2z_pytest_sharded_2_4_8.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 > gpurun_out/r2z_bench_q33_n8.json 2> gpurun_out/r2z_bench_q33_n8.err; echo "bench n8 rc=$?"
tail -c 800 gpurun_out/r2z_bench_q33_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 > gpurun_out/r2z_bench_q33_n4.json 2> gpurun_out/r2z_bench_q33_n4.err; echo "bench n4 rc=$?"
tail -c 400 gpurun_out/r2z_bench_q33_n4.json
