#!/bin/bash
# round 2, eighth GPU call (8 GPUs): multi-process sharded tests on 2 / 4 / 8 GPUs, bench at N=8 (with the 36-qubit
# leg), N=4, and the in-place exchange at N=8 for comparison
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt
QVM_TEST_BIG=1 timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -v -p no:cacheprovider > gpurun_out/r2h_pytest_sharded_2_4_8.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_sharded_2_4_8.txt
tail -6 gpurun_out/r2h_pytest_sharded_2_4_8.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2h_bench_q33_n8.json 2> gpurun_out/r2h_bench_q33_n8.err; echo "bench n8 rc=$?"
tail -c 1200 gpurun_out/r2h_bench_q33_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 > gpurun_out/r2h_bench_q33_n4.json 2> gpurun_out/r2h_bench_q33_n4.err; echo "bench n4 rc=$?"
tail -c 600 gpurun_out/r2h_bench_q33_n4.json
QVM_B200_SWAP=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --no-e2e --no-configs > gpurun_out/r2h_bench_q33_n8_p2p.json 2> gpurun_out/r2h_bench_q33_n8_p2p.err; echo "bench n8 p2p rc=$?"
tail -c 600 gpurun_out/r2h_bench_q33_n8_p2p.json
