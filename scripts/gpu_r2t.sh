#!/bin/bash
# round 2, twentieth GPU call (2 GPUs): final tree -- full GPU suite (incl. the 2-GPU multi-process case), bench at N = 2
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
QVM_TEST_BIG=1 timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2t_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_gpu.txt
tail -5 gpurun_out/r2t_pytest_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r2t_bench_q33_n2.json 2> gpurun_out/r2t_bench_q33_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/r2t_bench_q33_n2.json
