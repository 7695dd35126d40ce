#!/bin/bash
# round 2, second GPU call (2 GPUs): full GPU test suite, NVLink probe, bench at N=2 and N=1
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.txt
tail -5 gpurun_out/r2b_pytest_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2b_bench_q33_n2.json 2> gpurun_out/r2b_bench_q33_n2.err; echo "bench n2 rc=$?"
tail -c 3000 gpurun_out/r2b_bench_q33_n2.json
QVM_B200_FUSED_PUSH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --no-configs > gpurun_out/r2b_bench_q33_n2_unfused.json 2> gpurun_out/r2b_bench_q33_n2_unfused.err; echo "bench n2 unfused rc=$?"
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_bench_q33_n1.json 2> gpurun_out/r2b_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 2500 gpurun_out/r2b_bench_q33_n1.json
