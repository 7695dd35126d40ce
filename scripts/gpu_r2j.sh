#!/bin/bash
# round 2, tenth GPU call (2 GPUs): full GPU suite after the lazy reset, bench N=1 / N=2 (+ reset folding off for A/B)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
QVM_TEST_BIG=1 timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2j_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_gpu.txt
tail -5 gpurun_out/r2j_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2j_bench_q33_n1.json 2> gpurun_out/r2j_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 800 gpurun_out/r2j_bench_q33_n1.json
QVM_B200_FOLD_RESET=0 timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r2j_bench_q33_n1_nofold.json 2> gpurun_out/r2j_bench_q33_n1_nofold.err; echo "bench n1 nofold rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-configs > gpurun_out/r2j_bench_q33_n2.json 2> gpurun_out/r2j_bench_q33_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r2j_bench_q33_n2.json
