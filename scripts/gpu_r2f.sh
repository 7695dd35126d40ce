#!/bin/bash
# round 2, sixth GPU call (2 GPUs): full GPU suite (multi-process world 2 incl. the 16-qubit sharded rho), bench N=2, N=1
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
QVM_TEST_BIG=1 timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.txt
tail -5 gpurun_out/r2f_pytest_gpu.txt
QVM_TEST_BIG=1 QVM_TEST_PORT=29533 QVM_TEST_WORLD=2 bash -c 'for r in 0 1; do QVM_TEST_RANK=$r python tests/_sharded_gpu_worker.py > gpurun_out/r2f_sharded_worker_rank$r.txt 2>&1 & done; wait'
tail -4 gpurun_out/r2f_sharded_worker_rank0.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2f_bench_q33_n2.json 2> gpurun_out/r2f_bench_q33_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2f_bench_q33_n2.json
timeout 900 python bench.py > gpurun_out/r2f_bench_q33_n1.json 2> gpurun_out/r2f_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 1500 gpurun_out/r2f_bench_q33_n1.json
