#!/bin/bash
# round 2, last GPU minutes (1 GPU): first calls with the planner's choices replayed from the walk ahead, then the suite
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python bench.py --no-configs --no-cpu-baseline > gpurun_out/r3d_bench_q33_n1.json 2> gpurun_out/r3d_bench_q33_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r3d_bench_q33_n1.json
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r3d_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r3d_pytest_gpu.txt
tail -3 gpurun_out/r3d_pytest_gpu.txt
