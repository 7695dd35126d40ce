#!/bin/bash
# round 2 (1 GPU): round search of the encoder (fewer register-tile rounds per group), A/B on one box, then the GPU suite
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2w_bench_q33_n1.json 2> gpurun_out/r2w_bench_q33_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2w_bench_q33_n1.json
QVM_B200_ROUND_SEARCH=0 timeout 600 python bench.py --no-e2e --no-configs --no-cpu-baseline > gpurun_out/r2w_bench_q33_n1_noroundsearch.json 2> gpurun_out/r2w_bench_q33_n1_noroundsearch.err; echo "bench norounds rc=$?"
QVM_B200_PLAN_SEARCH=4 timeout 600 python bench.py --no-e2e --no-configs --no-cpu-baseline > gpurun_out/r2w_bench_q33_n1_depth4.json 2> gpurun_out/r2w_bench_q33_n1_depth4.err; echo "bench depth4 rc=$?"
QVM_B200_PLAN_SEARCH=3 timeout 600 python bench.py --no-e2e --no-configs --no-cpu-baseline > gpurun_out/r2w_bench_q33_n1_depth3.json 2> gpurun_out/r2w_bench_q33_n1_depth3.err; echo "bench depth3 rc=$?"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2w_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest_gpu.txt
tail -4 gpurun_out/r2w_pytest_gpu.txt
