#!/bin/bash
# round 2 (1 GPU): group search of the planner (qvm_plan_choose_group), A/B on one box: depth 2 (default) / 0 / 3, then the GPU suite
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python bench.py --no-configs > gpurun_out/r2v_bench_q33_n1_depth2.json 2> gpurun_out/r2v_bench_q33_n1_depth2.err; echo "bench depth2 rc=$?"
tail -c 300 gpurun_out/r2v_bench_q33_n1_depth2.json
QVM_B200_PLAN_SEARCH=0 timeout 600 python bench.py --no-e2e --no-configs --no-cpu-baseline > gpurun_out/r2v_bench_q33_n1_depth0.json 2> gpurun_out/r2v_bench_q33_n1_depth0.err; echo "bench depth0 rc=$?"
QVM_B200_PLAN_SEARCH=3 timeout 600 python bench.py --no-e2e --no-configs --no-cpu-baseline > gpurun_out/r2v_bench_q33_n1_depth3.json 2> gpurun_out/r2v_bench_q33_n1_depth3.err; echo "bench depth3 rc=$?"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2v_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest_gpu.txt
tail -4 gpurun_out/r2v_pytest_gpu.txt
