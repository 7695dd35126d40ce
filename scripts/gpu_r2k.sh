#!/bin/bash
# round 2, eleventh GPU call (1 GPU): GPU test suite + default bench line
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2k_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_gpu.txt
tail -5 gpurun_out/r2k_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2k_bench_q33_n1.json 2> gpurun_out/r2k_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 2500 gpurun_out/r2k_bench_q33_n1.json
