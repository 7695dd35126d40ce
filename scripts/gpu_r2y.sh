#!/bin/bash
# round 2 (1 GPU): final tree (group search by shard size, round search): GPU suite, smoke, bench line
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2y_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest_gpu.txt
tail -4 gpurun_out/r2y_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2y_smoke.txt 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2y_smoke.txt
timeout 900 python bench.py > gpurun_out/r2y_bench_q33_n1.json 2> gpurun_out/r2y_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 300 gpurun_out/r2y_bench_q33_n1.json
