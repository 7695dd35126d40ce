#!/bin/bash
# round 2, final confirmation on ONE GPU: what the driver runs at round end -- the GPU suite, smoke(), the bench line and
# the reference arm
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2u_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest_gpu.txt
tail -4 gpurun_out/r2u_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2u_smoke.txt 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r2u_smoke.txt
timeout 900 python bench.py > gpurun_out/r2u_bench_q33_n1.json 2> gpurun_out/r2u_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 400 gpurun_out/r2u_bench_q33_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err; echo "bench reference rc=$?"
tail -c 600 gpurun_out/r2u_bench_reference.json
