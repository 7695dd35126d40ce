#!/bin/bash
# round 2, thirteenth GPU call (1 GPU): filtered MEASURE histograms, cold-call tests, full GPU suite + bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2m_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.txt
tail -5 gpurun_out/r2m_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2m_bench_q33_n1.json 2> gpurun_out/r2m_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 1500 gpurun_out/r2m_bench_q33_n1.json
QVM_B200_PATIENT=0 timeout 600 python bench.py --no-configs --steps 1 --warmup 3 > gpurun_out/r2m_bench_q33_n1_impatient.json 2> gpurun_out/r2m_bench_q33_n1_impatient.err; echo "bench impatient rc=$?"
QVM_B200_AHEAD_COMPILE_S=0 timeout 600 python bench.py --no-configs --steps 1 --warmup 3 > gpurun_out/r2m_bench_q33_n1_nodefer.json 2> gpurun_out/r2m_bench_q33_n1_nodefer.err; echo "bench nodefer rc=$?"
