#!/bin/bash
# round 2, fourteenth GPU call (1 GPU): handle parking (full suite, config 2 A/B), cold-call A/B on ONE seed
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2n_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_gpu.txt
tail -5 gpurun_out/r2n_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2n_bench_q33_n1.json 2> gpurun_out/r2n_bench_q33_n1.err; echo "bench n1 rc=$?"
tail -c 1500 gpurun_out/r2n_bench_q33_n1.json
QVM_B200_PARK=0 timeout 600 python bench.py --no-e2e --steps 1 --warmup 3 > gpurun_out/r2n_bench_nopark.json 2> gpurun_out/r2n_bench_nopark.err; echo "bench nopark rc=$?"
QVM_B200_CACHE_DIR=/tmp/qc1 timeout 600 python bench.py --no-configs --steps 1 --warmup 3 --cold-seed 777 > gpurun_out/r2n_cold_patient_defer.json 2> gpurun_out/r2n_cold_patient_defer.err
QVM_B200_CACHE_DIR=/tmp/qc2 QVM_B200_AHEAD_COMPILE_S=0 timeout 600 python bench.py --no-configs --steps 1 --warmup 3 --cold-seed 777 > gpurun_out/r2n_cold_patient_nodefer.json 2> gpurun_out/r2n_cold_patient_nodefer.err
QVM_B200_CACHE_DIR=/tmp/qc3 QVM_B200_PATIENT=0 timeout 600 python bench.py --no-configs --steps 1 --warmup 3 --cold-seed 777 > gpurun_out/r2n_cold_impatient.json 2> gpurun_out/r2n_cold_impatient.err
echo done
