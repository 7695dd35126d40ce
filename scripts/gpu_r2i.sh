#!/bin/bash
# round 2, ninth GPU call (1 GPU): full GPU suite, then ncu -- launch list at 30 qubits and one --set full capture of the
# dominant kernel at 28 qubits (each only after the same command exited 0 without ncu)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2i_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_gpu.txt
tail -4 gpurun_out/r2i_pytest_gpu.txt
CMD30="python bench.py --qubits 30 --depth 40 --no-e2e --no-cpu-baseline --no-configs --steps 2 --warmup 3"
$CMD30 > gpurun_out/r2i_plain_q30.json 2> gpurun_out/r2i_plain_q30.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2i_launches_q30.csv $CMD30 > gpurun_out/r2i_ncu_launch.log 2>&1
tail -2 gpurun_out/r2i_ncu_launch.log
CMD28="python bench.py --qubits 28 --depth 40 --no-e2e --no-cpu-baseline --no-configs --steps 2 --warmup 3"
$CMD28 > gpurun_out/r2i_plain_q28.json 2> gpurun_out/r2i_plain_q28.err &&
ncu --set full --clock-control none --import-source on -k regex:spec_kernel --launch-skip 40 -c 2 -f -o gpurun_out/r2i_prof_spec $CMD28 > gpurun_out/r2i_ncu_full.log 2>&1
tail -3 gpurun_out/r2i_ncu_full.log
ls -la gpurun_out/r2i_prof_spec.ncu-rep
