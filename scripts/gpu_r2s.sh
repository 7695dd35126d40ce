#!/bin/bash
# round 2, nineteenth GPU call (1 GPU): ncu on the final tree -- launch list of the default workload (33 qubits) and
# one --set full capture of the dominant kernel at 28 qubits (each only after the same command exited 0 without ncu)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD33="python bench.py --no-e2e --no-cpu-baseline --no-configs --steps 2 --warmup 3"
$CMD33 > gpurun_out/r2s_plain_q33.json 2> gpurun_out/r2s_plain_q33.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches_q33.csv $CMD33 > gpurun_out/r2s_ncu_launch.log 2>&1
tail -2 gpurun_out/r2s_ncu_launch.log
CMD28="python bench.py --qubits 28 --depth 40 --no-e2e --no-cpu-baseline --no-configs --steps 2 --warmup 3"
$CMD28 > gpurun_out/r2s_plain_q28.json 2> gpurun_out/r2s_plain_q28.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spec_kernel --launch-skip 40 -c 2 -f -o gpurun_out/r2s_prof_spec $CMD28 > gpurun_out/r2s_ncu_full.log 2>&1
tail -3 gpurun_out/r2s_ncu_full.log
ls -la gpurun_out/r2s_prof_spec.ncu-rep
