#!/bin/bash
# round 2, fifteenth GPU call (1 GPU): parity of the rotated push orders (QVM_B200_PUSH_ORDER=1,2) with in-process ranks
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for mode in 1 2 0; do
QVM_B200_PUSH_ORDER=$mode timeout 900 python -m pytest tests/test_gpu_sharded_inproc.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider > gpurun_out/r2o_pytest_order$mode.txt 2>&1; echo "pytest order $mode rc=$?" >> gpurun_out/r2o_pytest_order$mode.txt
tail -3 gpurun_out/r2o_pytest_order$mode.txt
done
