#!/bin/bash
# round 2 (1 GPU): is the one failure of r2w (in-process ranks, 8 ranks, in-place exchange: a MEASURE outcome) repeatable?
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for i in 1 2 3 4; do
timeout 600 python -m pytest tests/test_gpu_sharded_inproc.py -m gpu -q -p no:cacheprovider -k "test_sharded_engine_on_one_gpu" > gpurun_out/r2x_inproc_$i.txt 2>&1; echo "run $i rc=$?" >> gpurun_out/r2x_inproc_$i.txt
tail -3 gpurun_out/r2x_inproc_$i.txt
done
QVM_B200_ROUND_SEARCH=0 timeout 600 python -m pytest tests/test_gpu_sharded_inproc.py -m gpu -q -p no:cacheprovider -k "test_sharded_engine_on_one_gpu" > gpurun_out/r2x_inproc_norounds.txt 2>&1; echo "rc=$?" >> gpurun_out/r2x_inproc_norounds.txt
tail -3 gpurun_out/r2x_inproc_norounds.txt
