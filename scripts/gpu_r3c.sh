#!/bin/bash
# round 2, last call (1 GPU): the GPU suite on the final tree
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r3c_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r3c_pytest_gpu.txt
tail -4 gpurun_out/r3c_pytest_gpu.txt
