#!/bin/bash
# round 2, fifth GPU call (1 GPU): TMA-pipelined specialised kernels -- parity subset, then A/B at 30 qubits; full suite
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
QVM_B200_SPEC_PIPE=2 timeout 900 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_kernels.py \
   "tests/test_gpu_sharded_inproc.py::test_sharded_engine_on_one_gpu" tests/test_gpu_scale.py \
   > gpurun_out/r2e_pytest_tma.txt 2>&1; echo "pytest tma rc=$?" >> gpurun_out/r2e_pytest_tma.txt
tail -5 gpurun_out/r2e_pytest_tma.txt
for mode in 0 2; do
  QVM_B200_SPEC_PIPE=$mode timeout 600 python bench.py --qubits 30 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs \
     > gpurun_out/r2e_ab_q30_pipe$mode.json 2> gpurun_out/r2e_ab_q30_pipe$mode.err; echo "ab pipe$mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_ab_q30_pipe$mode.json").read().strip().splitlines()[-1])
    print("pipe$mode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "ms/launch", round(d["roofline"]["avg_launch_ms"],3), "check", d["check"]["ok"], d["clocks"])
except Exception as e:
    print("pipe$mode failed", e)
PY
done
QVM_B200_SPEC_PIPE=2 QVM_B200_TILE_BITS=12 timeout 600 python bench.py --qubits 30 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs \
     > gpurun_out/r2e_ab_q30_pipe2_t12.json 2> gpurun_out/r2e_ab_q30_pipe2_t12.err; echo "ab pipe2 t12 rc=$?"
QVM_B200_SPEC_PIPE=2 QVM_B200_TILE_BITS=10 timeout 600 python bench.py --qubits 30 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs \
     > gpurun_out/r2e_ab_q30_pipe2_t10.json 2> gpurun_out/r2e_ab_q30_pipe2_t10.err; echo "ab pipe2 t10 rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2e_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.txt
tail -5 gpurun_out/r2e_pytest_gpu.txt
