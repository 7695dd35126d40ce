#!/bin/bash
# round 2, fourth GPU call (1 GPU): GPU test suite, then compute-sanitizer memcheck over the small-n kernel tests
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2d_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.txt
tail -5 gpurun_out/r2d_pytest_gpu.txt
# memcheck: relabelling rounds (interpreted + specialised), exchange kernels, push exchange + fused pass (in-process
# ranks), measure_hist, batched measure, the new scan / gather / expectation kernels -- small states only
QVM_B200_JIT_THREADS=2 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest -q -p no:cacheprovider -m gpu \
    "tests/test_gpu_kernels.py::test_relabelling_pass_on_device" \
    "tests/test_gpu_kernels.py::test_exchange_half_between_two_shards_on_one_device" \
    "tests/test_gpu_kernels.py::test_exchange_block_two_global_bits_on_one_device" \
    "tests/test_gpu_kernels.py::test_measure_hist_kernel_against_numpy" \
    "tests/test_gpu_kernels.py::test_regtile_mixed_ops_one_group" \
    "tests/test_gpu_kernels.py::test_sampling_matches_inverse_cdf" \
    "tests/test_gpu_api.py::test_batched_run_with_jumps_matches_the_per_trial_distribution" \
    "tests/test_gpu_api.py::test_expectation_of_operator_programs" \
    "tests/test_gpu_api.py::test_compiled_program_replay_matches_first_run_and_the_oracle" \
    "tests/test_gpu_sharded_inproc.py::test_sharded_engine_on_one_gpu[2-push]" \
    "tests/test_gpu_sharded_inproc.py::test_sharded_density_qvm_on_one_gpu[2]" \
    > gpurun_out/r2d_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2d_memcheck.txt
tail -15 gpurun_out/r2d_memcheck.txt
