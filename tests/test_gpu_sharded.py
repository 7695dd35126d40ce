"""Sharded product path on real GPUs (one process per GPU, NCCL): needs >= 2 devices, so it is
skipped on a single-GPU box; run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`."""
import os
import subprocess
import sys

import pytest

from test_sharded_gloo import free_port

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cuda_path_matches_oracle(world):
    if device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, QVM_TEST_PORT=str(port), QVM_TEST_RANK=str(rank), QVM_TEST_WORLD=str(world))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_sharded_gpu_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out[-4000:])
