"""N>1 host logic on CPU: world_size 2 and 4 gloo groups run the sharded engine (layout tracking,
chunked global<->local swaps, all-reduced MEASURE, sharded sampling) on the numpy shard backend and
diff every result against the oracle."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_engine_matches_oracle_under_gloo(world):
    port = free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, QVM_TEST_PORT=str(port), QVM_TEST_RANK=str(rank), QVM_TEST_WORLD=str(world),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_sharded_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out)
        assert "ok" in out
