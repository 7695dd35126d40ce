"""Ahead-of-time compilation (no GPU): the planner walk + qvm_spec_precompile fill the disk cache of cubins."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, os, sys
sys.path[:0] = [%(repo)r, os.path.join(%(repo)r, "reference-qvm_b200")]
import referenceqvm
from referenceqvm import _native as nat, precompile
from workloads import random_circuit_spec
out = {"dir": nat.jit_cache_dir()}
for gpus in (1, 2):
    spec = random_circuit_spec(14, 8, 5)
    out[str(gpus)] = precompile.precompile_spec(spec, 14, gpus)
out["counters"] = nat.jit_counters()
out["again"] = precompile.precompile_spec(random_circuit_spec(14, 8, 5), 14, 1)
out["counters_again"] = nat.jit_counters()
print(json.dumps(out))
"""


def test_precompile_fills_the_disk_cache(tmp_path):
    env = dict(os.environ, QVM_B200_CACHE_DIR=str(tmp_path / "cache"))
    proc = subprocess.run([sys.executable, "-c", SCRIPT % {"repo": REPO}], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=900)
    assert proc.returncode == 0, proc.stderr[-3000:]
    out = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["dir"] == str(tmp_path / "cache")
    cubins = [f for f in os.listdir(out["dir"]) if f.endswith(".cubin")]
    assert out["1"]["groups"] >= 2 and out["1"]["compiled"] >= 1
    assert out["2"]["groups"] >= 2 and out["2"]["compiled"] >= 1       # the sharded scheduler was walked too
    assert len(cubins) == out["counters"]["disk_stores"] >= out["1"]["compiled"]
    # a second walk finds every cubin on disk: nothing is compiled again
    assert out["counters_again"]["disk_stores"] == out["counters"]["disk_stores"]
    assert out["again"]["compiled"] == out["1"]["compiled"]
