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


SHARE_SCRIPT = r"""
import json, os, sys
sys.path[:0] = [%(repo)r, os.path.join(%(repo)r, "reference-qvm_b200")]
import referenceqvm
from referenceqvm import _native as nat, precompile
from workloads import random_circuit_spec
spec = random_circuit_spec(14, 8, 5)
offset = int(sys.argv[1])
nat.jit_share(2, offset)
nat.jit_defer(1)
res = precompile.precompile_spec(spec, 14, 1)
nat.jit_share(1, 0)
nat.jit_defer(0)
print(json.dumps({"res": res, "counters": nat.jit_counters()}))
"""


def test_two_processes_split_a_programs_structures(tmp_path):
    """qvm_jit_share: each of two processes walking the same program compiles every second structure into the
    shared disk cache; together they cover all of them, nothing is compiled twice."""
    env = dict(os.environ, QVM_B200_CACHE_DIR=str(tmp_path / "cache"))
    outs = []
    for offset in (0, 1):
        proc = subprocess.run([sys.executable, "-c", SHARE_SCRIPT % {"repo": REPO}, str(offset)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        assert proc.returncode == 0, proc.stderr[-3000:]
        outs.append(json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1]))
    groups = outs[0]["res"]["compiled"]
    assert groups >= 2 and outs[1]["res"]["compiled"] == groups
    sent = [o["counters"]["nvrtc_compiles"] for o in outs]
    assert sent[0] == (groups + 1) // 2 and sent[1] == groups // 2, (sent, groups)
    cubins = [f for f in os.listdir(str(tmp_path / "cache")) if f.endswith(".cubin")]
    assert len(cubins) == groups == sum(o["counters"]["disk_stores"] for o in outs)
