"""First call of a never-seen program on a large state: compile-ahead + paced launches."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("patient", ["1", "0"])
def test_cold_first_call_is_right_whatever_kernels_it_got(tmp_path, patient):
    env = dict(os.environ)
    env["QVM_B200_CACHE_DIR"] = str(tmp_path)            # nothing on disk
    env["QVM_B200_PATIENT"] = patient
    out = subprocess.run([sys.executable, os.path.join(HERE, "_cold_call_worker.py"), "30", "12"], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("COLD ")][-1]
    r = json.loads(line[5:])
    print(r)
    assert r["all_zero_bits"] and r["amp0_minus_1"] < 1e-10 and r["max_rest"] < 1e-10 and abs(r["norm2_minus_1"]) < 1e-10, r
    assert r["second_amp0_minus_1"] < 1e-10, r
    assert r["submitted"] >= r["passes"] // 2, r           # the walk ahead of the run sent the structures to NVRTC
    assert r["second_call_specialised_share"] >= 0.9, r
    if patient == "0":
        assert r["paced_waits"] == 0, r
