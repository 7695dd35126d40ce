"""bench.py's reference arm (the one arm that runs without a GPU): one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-qubits", "12", "--ref-depth", "2"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "gate-apps/s" and line["unit"] == "gate-apps/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    staged = os.path.isdir(os.path.join(REPO, "oracle", "_ref", "referenceqvm"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and "workload" in line["config"]
