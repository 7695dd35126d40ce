"""Opcode histogram of the register-tile encoder on the benchmark circuit (host only: runs the
product encoder through the test emulator's shared object on a small state with the same tiling)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "reference-qvm_b200"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import referenceqvm  # noqa: E402,F401
import _emu  # noqa: E402
from workloads import random_circuit_spec  # noqa: E402
from referenceqvm import _engine, _planner, gates as G  # noqa: E402

NAMES = {0: "X_RELABEL", 1: "H", 6: "S1", 11: "XS", 16: "DS", 17: "DST", 18: "D1", 23: "DIAG_G", 24: "M1", 29: "M2",
         31: "T"}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    reg_bits = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    spec = random_circuit_spec(n, depth, 1234)
    ops = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        ops.append(_engine.compile_gate(m(*params) if params else m, qubits))
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    hist = np.zeros(32, dtype=np.int64)
    groups = _planner.plan(ops, n, tile, 5)
    rounds = 0
    layers = 0
    dirty = 0
    for g in groups:
        psi, info = _emu.apply_group(psi, n, g.tile, g.ops, reg_bits=reg_bits)
        hist += info["opc_hist"]
        rounds += info["rounds"]
        layers += info["layers"]
        dirty += info["dirty_rounds"]
    print("n=%d depth=%d tile=%d: %d gates, %d groups, %d rounds (%d with a thread scalar), %d layers" % (n, depth, tile, len(ops), len(groups), rounds, dirty, layers))
    agg = {}
    for opc in range(32):
        base = max(k for k in NAMES if k <= opc)
        agg[NAMES[base]] = agg.get(NAMES[base], 0) + int(hist[opc])
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        if v:
            print("  %-10s %5d  (%.1f per group)" % (k, v, v / len(groups)))


if __name__ == "__main__":
    main()
