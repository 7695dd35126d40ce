"""Gate tables == the reference's (dumped by tests/golden/make_golden.py from reference gates.py)."""
import os

import numpy as np

from _programs import GOLDEN
from referenceqvm import gates


def _resolve(key, g):
    if key.startswith("noise:"):
        name, i = key[6:].split("@")
        return np.stack([np.asarray(k, dtype=complex) for k in gates.noise_gates[name](g["noise_p"][int(i)])])
    if key == "BARENCO":
        return gates.BARENCO(0.3, 0.7, -1.1)
    if "@" in key:
        name, i = key.split("@")
        return gates.gate_matrix[name](g["theta"][int(i)])
    if key in ("P0", "P1"):
        return gates.utility_gates[key]
    return gates.gate_matrix[key]


def test_gate_tables_bit_exact():
    g = np.load(os.path.join(GOLDEN, "gates.npz"))
    checked = 0
    for key in g.files:
        if key in ("theta", "noise_p"):
            continue
        assert np.array_equal(np.asarray(_resolve(key, g), dtype=complex), g[key]), key
        checked += 1
    assert checked >= 90


def test_gate_set_names():
    expected = set("I X Y Z H S T PHASE RX RY RZ CNOT CCNOT CPHASE00 CPHASE01 CPHASE10 CPHASE SWAP CSWAP "
                   "ISWAP PSWAP BARENCO CZ".split())
    assert set(gates.gate_matrix) == expected
    assert set(gates.stabilizer_gate_matrix) == set("I X Y Z H CNOT S CZ".split())
    assert set(gates.noise_gates) == {"relaxation", "dephasing", "depolarizing", "phase_flip", "bit_flip",
                                      "bitphase_flip"}


def test_kraus_sets_are_trace_preserving():
    # reference tests/test_density.py:135-180
    for name, f in gates.noise_gates.items():
        total = sum(np.conj(k.T).dot(k) for k in f(0.75))
        assert np.allclose(total, np.eye(2)), name


def test_parametric_sweeps():
    # reference tests/test_gates.py:4-42 (120-point sweeps)
    for phi in np.linspace(-2 * np.pi, 2 * np.pi, 120):
        assert np.allclose(gates.PHASE(phi), np.diag([1, np.exp(1j * phi)]))
        c, s = np.cos(phi / 2), np.sin(phi / 2)
        assert np.allclose(gates.RX(phi), [[c, -1j * s], [-1j * s, c]])
        assert np.allclose(gates.RY(phi), [[c, -s], [s, c]])
        assert np.allclose(gates.RZ(phi), np.diag([np.exp(-0.5j * phi), np.exp(0.5j * phi)]))
