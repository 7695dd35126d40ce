"""Synthetic workload generators shared by bench.py, the tests and the dev scripts.

These only *describe* circuits -- [(gate name, params, qubits)] -- and contain no simulation
arithmetic, so they are neither part of the oracle nor of the product path.
"""
import numpy as np


def random_circuit_spec(num_qubits, depth, seed):
    """The H/RZ/CNOT/CPHASE random circuit the headline metric is quoted on (SURVEY.md 8d /
    BASELINE.md section 3): per layer and per qubit q draw k in {0,1,2,3}: H(q), RZ(theta, q),
    CNOT(q, b) or CPHASE(theta, q, b) with a uniformly random partner b != q."""
    rng = np.random.RandomState(seed)
    spec = []
    for _ in range(depth):
        for q in range(num_qubits):
            kind = rng.randint(4)
            if kind == 0:
                spec.append(("H", [], [q]))
            elif kind == 1:
                spec.append(("RZ", [float(rng.uniform(0, 2 * np.pi))], [q]))
            else:
                b = int(rng.randint(num_qubits - 1))
                b += b >= q
                if kind == 2:
                    spec.append(("CNOT", [], [q, b]))
                else:
                    spec.append(("CPHASE", [float(rng.uniform(0, 2 * np.pi))], [q, b]))
    return spec


def qft_spec(num_qubits):
    """QFT in the gate ordering of the reference's QFT-8 fixture
    (``referenceqvm/tests/test_data/data_containers.py:3-44``), generalised to n qubits."""
    spec = []
    n = num_qubits
    for q in range(n - 1, -1, -1):
        for j in range(n - 1, q, -1):
            spec.append(("CPHASE", [float(np.pi / 2 ** (j - q))], [q, j]))
        spec.append(("H", [], [q]))
    for q in range(n // 2):
        spec.append(("SWAP", [], [q, n - 1 - q]))
    return spec


def density_circuit_spec(num_qubits, depth, seed):
    """BASELINE config 2: H on every qubit, then the random generator at ``depth``; run
    under NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)."""
    return [("H", [], [q]) for q in range(num_qubits)] + random_circuit_spec(num_qubits, depth, seed)


def inverse_spec(spec):
    """The adjoint circuit of a spec made of H / RZ / CNOT / CPHASE / SWAP (bench.py's U^dagger check)."""
    out = []
    for name, params, qubits in reversed(spec):
        if name in ("H", "CNOT", "SWAP", "X", "Z", "CZ"):
            out.append((name, list(params), list(qubits)))
        elif name in ("RZ", "CPHASE", "RX", "RY", "PHASE"):
            out.append((name, [-float(params[0])], list(qubits)))
        else:
            raise ValueError("no inverse rule for %s" % name)
    return out
