"""Full-size checks (BASELINE.json configs) through size-independent properties: the oracle cannot
run at these sizes, so parity rests on analytic results (QFT|0> is uniform, U U^dagger = 1,
norm = 1) and on measurement statistics."""
import numpy as np
import pytest

from _programs import spec_program
from oracle import qvm_oracle as orc
from referenceqvm import _engine, gates as G
from referenceqvm.api import QVMConnection

pytestmark = pytest.mark.gpu


def inverse_spec(spec):
    out = []
    for name, params, qubits in reversed(spec):
        if name in ("RZ", "CPHASE"):
            out.append((name, [-params[0]], qubits))
        else:
            out.append((name, params, qubits))      # H, CNOT, SWAP are self-inverse
    return out


def test_qft28_uniform_then_measure_all():
    """Config 3: 28-qubit QFT (420 gates) on |0...0> -> uniform 2^-14, then MEASURE every qubit."""
    n = 28
    eng = _engine.Engine(n)
    eng.reset()
    for name, params, qubits in orc.qft_spec(n):
        m = G.gate_matrix[name]
        eng.queue_matrix(m(*params) if params else m, qubits)
    eng.flush()
    assert abs(eng.norm2() - 1.0) < 1e-10
    amp = 2.0 ** -14
    for off in (0, 12345678, (1 << n) - 4096):
        chunk = eng.amplitudes(off, 4096)
        assert np.max(np.abs(chunk - amp)) < 1e-12
    rs = np.random.RandomState(3)
    outcomes = []
    for q in range(n):
        outcome, p0 = eng.measure(q, rs.random_sample())
        assert abs(p0 - 0.5) < 1e-9            # every conditional is 1/2 for the uniform state
        outcomes.append(outcome)
    index = sum(b << q for q, b in enumerate(outcomes))
    assert abs(eng.norm2() - 1.0) < 1e-10
    assert abs(abs(eng.amplitudes(index, 1)[0]) - 1.0) < 1e-10
    stats = eng.stats()
    eng.close()
    assert stats["hbm_passes"] < 60             # fusion: far fewer sweeps than 420 gates


def test_random_circuit_30q_forward_backward():
    """U then U^dagger on 30 qubits (16 GiB state) returns |0...0>; norm stays 1 in between."""
    n, depth = 30, 6
    spec = orc.random_circuit_spec(n, depth, 1234)
    eng = _engine.Engine(n)
    eng.reset()
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        eng.queue_matrix(m(*params) if params else m, qubits)
    eng.flush()
    assert abs(eng.norm2() - 1.0) < 1e-10
    for name, params, qubits in inverse_spec(spec):
        m = G.gate_matrix[name]
        eng.queue_matrix(m(*params) if params else m, qubits)
    eng.flush()
    head = eng.amplitudes(0, 1024)
    assert abs(head[0] - 1.0) < 1e-11 and np.max(np.abs(head[1:])) < 1e-11
    assert abs(eng.norm2() - 1.0) < 1e-10
    eng.close()


def test_density_14q_trace_and_purity_decay():
    """Config 2 shape at full size (rho = 2^28 entries = 4 GiB): trace stays 1, diag sums to 1."""
    from referenceqvm.qvm_density import INFINITY, NoiseModel
    n = 14
    spec = [("H", [], [q]) for q in range(n)] + orc.random_circuit_spec(n, 1, 1234)
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    qvm = QVMConnection(type_trans="density", noise_model=nm)
    rho = qvm.density(spec_program(spec))
    diag = rho.diagonal()
    assert abs(np.sum(diag) - 1.0) < 1e-10 and np.max(np.abs(diag.imag)) < 1e-14
    np.random.seed(0)
    res = np.asarray(qvm.run_and_measure(spec_program(spec), trials=2000))
    assert res.shape == (2000, n)
