"""GPU parity, API level: QVMConnection(...).wavefunction / run / run_and_measure / density /
unitary against (a) fixtures produced by the unmodified reference (tests/golden/) and (b) the
reference's own test-suite, re-expressed here test for test (file:line cited per test)."""
import json

import numpy as np
import pytest
import scipy.sparse as sps

from pyquil.gates import (CNOT, CZ, FALSE, H, HALT, I, NOP, PHASE, RX, RY, RZ, TRUE, WAIT, X, Y, Z)
from pyquil.quil import Program

from _programs import load_golden, oracle_ops, program_from_json, spec_program
from oracle import qvm_oracle as orc
from referenceqvm import gates as G
from referenceqvm.api import QVMConnection
from referenceqvm.qvm_density import INFINITY, NoiseModel
from referenceqvm.unitary_generator import apply_gate, lifted_gate, tensor_gates

pytestmark = pytest.mark.gpu
TOL = 1e-12


# ---- golden fixtures (reference outputs) -------------------------------------------------------
def test_golden_wavefunctions(qvm):
    meta, arr = load_golden("wavefunctions")
    for case in meta:
        wf, _ = qvm.wavefunction(program_from_json(case["program"]))
        assert qvm.num_qubits == case["n"]
        err = np.max(np.abs(np.asarray(wf.amplitudes) - arr[case["name"]]))
        assert err < TOL, (case["name"], err)


def test_golden_seeded_measurement(qvm):
    """Identical classical memory, collapsed state and RNG stream position for seeded programs."""
    meta, arr = load_golden("measure")
    done = 0
    for case in meta:
        key = case["key"]
        if key + ":wf" not in arr.files:
            continue
        np.random.seed(case["seed"])
        wf, mem = qvm.wavefunction(program_from_json(case["program"]))
        nxt = np.random.random()
        assert mem[:16] == arr[key + ":mem"].tolist(), key
        assert nxt == arr[key + ":next"][0], key
        assert np.max(np.abs(wf.amplitudes - arr[key + ":wf"])) < TOL, key
        done += 1
    assert done >= 25


def test_golden_run_streams_with_per_trial_path(qvm):
    meta, arr = load_golden("measure")
    progs = {c["name"]: c for c in meta}
    qvm.sampling_fast_path = False
    try:
        np.random.seed(42)
        res = qvm.run(program_from_json(progs["biased"]["program"]), [2, 3, 4], trials=50)
        assert np.array_equal(np.asarray(res), arr["run_biased@42"])
        case = progs["ram_random4"]
        np.random.seed(43)
        res = qvm.run_and_measure(program_from_json(case["program"]), case["qubits"], trials=case["trials"])
        assert np.array_equal(np.asarray(res), arr["ram_random4@43"])
    finally:
        qvm.sampling_fast_path = True


def test_golden_density():
    meta, arr = load_golden("density")
    done = 0
    for case in meta:
        if "program" not in case:
            continue
        noise = None
        if case["noise"] is not None:
            noise = NoiseModel(**{k: (INFINITY if v == "inf" else v) for k, v in case["noise"].items()})
        qvm = QVMConnection(type_trans="density", noise_model=noise)
        if "seed" in case:
            np.random.seed(case["seed"])
        rho = qvm.density(program_from_json(case["program"]))
        err = np.max(np.abs(np.asarray(rho.todense()) - arr[case["name"]]))
        assert err < TOL, (case["name"], err)
        if case["name"] + ":mem" in arr.files:
            assert qvm.classical_memory[:8].astype(int).tolist() == arr[case["name"] + ":mem"].tolist()
        done += 1
    assert done >= 17


def test_golden_apply_kraus():
    meta, arr = load_golden("density")
    rho0 = arr["kraus_rho0"]
    for case in meta:
        if "channel" not in case:
            continue
        qvm = QVMConnection(type_trans="density")
        qvm._density = sps.csc_matrix(rho0)
        qvm.num_qubits = case["n"]
        out = qvm.apply_kraus(G.noise_gates[case["channel"]](case["p"]), case["qubit"])
        assert np.max(np.abs(np.asarray(out.todense()) - arr[case["name"]])) < TOL, case["name"]
        assert np.array_equal(np.asarray(qvm._density.todense()), rho0)       # input left untouched


def test_golden_sampling_streams():
    meta, arr = load_golden("sampling")
    for case in meta:
        qvm = QVMConnection(type_trans="density")
        np.random.seed(case["seed"])
        prog = program_from_json(case["program"])
        if case["name"] == "density_run":
            res = qvm.run(prog, case["addresses"], trials=case["trials"])
            assert np.array_equal(np.asarray(res).astype(np.int8), arr["density_run:draws"])
            assert all(r.dtype == bool for r in res)
            continue
        res = np.asarray(qvm.run_and_measure(prog, trials=case["trials"]), dtype=np.int8)
        want = arr[case["name"] + ":draws"]
        assert res.shape == want.shape
        assert np.sum(np.any(res != want, axis=1)) <= 1, case["name"]      # cdf-boundary ties only


def test_golden_unitary(qvm_unitary):
    meta, arr = load_golden("unitary")
    for case in meta:
        u = qvm_unitary.unitary(program_from_json(case["program"]))
        assert u.shape == arr[case["name"]].shape
        assert np.max(np.abs(u - arr[case["name"]])) < TOL, case["name"]


def test_golden_operators():
    meta, arr = load_golden("operators")
    for case in meta:
        if "lift" in case:
            assert np.array_equal(lifted_gate(case["lift"], G.SWAP, case["n"]).toarray(), arr[case["key"]])
        elif "vec" in case:
            got = apply_gate(arr["G%d" % case["k"]], case["args"], case["n"]).dot(arr[case["vec"]])
            assert np.max(np.abs(got - arr[case["key"]])) < 1e-12 * np.max(np.abs(arr[case["key"]]))
        else:
            got = apply_gate(arr["G%d" % case["k"]], case["args"], case["n"]).toarray()
            assert np.max(np.abs(got - arr[case["key"]])) < TOL, case["key"]


# ---- reference tests/test_unitary_generator.py ----------------------------------------------------
def test_lifted_swap():                      # :14-53
    assert np.allclose(lifted_gate(0, G.SWAP, 2).toarray(), G.SWAP)
    assert np.allclose(lifted_gate(0, G.SWAP, 3).toarray(), np.kron(np.eye(2), G.SWAP))
    assert np.allclose(lifted_gate(0, G.SWAP, 4).toarray(), np.kron(np.eye(4), G.SWAP))
    assert np.allclose(lifted_gate(1, G.SWAP, 3).toarray(), np.kron(G.SWAP, np.eye(2)))
    assert np.allclose(lifted_gate(1, G.SWAP, 4).toarray(), np.kron(np.eye(2), np.kron(G.SWAP, np.eye(2))))
    assert np.allclose(lifted_gate(2, G.SWAP, 4).toarray(), np.kron(G.SWAP, np.eye(4)))
    assert np.allclose(lifted_gate(8, G.SWAP, 10).toarray(), np.kron(G.SWAP, np.eye(2 ** 8)))


def test_two_qubit_gates():                  # :56-136 argument-order convention
    P0, P1 = G.P0, G.P1
    assert np.allclose(apply_gate(G.CNOT, [1, 0], 2).toarray(), np.kron(P0, np.eye(2)) + np.kron(P1, G.X))
    assert np.allclose(apply_gate(G.CNOT, [0, 1], 2).toarray(), np.kron(np.eye(2), P0) + np.kron(G.X, P1))
    assert np.allclose(apply_gate(G.CNOT, [2, 1], 3).toarray(), np.kron(G.CNOT, np.eye(2)))
    v = np.kron(np.eye(2), G.SWAP)
    far = v.dot(np.kron(G.CNOT, np.eye(2))).dot(v)            # CNOT(2, 0) via swap conjugation
    assert np.allclose(apply_gate(G.CNOT, [2, 0], 3).toarray(), far)
    assert np.allclose(apply_gate(G.ISWAP, [0, 1], 3).toarray(), np.kron(np.eye(2), G.ISWAP))
    assert np.allclose(apply_gate(G.SWAP, [0, 2], 3).toarray(), apply_gate(G.SWAP, [2, 0], 3).toarray())


def test_single_qubit_lifting_and_tensor_gates():     # :139-213
    for n in (4, 5):
        for q in range(n):
            want = np.kron(np.eye(2 ** (n - 1 - q)), np.kron(G.H, np.eye(2 ** q)))
            assert np.allclose(apply_gate(G.H, [q], n).toarray(), want)
    t = tensor_gates(G.gate_matrix, {}, H(1), 3)
    assert np.allclose(t.toarray(), np.kron(np.eye(2), np.kron(G.H, np.eye(2))))
    t = tensor_gates(G.gate_matrix, {}, RX(0.2, 3), 4)
    assert np.allclose(t.toarray(), np.kron(G.RX(0.2), np.eye(8)))
    t = tensor_gates(G.gate_matrix, {}, CNOT(0, 1), 2)
    assert np.allclose(t.toarray(), np.kron(np.eye(2), G.P0) + np.kron(G.X, G.P1))


# ---- reference tests/test_wavefunction.py, test_system.py, test_arbitrary_state_prep.py ------------
def test_belltest(qvm):                      # test_wavefunction.py:27-36
    bellout, _ = qvm.wavefunction(Program().inst([H(0), CNOT(0, 1)]))
    bell = np.zeros(4)
    bell[0] = bell[-1] = 1.0 / np.sqrt(2)
    assert np.allclose(bellout.amplitudes, bell)


def test_occupation_basis(qvm):              # :38-44
    state = np.zeros(16)
    state[3] = 1.0
    wf, _ = qvm.wavefunction(Program().inst([X(0), X(1), I(2), I(3)]))
    assert np.allclose(wf.amplitudes, state)


def test_exp_circuit_norm(qvm):              # :46-68 (shim-dependent: pyquil.paulis.exponentiate)
    from pyquil.paulis import PauliTerm, exponentiate
    op = PauliTerm("X", 1, -0.25) * PauliTerm("Y", 2)
    op += PauliTerm("Y", 1, 0.25) * PauliTerm("Y", 2)
    op += PauliTerm("Y", 1, 0.25) * PauliTerm("X", 2)
    op += PauliTerm("X", 1, 0.25) * PauliTerm("X", 2)
    op += PauliTerm("I", 0, 1.0)
    prog = Program()
    for term in op.terms:
        prog += exponentiate(term)
    wf, _ = qvm.wavefunction(prog)
    a = np.reshape(wf.amplitudes, -1)
    assert np.allclose(a.dot(np.conj(a).T), 1.0)


QAOA2 = [RY(np.pi / 2, 0), RX(np.pi, 0), RY(np.pi / 2, 1), RX(np.pi, 1), CNOT(0, 1), RX(-np.pi / 2, 1),
         RY(4.71572463191, 1), RX(np.pi / 2, 1), CNOT(0, 1), RX(-2 * 2.74973750579, 0), RX(-2 * 2.74973750579, 1)]
QAOA2_WF = [0.00167784 + 1.00210180e-05 * 1j, 0.50000000 - 4.99997185e-01 * 1j,
            0.50000000 - 4.99997185e-01 * 1j, 0.00167784 + 1.00210180e-05 * 1j]


def test_qaoa_circuit(qvm):                  # :71-81
    wf, _ = qvm.wavefunction(Program().inst(QAOA2))
    assert np.allclose(wf.amplitudes, QAOA2_WF)


def test_against_cloud_fixtures(qvm):        # test_system.py:7-23 + data_containers QFT-8
    p = Program(H(0))
    assert len(qvm.run(p, classical_addresses=[0], trials=10)) == 10
    wf, _ = qvm.wavefunction(p)
    assert np.allclose([0.70710678, 0.70710678], wf.amplitudes)
    wf, _ = qvm.wavefunction(spec_program(orc.qft_spec(8)))
    assert np.allclose(wf.amplitudes, np.full(256, 0.0625))


# ---- reference tests/test_measurement.py -------------------------------------------------------------
def test_measurement(qvm):                   # :6-17
    p = Program().inst(X(0))
    results = qvm.run_and_measure(p)
    assert len(results) == 1 and len(results[0]) == 0
    results = qvm.run_and_measure(p, qubits=[0, 1, 2, 3])
    assert len(results) == 1 and len(results[0]) == 4
    assert np.allclose(results, [1, 0, 0, 0])
    with pytest.raises(TypeError):
        qvm.run_and_measure(p, qubits=(0, 1))


def test_many_sample_measurements(qvm):      # :20-26
    prog = Program().inst(H(0)).measure(0, [0]).measure(0, [1])
    for fast in (True, False):
        qvm.sampling_fast_path = fast
        samples = 1000 if fast else 200
        result = qvm.run(prog, [0, 1], trials=samples)
        assert len(result) == samples and all(x[0] == x[1] for x in result)
        bias = sum(x[0] for x in result) / float(samples)
        assert np.isclose(0.5, bias, atol=0.12 if not fast else 0.05, rtol=0.05)
    qvm.sampling_fast_path = True


def test_biased_coin(qvm):                   # :29-34
    prog = Program().inst(RX(np.pi / 3, 0))
    results = qvm.run_and_measure(prog, qubits=[0], trials=1000)
    assert np.isclose(sum(x[0] for x in results) / 1000.0, 0.25, atol=0.05, rtol=0.05)


def test_run_and_measure_chi_squared_against_reference_distribution(qvm):
    spec = orc.random_circuit_spec(5, 6, 77)
    psi, _ = orc.fast_run_wavefunction(oracle_ops(spec_program(spec)), 5)
    probs = psi.real ** 2 + psi.imag ** 2
    shots = 100000
    np.random.seed(5)
    res = np.asarray(qvm.run_and_measure(spec_program(spec), qubits=[0, 1, 2, 3, 4, 7], trials=shots))
    assert res.shape == (shots, 6) and not res[:, 5].any()
    idx = res[:, :5].dot(1 << np.arange(5))
    counts = np.bincount(idx, minlength=32)
    keep = probs > 1e-12
    chi2 = np.sum((counts[keep] - shots * probs[keep]) ** 2 / (shots * probs[keep]))
    assert counts[~keep].sum() == 0 and chi2 < 75.0


# ---- reference tests/test_controlflow.py, test_qvm_funcs.py ---------------------------------------------
def test_if_then(qvm):                       # :6-21
    def build():
        main = Program().inst(X(0))
        main.if_then(0, Program().inst(X(0)), Program().inst())
        return main
    assert qvm.run_and_measure(Program().inst(TRUE(0)) + build(), [0])[0][0] == 0
    assert qvm.run_and_measure(Program().inst(FALSE(0)) + build(), [0])[0][0] == 1


def test_while(qvm):                         # :24-37
    loop = Program(TRUE([2])).while_do(2, Program(X(0), H(0)).measure(0, 2))
    _, cregs = qvm.wavefunction(loop, [2])
    assert cregs[0] == False  # noqa: E712


def test_halt(qvm):                          # :40-50
    prog = Program(X(0))
    prog.inst(HALT)
    prog.inst(X(0))
    assert qvm.run_and_measure(prog, [0])[0][0] == 1
    assert qvm.run_and_measure(Program(X(0)).inst(X(0)), [0])[0][0] == 0


def test_unsupported_instructions(qvm):      # :53-68
    for bad in (NOP, WAIT):
        prog = Program(bad)
        with pytest.raises(TypeError):
            qvm.wavefunction(prog)
        with pytest.raises(TypeError):
            qvm.run(prog)
        with pytest.raises(TypeError):
            qvm.run_and_measure(prog)


def test_identify_bits_through_api(qvm):     # test_qvm_funcs.py:5-31
    p = Program()
    for i in range(5):
        p.inst(X(i))
    for i in range(670):
        p.inst(TRUE(i))
    qvm.wavefunction(p)
    qvm.run(p)
    qvm.run_and_measure(p)
    assert qvm.num_qubits == 5 and len(qvm.classical_memory) == 670
    p = Program()
    for i in range(670):
        p.inst(TRUE(i))
    wf, mem = qvm.wavefunction(p)
    qvm.run(p)
    qvm.run_and_measure(p)
    assert qvm.num_qubits == 0 and len(qvm.classical_memory) == 670
    assert np.allclose(wf.amplitudes, [1.0]) and sum(mem) == 670


def test_classical_address_checks(qvm):      # qvm_wavefunction.py:340-347
    p = Program(X(0)).measure(0, [3])
    _, mem = qvm.wavefunction(p, [3, 0])
    assert mem == [1, 0]
    with pytest.raises(RuntimeError):
        qvm.wavefunction(p, [512])
    with pytest.raises(AssertionError):
        qvm.wavefunction(p, [1, 1])
    with pytest.raises(ValueError):
        qvm.wavefunction(Program().inst(("NOSUCHGATE", 0)))


# ---- reference tests/test_density.py -------------------------------------------------------------------
def random_1q_density():
    state = np.random.random(2) + 1j * np.random.random()
    state /= np.sqrt(np.conj(state).T.dot(state))
    state = state.reshape((-1, 1))
    return state.dot(np.conj(state).T)


def test_qaoa_density():                     # :36-51
    wf_true = np.reshape(np.array(QAOA2_WF), (4, 1))
    rho_true = np.dot(wf_true, np.conj(wf_true).T)
    rho = QVMConnection(type_trans="density").density(Program().inst(QAOA2))
    assert np.isclose(rho_true, rho.todense()).all()


@pytest.mark.parametrize("chan,expect", [
    ("bit_flip", lambda r, p: (1 - p) * r + p * G.X.dot(r).dot(G.X)),
    ("phase_flip", lambda r, p: (1 - p) * r + p * G.Z.dot(r).dot(G.Z)),
    ("bitphase_flip", lambda r, p: (1 - p) * r + p * G.Y.dot(r).dot(G.Y)),
    ("relaxation", lambda r, p: np.array([[r[0, 0] + r[1, 1] * p, np.sqrt(1 - p) * r[0, 1]],
                                          [np.sqrt(1 - p) * r[1, 0], (1 - p) * r[1, 1]]])),
    ("dephasing", lambda r, p: np.array([[r[0, 0], (1 - p) * r[0, 1]], [(1 - p) * r[1, 0], r[1, 1]]])),
    ("depolarizing", lambda r, p: (1 - p) * r + (p / 3) * (G.X.dot(r).dot(G.X) + G.Y.dot(r).dot(G.Y) +
                                                          G.Z.dot(r).dot(G.Z))),
])
def test_kraus_application(chan, expect):    # :183-257
    qvm = QVMConnection(type_trans="density")
    rho = random_1q_density()
    qvm._density = sps.csc_matrix(rho)
    qvm.num_qubits = 1
    p = 0.372
    out = qvm.apply_kraus(G.noise_gates[chan](p), 0)
    assert np.allclose(out.todense(), expect(rho, p))


def test_kraus_compound_T1T2_application():  # :260-272
    qvm = QVMConnection(type_trans="density")
    rho = random_1q_density()
    qvm._density = sps.csc_matrix(rho)
    qvm.num_qubits = 1
    p1, p2 = 0.372, 0.45
    want = np.array([[rho[0, 0] + rho[1, 1] * p1, (1 - p2) * np.sqrt(1 - p1) * rho[0, 1]],
                     [(1 - p2) * np.sqrt(1 - p1) * rho[1, 0], (1 - p1) * rho[1, 1]]])
    qvm._density = qvm.apply_kraus(G.noise_gates["relaxation"](p1), 0)
    qvm._density = qvm.apply_kraus(G.noise_gates["dephasing"](p2), 0)
    assert np.allclose(qvm._density.todense(), want)


@pytest.mark.parametrize("kw", [dict(T2=INFINITY, ro_fidelity=1.0), dict(T1=INFINITY, ro_fidelity=1.0),
                                dict(ro_fidelity=1.0)])
def test_kraus_through_qvm(kw):              # :275-318
    noise = NoiseModel(**kw)
    qvm = QVMConnection(type_trans="density", noise_model=noise)
    rho = random_1q_density()
    qvm._density = sps.csc_matrix(rho)
    qvm.num_qubits = 1
    qvm.load_program(Program().inst(I(0)))
    qvm.kernel()
    p1 = 0.0 if noise.T1 == INFINITY else 1 - np.exp(-noise.gate_time_1q / noise.T1)
    p2 = 0.0 if noise.T2 == INFINITY else 1 - np.exp(-noise.gate_time_1q / noise.T2)
    want = np.array([[rho[0, 0] + p1 * rho[1, 1], np.sqrt(1 - p1) * (1 - p2) * rho[0, 1]],
                     [np.sqrt(1 - p1) * (1 - p2) * rho[1, 0], (1 - p1) * rho[1, 1]]])
    assert np.allclose(want, qvm._density.todense())


# ---- reference tests/test_sampling.py ---------------------------------------------------------------------
def test_density_sampling_statistics():      # :12-54
    qvm = QVMConnection(type_trans="density")
    res = qvm.run_and_measure(Program().inst(H(0)), trials=10000)
    assert np.isclose(sum(x[0] for x in res) / 10000.0, 0.5, atol=0.05, rtol=0.05)
    res = qvm.run_and_measure(Program().inst(H(0), CNOT(0, 1)), trials=100000)
    assert all(rr[0] == rr[1] for rr in res)
    assert np.isclose(sum(x[0] for x in res) / 100000.0, 0.5, atol=0.05, rtol=0.05)
    res = qvm.run_and_measure(Program().inst(RX(np.pi / 3, 0)), trials=100000)
    assert np.isclose(sum(x[0] for x in res) / 100000.0, 0.25, atol=0.05, rtol=0.05)
    prog = Program().inst(H(0)).measure(0, [0]).measure(0, [1])
    result = qvm.run(prog, [0, 1], trials=300)
    assert all(x[0] == x[1] for x in result)
    assert np.isclose(0.5, sum(x[0] for x in result) / 300.0, atol=0.1, rtol=0.05)


# ---- reference tests/test_unitary.py ------------------------------------------------------------------------
def test_unitary_qvm(qvm_unitary):           # :8-52
    u = qvm_unitary.unitary(Program().inst([H(0), H(1), H(0)]))
    assert np.allclose(u, np.kron(G.H, np.eye(2)))
    u = qvm_unitary.unitary(Program().inst([H(0), X(1), Y(2), Z(3)]))
    assert np.allclose(u, np.kron(G.Z, np.kron(G.Y, np.kron(G.X, G.H))))
    u = qvm_unitary.unitary(Program().inst([X(2), CNOT(2, 1), CNOT(1, 0)]))
    want = np.kron(np.eye(2), G.CNOT).dot(np.kron(G.CNOT, np.eye(2))).dot(np.kron(G.X, np.eye(4)))
    assert np.allclose(u, want)
    assert np.allclose(qvm_unitary.unitary(Program()), np.eye(1))
    u = qvm_unitary.unitary(Program().inst(QAOA2))
    assert np.allclose(u.dot([1, 0, 0, 0]), QAOA2_WF)


def test_unitary_errors(qvm_unitary):        # :55-70
    prog = Program().inst([H(0), H(1)])
    prog.measure(0, [0])
    with pytest.raises(TypeError):
        qvm_unitary.unitary(prog)
    prog = Program()
    prog.defgate("hello", np.array([[0, 1], [1, 0]]))
    prog.inst(("hello2", 0))
    with pytest.raises(TypeError):
        qvm_unitary.unitary(prog)


# ---- differential, at sizes the reference's algorithm cannot reach: closed-form oracle ---------------------
@pytest.mark.parametrize("n,depth,seed", [(16, 10, 1234), (20, 6, 1), (22, 4, 2)])
def test_random_circuit_vs_closed_form_oracle(qvm, n, depth, seed):
    spec = orc.random_circuit_spec(n, depth, seed)
    prog = spec_program(spec)
    wf, _ = qvm.wavefunction(prog)
    want, _ = orc.fast_run_wavefunction(oracle_ops(prog), n)
    assert np.max(np.abs(wf.amplitudes - want)) < TOL


def test_qft_on_product_state_vs_oracle(qvm):
    n = 16
    rs = np.random.RandomState(n)
    prep = [("RY", [float(rs.uniform(0, np.pi))], [q]) for q in range(n)] + \
           [("RZ", [float(rs.uniform(0, 2 * np.pi))], [q]) for q in range(n)]
    prog = spec_program(prep + orc.qft_spec(n))
    wf, _ = qvm.wavefunction(prog)
    want, _ = orc.fast_run_wavefunction(oracle_ops(prog), n)
    assert np.max(np.abs(wf.amplitudes - want)) < TOL


def test_density_config2_shape_vs_oracle():
    """Config 2 at a size the oracle finishes quickly: H layer + random circuit + T2 dephasing."""
    n = 7
    spec = [("H", [], [q]) for q in range(n)] + orc.random_circuit_spec(n, 3, 1234)
    prog = spec_program(spec)
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    rho = QVMConnection(type_trans="density", noise_model=nm).density(prog)
    vec = np.zeros(4 ** n, dtype=complex)
    vec[0] = 1.0
    for _, m, q in oracle_ops(prog):
        vec = orc.fast_density_gate(vec, m, q, n)
        p = 1.0 - np.exp(-(50e-9 if len(q) == 1 else 150e-9) / 30e-6)
        for i in range(n):
            vec = orc.fast_apply_kraus(vec, G.dephasing(p), i, n)
    got = np.asarray(rho.todense())
    assert np.max(np.abs(got - orc.vec_to_rho(vec, n))) < TOL
    assert abs(np.trace(got) - 1.0) < 1e-12


def test_swap_gates_are_relabels_with_identical_results(qvm):
    """SWAP instructions never touch the amplitudes (the QVM routes later gates, MEASURE and sampling
    through a program-qubit -> engine-qubit table); every observable result must equal the oracle's."""
    from pyquil.gates import CNOT, H, MEASURE, RY, SWAP, X
    n = 6
    prog = Program(X(0), H(1), SWAP(0, 3), CNOT(1, 4), SWAP(1, 5), RY(0.7, 2), SWAP(2, 3), SWAP(0, 5), CNOT(3, 0),
                   SWAP(4, 1))
    ops = []
    for ins in prog:
        m = G.gate_matrix[ins.name]
        ops.append(("gate", np.asarray(m(*ins.params) if ins.params else m, dtype=complex),
                    [q.index if hasattr(q, "index") else q for q in ins.qubits]))
    want, _ = orc.fast_run_wavefunction(ops, n)
    wf, _ = qvm.wavefunction(prog)
    assert np.max(np.abs(wf.amplitudes - want)) < 1e-12
    passes_with_readback = qvm._engine.stats()["hbm_passes"]
    # MEASURE after swaps: deterministic bits (X(0) travels 0 -> 3 -> 2, then CNOT(3,0) does nothing ...)
    probs = np.abs(want) ** 2
    mprog = Program().inst(prog)
    for q in range(n):
        mprog.measure(q, [q])
    np.random.seed(5)
    _, mem = qvm.wavefunction(mprog, list(range(n)))
    rng = np.random.RandomState(5)
    _, cmem = orc.fast_run_wavefunction(ops + [("measure", q, q) for q in range(n)], n, rng)
    assert mem == [cmem[q] for q in range(n)]
    # sampling fast path honours the relabelling and the caller's qubit order
    np.random.seed(6)
    order = [5, 0, 3, 2, 4, 1, 9]
    shots = np.asarray(qvm.run_and_measure(prog, order, trials=4000))
    u = np.random.RandomState(6).random_sample(4000)
    # the engine samples in ENGINE index order: compare distributions, not streams
    idx = sum(shots[:, j].astype(np.int64) << q for j, q in enumerate(order) if q < n)
    freq = np.bincount(idx, minlength=1 << n) / 4000.0
    assert np.max(np.abs(freq - probs)) < 0.05 and not shots[:, -1].any()
    assert passes_with_readback >= 1


# ---- batched trials: mid-circuit MEASURE without control flow ------------------------------
def test_run_with_midcircuit_measure_batches_trials_and_matches_the_per_trial_loop(qvm):
    """run(trials) on a program that measures in mid-circuit and keeps going: all trials execute side by
    side (one segment of the device buffer each).  With the same seed the batched path must return exactly
    the bits of the per-trial loop (the reference's algorithm, qvm_wavefunction.py:279-283)."""
    from pyquil.gates import AND, EXCHANGE, MEASURE, MOVE, NOT, OR, SWAP
    prog = Program(H(0), CNOT(0, 1), RY(0.9, 2), MEASURE(0, 0), CNOT(1, 2), H(1), MEASURE(1, 1), NOT(1),
                   RX(1.1, 3), CZ(2, 3), SWAP(0, 3), MEASURE(2, 2), MOVE(0, 5), AND(1, 2), OR(0, 2), EXCHANGE(1, 5),
                   H(0), MEASURE(0, 3), TRUE(6), FALSE(7), MEASURE(3, 4))
    addresses = [0, 1, 2, 3, 4, 5, 6, 7]
    trials = 300
    assert qvm._batchable(prog) and qvm._terminal_measure_block(prog) is None
    np.random.seed(77)
    batched = qvm.run(prog, addresses, trials=trials)
    stream_after = np.random.random()
    batched_state = np.asarray(qvm.wf).copy()
    batched_memory = qvm.classical_memory.copy()
    qvm.sampling_fast_path = False
    try:
        np.random.seed(77)
        looped = qvm.run(prog, addresses, trials=trials)
        assert np.random.random() == stream_after              # same number of draws, same order
        looped_state = np.asarray(qvm.wf).copy()
    finally:
        qvm.sampling_fast_path = True
    assert batched == looped
    assert len({tuple(r) for r in batched}) > 4                # the trials really differ
    # the last trial's wavefunction and classical memory are left behind, as in the reference
    assert np.max(np.abs(batched_state - looped_state)) < 1e-12
    assert np.array_equal(batched_memory, qvm.classical_memory)
    # more trials than one batch holds, and all memory (classical_addresses=None)
    old = qvm.BATCH_BITS
    type(qvm).BATCH_BITS = 4 + 5
    try:
        np.random.seed(5)
        chunked = qvm.run(prog, None, trials=70)
        qvm.sampling_fast_path = False
        np.random.seed(5)
        ref = qvm.run(prog, None, trials=70)
    finally:
        qvm.sampling_fast_path = True
        type(qvm).BATCH_BITS = old
    assert chunked == ref and len(chunked[0]) >= 512


def test_batched_measure_kernel_against_oracle():
    from referenceqvm import _native as nat
    n, t = 9, 5
    rs = np.random.RandomState(8)
    st = nat.DeviceState(n + t)
    segs = []
    for s in range(1 << t):
        v = rs.standard_normal(1 << n) + 1j * rs.standard_normal(1 << n)
        segs.append(v / np.linalg.norm(v))
    st.set_state(np.concatenate(segs))
    for qubit in (0, 4, 8):
        u = rs.random_sample(1 << t)
        outcomes, p0 = st.measure_batched(n, qubit, u)
        got = st.get_state().reshape(1 << t, 1 << n)
        for s in range(1 << t):
            o_want, p_want, psi = orc.fast_measure(segs[s], qubit, n, u[s])
            assert outcomes[s] == o_want and abs(p0[s] - p_want) < 1e-12
            assert np.max(np.abs(got[s] - psi)) < 1e-12
            segs[s] = psi
    st.reset_batched(n)
    got = st.get_state().reshape(1 << t, 1 << n)
    assert np.all(got[:, 0] == 1.0) and np.count_nonzero(got) == 1 << t
    st.close()


def test_runs_of_measures_are_fused_and_keep_the_seeded_stream(qvm):
    """wavefunction() on a program that ends in MEASUREs of all qubits (shuffled order, plus a repeated
    qubit and gates in between): same classical memory, same RNG stream position and same collapsed state
    as the oracle's one-by-one measurements."""
    from pyquil.gates import MEASURE
    n = 12
    spec = orc.random_circuit_spec(n, 4, 21)
    rs = np.random.RandomState(2)
    order = [int(q) for q in rs.permutation(n)]
    prog = spec_program(spec)
    ops = list(oracle_ops(prog))
    for c, q in enumerate(order):
        prog.inst(MEASURE(q, c))
        ops.append(("measure", q, c))
    prog.inst(MEASURE(order[0], 20), H(order[1]), MEASURE(order[1], 21), MEASURE(order[2], 22))
    ops += [("measure", order[0], 20), ("gate", np.asarray(G.H, dtype=complex), [order[1]]), ("measure", order[1], 21),
            ("measure", order[2], 22)]
    np.random.seed(99)
    wf, mem = qvm.wavefunction(prog, list(range(24)))
    nxt = np.random.random()
    rng = np.random.RandomState(99)
    want, cmem = orc.fast_run_wavefunction(ops, n, rng)
    assert mem == [int(cmem.get(i, 0)) for i in range(24)]
    assert nxt == rng.random_sample()
    assert np.max(np.abs(wf.amplitudes - want)) < TOL
    assert qvm._engine.stats()["hbm_passes"] < 2 * (n + 3)          # fused: far fewer than two passes per MEASURE


# ---- round-2 review items --------------------------------------------------------------------------------
def test_batched_run_with_more_than_2_pow_20_trials(qvm):
    """run() on a tiny program with a mid-circuit MEASURE and more than 2^20 trials: more batches, not an error
    (the batched MEASURE kernel takes at most 2^20 trial segments per launch)."""
    prog = Program().inst(H(0)).measure(0, [0]).inst(CNOT(0, 1)).measure(1, [1])
    trials = (1 << 20) + 1
    np.random.seed(5)
    got = np.asarray(qvm.run(prog, [0, 1], trials=trials))
    assert got.shape == (trials, 2)
    assert np.array_equal(got[:, 0], got[:, 1])                 # CNOT copies the measured bit
    assert abs(got[:, 0].mean() - 0.5) < 0.005


def test_density_returns_independent_snapshots():
    """density() hands out snapshots (the reference returns a fresh matrix per call); the `_density` attribute
    is a live view that follows the QVM and goes stale with an error when the QVM replaces its state."""
    qvm = QVMConnection(type_trans="density")
    rho_a = qvm.density(Program().inst(H(0)).inst(CNOT(0, 1)))
    a = np.array(rho_a.todense())
    rho_b = qvm.density(Program().inst(X(0)).inst(I(1)))
    assert np.allclose(np.array(rho_a.todense()), a)            # untouched by the second call
    assert abs(np.array(rho_b.todense())[1, 1] - 1.0) < 1e-15
    view = qvm._density
    assert abs(view[1, 1] - 1.0) < 1e-15
    qvm.density(Program().inst(X(0)).inst(X(1)).inst(I(2)))      # other size: the old buffer is gone
    with pytest.raises(RuntimeError):
        view.toarray()


def test_subclass_hooks_see_every_measure():
    """A subclass that overrides a hook is called around every instruction, MEASURE runs included (the
    fused joint-distribution path is only taken for the base class)."""
    from referenceqvm.qvm_wavefunction import QVM_Wavefunction

    class Counting(QVM_Wavefunction):
        calls = 0

        def post(self):
            Counting.calls += 1

    n = 12
    prog = Program()
    for q in range(n):
        prog.inst(H(q))
    for q in range(n):
        prog.measure(q, [q])
    qvm = Counting(gate_set=G.gate_matrix)
    qvm.wavefunction(prog, list(range(n)))
    assert Counting.calls == 2 * n


def test_sampling_cdf_scan_long_arrays_matches_numpy():
    """The three-phase CDF scan (arrays longer than one chunk of 4096 block sums: >= 24 qubits) against
    numpy's inverse CDF on the same probabilities."""
    from referenceqvm import _engine
    n = 24
    eng = _engine.Engine(n)
    try:
        eng.reset()
        rs = np.random.RandomState(3)
        for q in range(n):
            eng.queue_matrix(G.RY(float(rs.uniform(0.2, 2.9))), [q])
        eng.flush()
        probs = np.abs(eng.amplitudes()) ** 2
        u = rs.random_sample(20000)
        idx, total = eng.sample(u, physical=True)
        assert abs(total - 1.0) < 1e-12
        phys = np.arange(1 << n, dtype=np.int64)
        logical = np.zeros_like(phys)
        for q, p in enumerate(eng.pos_of):
            logical |= ((phys >> p) & 1) << q
        ref = np.asarray(orc.inverse_cdf_sample(probs[logical], u), dtype=np.int64)      # CDF in HBM order
        assert np.mean(idx.astype(np.int64) == ref) > 0.999
    finally:
        eng.close()


# ---- compiled programs (plan cache + CUDA graph) -------------------------------------------------------------
def test_compiled_program_replay_matches_first_run_and_the_oracle():
    """The second and later runs of a Program object replay its compiled gate prefix (one library call, a CUDA
    graph once every group has its kernel): same amplitudes as the first run, which match the oracle."""
    from referenceqvm import _compiled
    n, depth = 16, 10
    spec = orc.random_circuit_spec(n, depth, 21)
    prog = spec_program(spec)
    want, _ = orc.fast_run_wavefunction(oracle_ops(prog), n)
    qvm = QVMConnection()
    first = np.asarray(qvm.wavefunction(prog)[0].amplitudes)
    assert np.max(np.abs(first - want)) < 1e-12
    entry = _compiled.lookup(prog, qvm._plan_key())
    assert entry is not None and entry.k == len(prog) and entry.compiled.plan.launches >= 2
    qvm._engine.state.set_jit(2)                      # kernels at once: every group specialised from the next run on
    for _ in range(6):
        again = np.asarray(qvm.wavefunction(prog)[0].amplitudes)
        assert np.array_equal(again, first)
    st = qvm._engine.stats()
    assert st["graph_launches"] >= 1, st               # the launch sequence became one graph launch
    # appending to the program invalidates the cache; the new program is right too
    prog.inst(H(3)).inst(CNOT(3, 9))
    want2, _ = orc.fast_run_wavefunction(oracle_ops(prog), n)
    got2 = np.asarray(qvm.wavefunction(prog)[0].amplitudes)
    assert np.max(np.abs(got2 - want2)) < 1e-12
    # the cache can be switched off per QVM
    plain = QVMConnection()
    plain.plan_cache = False
    assert np.max(np.abs(np.asarray(plain.wavefunction(spec_program(spec))[0].amplitudes) - want)) < 1e-12


def test_compiled_prefix_with_measurements_and_swaps_keeps_the_seeded_stream():
    """Programs whose gate prefix is followed by MEASUREs: the prefix is replayed, MEASUREs step as before;
    seeded outcomes equal the uncached QVM's.  SWAP relabels inside the prefix are remembered."""
    n = 12
    spec = orc.random_circuit_spec(n, 5, 8)
    prog = spec_program(spec)
    prog.inst(("SWAP", 0, 7)).inst(("SWAP", 3, 11)).inst(H(7))
    for q in (7, 0, 5):
        prog.measure(q, [q])
    prog.inst(X(2)).measure(2, [2])
    cached, plain = QVMConnection(), QVMConnection()
    plain.plan_cache = False
    for rep in range(3):
        np.random.seed(100 + rep)
        a = cached.wavefunction(prog, [7, 0, 5, 2])
        np.random.seed(100 + rep)
        b = plain.wavefunction(prog, [7, 0, 5, 2])
        assert a[1] == b[1]
        assert np.max(np.abs(np.asarray(a[0].amplitudes) - np.asarray(b[0].amplitudes))) < 1e-13
    np.random.seed(3)
    r1 = cached.run(prog, [7, 0, 5, 2], trials=20)
    np.random.seed(3)
    r2 = plain.run(prog, [7, 0, 5, 2], trials=20)
    assert r1 == r2


def test_compiled_density_program_matches_uncached():
    n = 6
    spec = [("H", [], [q]) for q in range(n)] + orc.random_circuit_spec(n, 3, 4)
    prog = spec_program(spec)
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    cached = QVMConnection(type_trans="density", noise_model=nm)
    plain = QVMConnection(type_trans="density", noise_model=nm)
    plain.plan_cache = False
    want = np.asarray(plain.density(prog).todense())
    for _ in range(3):
        got = np.asarray(cached.density(prog).todense())
        assert np.max(np.abs(got - want)) < 1e-14
    # another noise model on the same Program object is another plan
    nm2 = NoiseModel(T1=20e-6, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    a = np.asarray(QVMConnection(type_trans="density", noise_model=nm2).density(prog).todense())
    off = QVMConnection(type_trans="density", noise_model=nm2)
    off.plan_cache = False
    assert np.max(np.abs(a - np.asarray(off.density(prog).todense()))) < 1e-14
    assert np.max(np.abs(a - want)) > 1e-6


# ---- expectation values -----------------------------------------------------------------------------------------
def _golden_terms(case):
    from pyquil.paulis import PauliSum, PauliTerm
    terms = []
    for t in case["terms"]:
        term = PauliTerm("I", 0, complex(t["re"], t["im"]))
        for q, p in t["ops"].items():
            term = term * PauliTerm(p, int(q))
        terms.append(term)
    return PauliSum(terms)


def test_golden_expectation_values(qvm, qvm_unitary):
    """expectation() of the wavefunction, unitary and density QVMs against values assembled from the reference's
    own wavefunction / density and tensor_up (the reference's expectation() is NotImplementedError)."""
    meta, arrays = load_golden("expectation")
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    dq = QVMConnection(type_trans="density", noise_model=nm)
    by_program = {}
    for case in meta:
        by_program.setdefault(json.dumps(case["spec"]), []).append(case)
    for spec_json, cases in by_program.items():
        prog = spec_program([(a, b, c) for a, b, c in json.loads(spec_json)])
        sums = [_golden_terms(c) for c in cases]
        got_w = qvm.expectation(prog, sums)
        got_u = qvm_unitary.expectation(prog, sums)
        got_d = dq.expectation(prog, sums)
        for case, w, u, d in zip(cases, got_w, got_u, got_d):
            want_psi, want_rho = arrays[case["key"]]
            assert abs(w - want_psi) < TOL and abs(u - want_psi) < TOL
            assert abs(d - want_rho) < TOL


def test_expectation_of_operator_programs(qvm):
    """Operator PROGRAMS (the pyquil 1.x call shape): Pauli-gate programs take the reduction path, anything else is
    applied to a copy of the state; the prepared state stays untouched between operators; default = identity."""
    from pyquil.paulis import PauliTerm
    n = 14
    spec = orc.random_circuit_spec(n, 5, 12)
    prog = spec_program(spec)
    psi, _ = orc.fast_run_wavefunction(oracle_ops(prog), n)
    ops = [Program().inst(X(0)).inst(Z(13)), Program().inst(H(3)).inst(CNOT(3, 9)), Program().inst(RZ(0.3, 5)),
           Program().inst(X(2)).inst(X(2)).inst(Y(2)), Program(), PauliTerm("Y", 13) * PauliTerm("X", 12, 2.0)]
    got = qvm.expectation(prog, ops)
    want = []
    for op in ops[:5]:
        phi, _ = orc.fast_run_wavefunction(oracle_ops(op), n, psi0=psi) if len(op) else (psi, None)
        want.append(np.vdot(psi, phi))
    want.append(orc.fast_expectation_state(psi, [(2.0, {13: "Y", 12: "X"})], n))
    for g, w in zip(got, want):
        assert abs(g - w) < TOL
    assert qvm.expectation(prog) == [1.0] or abs(qvm.expectation(prog)[0] - 1.0) < TOL
    with pytest.raises(IndexError):
        qvm.expectation(prog, [PauliTerm("X", n)])
    # SWAP relabels in the preparation are honoured
    sprog = spec_program(spec)
    sprog.inst(("SWAP", 0, 9))
    spsi, _ = orc.fast_run_wavefunction(oracle_ops(sprog), n)
    assert abs(qvm.expectation(sprog, [PauliTerm("Z", 9) * PauliTerm("X", 0)])[0]
               - orc.fast_expectation_state(spsi, [(1.0, {9: "Z", 0: "X"})], n)) < TOL


# ---- batched trials under classical control ----------------------------------------------------------------------
def test_batched_run_with_jumps_matches_the_per_trial_distribution(qvm):
    """run(trials) on programs with if_then / while_do: trials are batched and regrouped at every divergent jump;
    deterministic consequences are exact, distributions match the per-trial loop (chi-squared)."""
    from pyquil.gates import MEASURE
    # if_then: measure a coin, flip another qubit only when it came up 1 -> the two bits are always equal
    branch = Program().inst(H(0)).measure(0, [0])
    branch.if_then(0, Program().inst(X(1)), Program().inst(I(1)))
    branch.measure(1, [1])
    got = np.asarray(qvm.run(branch, [0, 1], trials=4000))
    assert np.array_equal(got[:, 0], got[:, 1]) and abs(got[:, 0].mean() - 0.5) < 0.04
    # while_do: keep tossing until tails; bit 2 counts parity of the tosses -> geometric law
    loop = Program().inst(X(0)).measure(0, [0])
    loop.while_do(0, Program().inst(H(0)).measure(0, [0]).inst(X(1)))
    loop.measure(1, [1])
    trials = 6000
    batched = np.asarray(qvm.run(loop, [0, 1], trials=trials))
    assert not batched[:, 0].any()                          # the loop only ends with bit 0 cleared
    # parity of the number of iterations: P(odd) = sum_{k odd} 2^-k = 2/3
    assert abs(batched[:, 1].mean() - 2.0 / 3.0) < 0.03
    plain = QVMConnection()
    plain.sampling_fast_path = False
    ref = np.asarray(plain.run(loop, [0, 1], trials=300))
    assert not ref[:, 0].any() and abs(ref[:, 1].mean() - 2.0 / 3.0) < 0.12
    # nested control flow with a rotation: outcome probabilities against the closed form
    theta = 1.1
    prog = Program().inst(RX(theta, 0)).measure(0, [0])
    prog.if_then(0, Program().inst(RX(theta, 1)).measure(1, [1]).if_then(1, Program().inst(X(2))), Program().inst(H(2)))
    prog.measure(2, [2])
    res = np.asarray(qvm.run(prog, [0, 1, 2], trials=8000))
    p1 = np.sin(theta / 2) ** 2
    assert abs(res[:, 0].mean() - p1) < 0.03
    assert abs((res[:, 0] & res[:, 1]).mean() - p1 * p1) < 0.03
    want2 = (1 - p1) * 0.5 + p1 * p1
    assert abs(res[:, 2].mean() - want2) < 0.03
    # the last trial's state is left in self.wf like the reference's loop
    assert abs(np.vdot(qvm.wf, qvm.wf).real - 1.0) < 1e-12


def test_stabilizer_qvm_against_the_wavefunction_qvm(qvm):
    """Reference tests/test_stabilizer_qvm.py:366-400 (test_random_stabilizer_circuit): a random H / S / CNOT
    circuit on the tableau simulator and on the (GPU) wavefunction QVM give the same state up to a global phase."""
    from pyquil.gates import S
    rs = np.random.RandomState(42)
    n = 5
    prog = Program()
    for _ in range(300):
        g = rs.randint(3)
        if g == 0:
            prog.inst(H(int(rs.randint(n))))
        elif g == 1:
            prog.inst(S(int(rs.randint(n))))
        else:
            a = int(rs.randint(n))
            prog.inst(CNOT(a, int((a + 1 + rs.randint(n - 1)) % n)))
    stab = QVMConnection(type_trans="stabilizer")
    a = np.asarray(stab.wavefunction(prog).amplitudes)
    b = np.asarray(qvm.wavefunction(prog)[0].amplitudes)
    assert abs(abs(np.vdot(a, b)) - 1.0) < 1e-12
    rho = np.asarray(stab.density(prog))
    assert np.max(np.abs(rho - np.outer(b, b.conj()))) < 1e-12
