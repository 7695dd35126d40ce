"""The stabilizer QVM (host-side tableau simulator, no GPU involved) against fixtures produced by the reference's
own ``QVM_Stabilizer`` (tests/golden/make_golden.py::golden_stabilizer): final tableau bit for bit, projected
wavefunction and density matrix, seeded run() outcomes and RNG stream position."""
import numpy as np
import pytest

from pyquil.gates import CNOT, CZ, H, S, X, Y, Z
from pyquil.quil import Program

from _programs import load_golden, program_from_json
from referenceqvm.api import QVMConnection
from referenceqvm.qvm_stabilizer import (QVM_Stabilizer, binary_stabilizer_to_pauli_stabilizer,
                                         pauli_stabilizer_to_binary_stabilizer, project_stabilized_state,
                                         symplectic_inner_product)


def test_golden_stabilizer_programs():
    meta, arrays = load_golden("stabilizer")
    qvm = QVMConnection(type_trans="stabilizer")
    assert isinstance(qvm, QVM_Stabilizer)
    for case in meta:
        name, n = case["name"], case["n"]
        prog = program_from_json(case["program"])
        wf = qvm.wavefunction(prog)
        assert np.array_equal(qvm.tableau, arrays[name + ":tableau"]), name
        ref = arrays[name + ":wf"]
        mine = np.asarray(wf.amplitudes)
        if abs(np.vdot(ref, ref).real - 1.0) < 1e-9:
            assert np.max(np.abs(mine - ref)) < 1e-12, name
        else:
            # |+>^n is orthogonal to this state: the reference's projection divides by zero and returns an
            # all-zero / NaN vector; here another start state is projected.  Check the defining property instead:
            # every stabilizer generator fixes the state.
            from referenceqvm.qvm_stabilizer import _apply_pauli_term
            assert abs(np.vdot(mine, mine).real - 1.0) < 1e-12, name
            for g in binary_stabilizer_to_pauli_stabilizer(qvm.stabilizer_tableau()):
                assert np.max(np.abs(_apply_pauli_term(mine, g, n) - mine)) < 1e-12, name
        assert np.max(np.abs(np.asarray(qvm.density(prog)) - arrays[name + ":rho"])) < 1e-12, name
        mprog = Program().inst(prog)
        for q in range(n):
            mprog.measure(q, [q])
        np.random.seed(case["seed"])
        runs = np.array(qvm.run(mprog, list(range(n)), trials=25))
        if case["runs"]:
            assert np.array_equal(runs, arrays[name + ":runs"]), name
            assert np.random.random() == arrays[name + ":rng_after"][0], name
        else:
            # (the reference raises on this program: a destabilizer product with phase +-i, see make_golden.py);
            # here the run completes; its outcomes must be consistent with the state's distribution
            probs = np.abs(mine) ** 2
            for bits in runs:
                assert probs[sum(int(b) << q for q, b in enumerate(bits))] > 1e-12, name


def test_pauli_gates_and_cz_beyond_the_reference():
    """X, Y, Z and CZ are in the reference's stabilizer gate set but its transition refuses them; here they run.
    Checked through Clifford identities: X = H Z H, Z = S S, Y ~ S X S^dagger (same stabilizer state), CZ = H CNOT H."""
    qvm = QVMConnection(type_trans="stabilizer")
    prep = [H(0), CNOT(0, 1), S(1), H(2), CNOT(2, 0)]
    pairs = [([X(1)], [H(1), S(1), S(1), H(1)]), ([Z(2)], [S(2), S(2)]),
             ([Y(0)], [S(0), S(0), S(0), H(0), S(0), S(0), H(0), S(0)]),       # S^3 X S = Y up to a phase
             ([CZ(1, 2)], [H(2), CNOT(1, 2), H(2)])]
    for lhs, rhs in pairs:
        a = np.asarray(qvm.density(Program().inst(prep + lhs)))
        b = np.asarray(qvm.density(Program().inst(prep + rhs)))
        assert np.max(np.abs(a - b)) < 1e-12
    with pytest.raises(TypeError):
        qvm.run(Program().inst(H(0)).inst(("T", 0)))


def test_stabilizer_utilities_roundtrip():
    qvm = QVMConnection(type_trans="stabilizer")
    qvm.wavefunction(Program().inst(H(0), CNOT(0, 1), S(0), H(2), CNOT(2, 1)))
    stab = qvm.stabilizer_tableau()
    terms = binary_stabilizer_to_pauli_stabilizer(stab)
    assert np.array_equal(pauli_stabilizer_to_binary_stabilizer(terms), stab)
    n = qvm.num_qubits
    for i in range(n):                     # stabilizers commute with each other; destabilizer i anticommutes with stabilizer i only
        for j in range(n):
            flipped = np.concatenate((stab[j, n:2 * n], stab[j, :n]))          # symplectic form: x.z' + z.x'
            assert symplectic_inner_product(stab[i, :2 * n], flipped) == 0
            assert symplectic_inner_product(qvm.destabilizer_tableau()[i, :2 * n], flipped) == (1 if i == j else 0)
    psi = project_stabilized_state(terms, n).flatten()
    assert abs(np.vdot(psi, psi) - 1.0) < 1e-12
    # a state orthogonal to |+>^n (qubit in |->): the reference divides by zero, here another start state is projected
    minus = qvm.wavefunction(Program().inst(H(0), S(0), S(0)))
    assert np.allclose(np.abs(minus.amplitudes), [2 ** -0.5, 2 ** -0.5])
