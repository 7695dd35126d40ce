"""Host-side state machine pieces that need no device (reference tests/test_qvm_funcs.py,
parts of tests/test_controlflow.py and the API factory)."""
import numpy as np
import pytest

from pyquil.gates import CNOT, H, MEASURE, TRUE, FALSE, NOT, AND, OR, MOVE, EXCHANGE, X, HALT
from pyquil.quil import Program
from pyquil.quilbase import Jump, JumpTarget, JumpWhen, Label

from referenceqvm.api import QVMConnection
from referenceqvm.qvm_density import QVM_Density, NoiseModel, INFINITY, kraus_superoperator
from referenceqvm.qvm_unitary import QVM_Unitary
from referenceqvm.qvm_wavefunction import QVM_Wavefunction
from referenceqvm import gates as G
from referenceqvm import unitary_generator as ug


def test_factory_types_and_errors():
    assert isinstance(QVMConnection(), QVM_Wavefunction)
    assert isinstance(QVMConnection(type_trans="wavefunction"), QVM_Wavefunction)
    assert isinstance(QVMConnection(type_trans="unitary"), QVM_Unitary)
    nm = NoiseModel(T1=INFINITY)
    d = QVMConnection(type_trans="density", noise_model=nm)
    assert isinstance(d, QVM_Density) and d.noise_model is nm
    with pytest.raises(TypeError):
        QVMConnection(type_trans="tensor-network")
    from referenceqvm.qvm_stabilizer import QVM_Stabilizer
    assert isinstance(QVMConnection(type_trans="stabilizer"), QVM_Stabilizer)


def test_identify_bits_like_reference():
    qvm = QVMConnection()
    p = Program()
    for i in range(5):
        p.inst(X(i))
    for i in range(670):
        p.inst(TRUE(i))
    qvm.load_program(p)
    assert qvm.num_qubits == 5 and len(qvm.classical_memory) == 670
    p = Program()
    for i in range(670):
        p.inst(TRUE(i))
    qvm.load_program(p)
    assert qvm.num_qubits == 0 and len(qvm.classical_memory) == 670
    qvm.load_program(Program(X(0)))
    assert qvm.num_qubits == 1 and len(qvm.classical_memory) == 512
    assert qvm.classical_memory.dtype == bool and qvm.program_counter == 0


def test_load_program_errors():
    qvm = QVMConnection()
    with pytest.raises(TypeError):
        qvm.load_program([H(0)])
    with pytest.raises(RuntimeError):
        qvm.load_program(Program(H(51)))
    qvm.load_program(Program(H(50)))
    assert qvm.num_qubits == 51
    qu = QVMConnection(type_trans="unitary")
    with pytest.raises(TypeError):
        qu.load_program(Program(H(0)).measure(0, [0]))
    prog = Program()
    prog.defgate("hello", np.array([[0, 1], [1, 0]]))
    prog.inst(("hello2", 0))
    with pytest.raises(TypeError):
        qu.load_program(prog)


def test_classical_and_control_flow_transitions_without_device():
    qvm = QVMConnection()
    lbl = Label("L")
    prog = Program(TRUE(0), FALSE(1), NOT(1), AND(0, 2), OR(0, 2), MOVE(2, 3), EXCHANGE(3, 4), FALSE(0))
    prog.inst(JumpWhen(lbl, 0), TRUE(7), JumpTarget(lbl), TRUE(8), HALT, TRUE(9))
    qvm.load_program(prog)
    while qvm.program_counter < len(qvm.program):
        assert qvm._classical_transition(qvm.current_instruction())
    mem = qvm.classical_memory
    assert list(mem[:5].astype(int)) == [0, 1, 1, 0, 1]
    assert mem[7] and mem[8] and not mem[9]
    with pytest.raises(RuntimeError):
        qvm.find_label(Label("missing"))
    qvm.load_program(Program(Jump(Label("nowhere"))))
    with pytest.raises(RuntimeError):
        qvm._classical_transition(qvm.current_instruction())


def test_operator_descriptor_validation():
    # reference tests/test_unitary_generator.py:34-41
    for i, n in [(2, 3), (3, 3), (-1, 3), (3, 4)]:
        with pytest.raises(ValueError):
            ug.lifted_gate(i, G.SWAP, n)
    op = ug.lifted_gate(8, G.SWAP, 10)
    assert op.qubits == (9, 8) and op.shape == (1024, 1024)
    assert ug.apply_gate(G.CNOT, [1, 0], 2).qubits == (1, 0)
    with pytest.raises(ValueError):
        ug.apply_gate(G.H, [3], 3)
    with pytest.raises(ValueError):
        ug.apply_gate(G.CNOT, [1, 1], 3)
    with pytest.raises(TypeError):
        ug.apply_gate(np.eye(3), [0], 3)
    with pytest.raises(TypeError):
        ug.apply_gate(G.CCNOT, [0, 1, 2], 2)
    with pytest.raises(ValueError):
        ug.apply_gate(G.H, [0], 0)
    from pyquil.quilbase import Gate
    with pytest.raises(ValueError):
        ug.tensor_gates(G.gate_matrix, {}, Gate("NOPE", [], [0]), 2)
    t = ug.tensor_gates(G.gate_matrix, {}, Gate("RX", [0.2], [1]), 3)
    assert np.allclose(t.matrix, G.RX(0.2)) and t.qubits == (1,)


def test_value_get():
    from pyquil.quilbase import Addr, Qubit
    assert ug.value_get(Qubit(3)) == 3 and ug.value_get(Addr(9)) == 9 and ug.value_get(Label("a")) == "a"
    assert ug.value_get(2.5) == 2.5
    with pytest.raises(TypeError):
        ug.value_get("q")


def test_kraus_superoperator_forms():
    p = 0.3
    s = kraus_superoperator(G.dephasing(p))
    assert np.allclose(s, np.diag([1, 1 - p, 1 - p, 1]))
    s = kraus_superoperator(G.relaxation(p))
    rho = np.array([[0.6, 0.2 - 0.1j], [0.2 + 0.1j, 0.4]])
    out = (s @ rho.reshape(-1)).reshape(2, 2)
    want = sum(k @ rho @ k.conj().T for k in G.relaxation(p))
    assert np.allclose(out, want)


def test_terminal_measure_analysis():
    qvm = QVMConnection()
    f = qvm._terminal_measure_block
    assert f(Program(H(0), CNOT(0, 1))) == 2
    assert f(Program(H(0)).measure(0, [0]).measure(0, [1])) == 1
    assert f(Program(H(0)).measure(0, [0]).inst(X(0))) is None
    assert f(Program(H(0), HALT).measure(0, [0])) is None
    assert f(Program(TRUE(0)).while_do(0, Program(X(0)).measure(0, 0))) is None


def test_tensor_up_small():
    from pyquil.paulis import PauliSum, PauliTerm
    ps = PauliSum([PauliTerm("X", 0, 1.0) * PauliTerm("Y", 1), PauliTerm("Y", 1, 1.0) * PauliTerm("X", 0)])
    m = ug.tensor_up(ps, 2)
    assert np.allclose(m, 2 * np.kron(G.Y, G.X))
    with pytest.raises(TypeError):
        ug.tensor_up("X0", 2)


# ---- runs of MEASUREs: the host replay of the sequential decisions (Engine.measure_many) -----------------
class _HistState(object):
    """numpy model of qvm_measure_hist (tests only): collapse, then |amp|^2 summed per (high-bin, lane)."""

    def __init__(self, psi, n):
        self.psi = np.array(psi, dtype=np.complex128)
        self.n_local, self.device, self.passes = n, 0, 0

    def measure_hist(self, high_positions, collapse=None, want_hist=True, apply=True):
        self.passes += 1
        idx = np.arange(self.psi.size, dtype=np.int64)
        psi = self.psi
        if collapse is not None:
            mask, value, scale = collapse
            psi = np.where((idx & mask) == value, self.psi * scale, 0.0)
            if apply:
                self.psi = psi
        if not want_hist:
            return None
        bins = np.zeros_like(idx)
        for i, p in enumerate(high_positions):
            assert p >= 5
            bins |= ((idx >> p) & 1) << i
        hist = np.zeros((1 << len(high_positions), 32))
        np.add.at(hist, (bins, idx & 31), np.abs(psi) ** 2)
        return hist

    def close(self):
        pass


@pytest.mark.parametrize("n,seed", [(10, 0), (12, 1), (13, 2)])
def test_measure_many_replays_the_sequential_measurements(n, seed):
    from referenceqvm import _engine
    from oracle import qvm_oracle as orc
    rs = np.random.RandomState(seed)
    psi = rs.standard_normal(1 << n) + 1j * rs.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    for order in (list(range(n)), list(rs.permutation(n)), list(rs.permutation(n)[:7]), [n - 1, 0]):
        eng = _engine.Engine(n, _state=_HistState(psi, n))
        eng.pos_of = [int(p) for p in rs.permutation(n)]          # a relabelled layout
        logical = np.arange(1 << n, dtype=np.int64)
        phys = np.zeros_like(logical)
        for q, p in enumerate(eng.pos_of):
            phys |= ((logical >> q) & 1) << p
        stored = np.empty_like(psi)
        stored[phys] = psi
        eng.state.psi = stored
        uniforms = rs.random_sample(len(order))
        got = eng.measure_many(order, uniforms)
        want = psi
        for (outcome, p0), q, r in zip(got, order, uniforms):
            o_want, p_want, want = orc.fast_measure(want, int(q), n, r)
            assert outcome == o_want and abs(p0 - p_want) < 1e-12
        assert np.max(np.abs(eng.state.psi[phys] - want)) < 1e-12
        n_high = sum(1 for q in order if eng.pos_of[q] >= 5)
        assert eng.state.passes <= (n_high + 4) // 5 + 2           # chunks of <= 5 high positions + final collapse


# ---- batched trials: host pieces that need no device ---------------------------------------------------------
def test_batched_classical_matches_the_per_trial_state_machine():
    """QVM_Wavefunction._batched_classical on a (trials x addresses) bit matrix == QAM._classical_transition
    applied to every row on its own (qam.py's semantics for TRUE/FALSE/NOT/AND/OR/MOVE/EXCHANGE)."""
    from pyquil.gates import AND, EXCHANGE, FALSE, MOVE, NOT, OR, TRUE
    from pyquil.quil import Program
    from referenceqvm.qvm_wavefunction import QVM_Wavefunction
    from referenceqvm.gates import gate_matrix
    rs = np.random.RandomState(0)
    instrs = [TRUE(0), FALSE(1), NOT(2), AND(0, 3), OR(1, 4), MOVE(2, 5), EXCHANGE(3, 6), NOT(0), AND(5, 5), OR(6, 2),
              EXCHANGE(1, 1), MOVE(7, 0)]
    mem = rs.randint(0, 2, size=(64, 8)).astype(bool)
    batched = mem.copy()
    for ins in instrs:
        QVM_Wavefunction._batched_classical(ins, batched)
    qvm = QVM_Wavefunction(gate_set=gate_matrix, defgate_set={})
    qvm.program = Program(*instrs)
    for row in range(mem.shape[0]):
        qvm.classical_memory = mem[row].copy()
        qvm.program_counter = 0
        for ins in instrs:
            assert qvm._classical_transition(ins)
        assert np.array_equal(qvm.classical_memory, batched[row]), row


def test_batchable_decision():
    from pyquil.gates import CNOT, H, MEASURE, NOT, X
    from pyquil.quil import Program
    from referenceqvm.qvm_wavefunction import QVM_Wavefunction
    from referenceqvm.gates import gate_matrix
    qvm = QVM_Wavefunction(gate_set=gate_matrix, defgate_set={})
    mid = Program(H(0), MEASURE(0, 0), CNOT(0, 1), NOT(0), MEASURE(1, 1))
    assert qvm._batchable(mid) and qvm._terminal_measure_block(mid) is None
    assert qvm._terminal_measure_block(Program(H(0), CNOT(0, 1), MEASURE(0, 0), MEASURE(1, 1))) == 2
    looped = Program(X(0), MEASURE(0, 0)).while_do(0, Program(X(0), MEASURE(0, 0))) if hasattr(Program, "while_do") else None
    if looped is not None:                           # jumps: batched too, through the regrouping flow
        assert qvm._batchable(looped) and qvm._has_control_flow(looped) and not qvm._has_control_flow(mid)
    wide = Program(H(QVM_Wavefunction.BATCH_MAX_QUBITS), MEASURE(0, 0), H(0))
    assert not qvm._batchable(wide)                  # one block per trial: segments above 2^20 amplitudes keep the loop
    assert not qvm._batchable(Program())


def test_result_rows_helper_matches_numpy(monkeypatch):
    """csrc/qvm_pyrows.c (per-trial result lists built in C) against the numpy path it replaces."""
    from referenceqvm import qvm_wavefunction as qw
    assert qw._pyrows_lib(), "referenceqvm/_pyrows.so is built by build.py"
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 2 ** 63, size=1000, dtype=np.uint64)
    idx[:4] = [0, 1, 2 ** 63, 2 ** 64 - 1]
    sel = [0, 5, 63, -1, 33, 64, 5]
    fast = qw._rows_to_lists(indices=idx, sel=sel)
    import gc
    assert type(fast[0]) is list and not gc.is_tracked(fast[0])      # invisible to the cyclic collector
    mat = rng.integers(-2 ** 62, 2 ** 62, size=(50, 12), dtype=np.int64)
    mat[:, ::2] = rng.integers(0, 2, size=(50, 6))
    fast_m = qw._rows_to_lists(mat[::2, ::-1])
    monkeypatch.setattr(qw, "_pyrows", [False])
    slow = qw._rows_to_lists(indices=idx, sel=sel)
    assert fast == slow and all(type(v) is int for v in fast[3])
    assert fast[3] == [1, 1, 1, 0, 1, 0, 1] and fast[0] == [0] * 7
    assert fast_m == mat[::2, ::-1].tolist() == qw._rows_to_lists(mat[::2, ::-1])
    assert qw._rows_to_lists(np.zeros((0, 3), dtype=np.int64)) == []
