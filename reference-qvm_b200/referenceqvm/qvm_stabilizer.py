"""Stabilizer QVM: Clifford circuits on a tableau (Aaronson & Gottesman, Phys. Rev. A 70, 052328).

Host-side integer work, polynomial in the number of qubits: there is no state vector and nothing for the GPU to
do, so this module stays numpy -- it exists so that ``QVMConnection(type_trans='stabilizer')`` is as usable here
as in the reference (``referenceqvm/qvm_stabilizer.py``).  Same surface: ``run`` (:203-232), ``density``
(:234-251), ``wavefunction`` (:253-269), ``tableau`` / ``stabilizer_tableau`` / ``destabilizer_tableau``
(:133-139, :271-286), ``load_program`` (:70-131: gates of the stabilizer gate set and MEASURE only).

The tableau is the reference's: a (2n) x (2n + 1) integer array, rows 0..n-1 destabilizers, rows n..2n-1
stabilizers, columns x_0..x_(n-1) | z_0..z_(n-1) | sign.  Gates update whole columns at once (the reference
loops over rows, :377-425); MEASURE follows :427-479 step for step, including its single
``np.random.random() > 0.5`` draw per random outcome, so seeded programs give the reference's bits.
X, Y, Z and CZ are in the reference's stabilizer gate set (gates.py:229-238) but its ``_transition`` refuses
them ("the impossible has happened", :178-180); here they are applied (sign flips / H.CNOT.H).
"""
from functools import reduce

import numpy as np

from pyquil.paulis import PauliSum, PauliTerm, sI, sX, sY, sZ
from pyquil.quil import Program, get_classical_addresses_from_program
from pyquil.quilbase import Gate, Jump, JumpTarget, Measurement
from pyquil.wavefunction import Wavefunction

from referenceqvm.gates import stabilizer_gate_matrix
from referenceqvm.qam import CBIT_FLOOR, QAM
from referenceqvm.unitary_generator import tensor_up, value_get


def binary_stabilizer_to_pauli_stabilizer(stabilizer_tableau):
    """Rows of a tableau as signed PauliTerms (reference stabilizer_utils.py:176-207)."""
    tableau = np.asarray(stabilizer_tableau)
    n = (tableau.shape[1] - 1) // 2
    letters = {(1, 0): sX, (0, 1): sZ, (1, 1): sY}
    terms = []
    for row in tableau:
        factors = [letters[(int(row[q]), int(row[q + n]))](q) for q in range(n) if row[q] or row[q + n]]
        term = reduce(lambda a, b: a * b, factors) if factors else sI(0)
        terms.append(term * ((-1) ** int(row[-1])))
    return terms


def pauli_stabilizer_to_binary_stabilizer(stabilizer_list):
    """The inverse (reference stabilizer_utils.py:145-174): PauliTerms with coefficient +-1 -> tableau rows."""
    if not all(isinstance(t, PauliTerm) for t in stabilizer_list):
        raise TypeError("At least one element of stabilizer_list is not a PauliTerm")
    n = max(max(t.get_qubits()) for t in stabilizer_list) + 1
    out = np.zeros((len(stabilizer_list), 2 * n + 1), dtype=int)
    for r, term in enumerate(stabilizer_list):
        for q, letter in term:
            out[r, q] = letter in ("X", "Y")
            out[r, q + n] = letter in ("Z", "Y")
        if not (np.isclose(term.coefficient, 1) or np.isclose(term.coefficient, -1)):
            raise ValueError("stabilizers must have a +/- coefficient")
        out[r, -1] = 0 if term.coefficient.real > 0 else 1
    return out


def symplectic_inner_product(vector1, vector2):
    """0 when the two (sign-less) binary Pauli vectors commute, 1 when they anticommute -- computed as the
    reference does (stabilizer_utils.py:210-226: XOR over the element-wise product)."""
    vector1, vector2 = np.asarray(vector1), np.asarray(vector2)
    if vector1.shape != vector2.shape:
        raise ValueError("vectors must be the same size.")
    return int(np.bitwise_xor.reduce(np.multiply(vector1, vector2)))


def _apply_pauli_term(state, term, n):
    """term |state> on a dense little-endian vector (qubit q = index bit q)."""
    idx = np.arange(state.size)
    out_idx = idx.copy()
    phase = np.full(state.size, complex(term.coefficient))
    for q, letter in term:
        bit = (idx >> q) & 1
        if letter == "X":
            out_idx ^= 1 << q
        elif letter == "Y":
            out_idx ^= 1 << q
            phase = phase * np.where(bit == 0, 1j, -1j)       # Y|0> = i|1>, Y|1> = -i|0>
        elif letter == "Z":
            phase = phase * np.where(bit == 0, 1.0, -1.0)
    out = np.zeros_like(state)
    out[out_idx] = phase * state
    return out


def project_stabilized_state(stabilizer_list, num_qubits=None, classical_state=None):
    """|psi> = normalised prod_i (1 + G_i)/2 applied to |+>^n (or to a classical bit list, qubit 0 first):
    the state the stabilizers fix, with the reference's phase convention (stabilizer_utils.py:100-142).
    Returned as a dense column (2^n x 1) like the reference's ``.toarray()``."""
    n = len(stabilizer_list) if num_qubits is None else num_qubits
    if classical_state is None:
        state = np.full(2 ** n, 1.0 / np.sqrt(2 ** n), dtype=complex)
    else:
        if not isinstance(classical_state, list):
            raise TypeError("I only accept lists as the classical state")
        if len(classical_state) != n:
            raise TypeError("Classical state does not match the number of qubits")
        state = np.zeros(2 ** n, dtype=complex)
        state[sum(int(b) << q for q, b in enumerate(classical_state))] = 1.0
    start = state
    for attempt in range(2 ** n + 1):
        state = start
        for generator in stabilizer_list:
            state = (state + _apply_pauli_term(state, generator, n)) / 2
        norm = float(np.vdot(state, state).real)
        if norm > 1e-12 or classical_state is not None or attempt == 2 ** n:
            break
        # |+>^n is orthogonal to the stabilized state (e.g. a qubit in |->): the reference divides by zero here;
        # project computational basis states instead until one overlaps
        start = np.zeros(2 ** n, dtype=complex)
        start[attempt] = 1.0
    return (state / np.sqrt(norm)).reshape(-1, 1)


class QVM_Stabilizer(QAM):
    """Supports run(), density() and wavefunction() for programs of I, X, Y, Z, H, S, CNOT, CZ and MEASURE."""

    def __init__(self, num_qubits=None, program=None, program_counter=None, classical_memory=None,
                 gate_set=stabilizer_gate_matrix, defgate_set=None):
        super(QVM_Stabilizer, self).__init__(num_qubits=num_qubits, program=program, program_counter=program_counter,
                                             classical_memory=classical_memory, gate_set=gate_set,
                                             defgate_set=defgate_set)
        self.all_inst = False
        self.tableau = None if num_qubits is None else self._n_qubit_tableau(num_qubits)

    # -- loading (reference :70-131: gates of the set and MEASURE pass, anything else is refused) ---------------
    def load_program(self, pyquil_program):
        if not isinstance(pyquil_program, Program):
            raise TypeError("I can only generate from pyQuil programs")
        for instruction in pyquil_program:
            ok = (isinstance(instruction, Gate) and instruction.name in self.gate_set) or \
                isinstance(instruction, Measurement)
            if not ok and self.all_inst is False:
                raise TypeError("Some gates used are not allowed in this QAM")
        self.program = pyquil_program
        self.program_counter = 0
        q_max, c_max = self.identify_bits()
        self.num_qubits = q_max
        self.classical_memory = np.zeros(max(c_max, CBIT_FLOOR)).astype(bool)

    def _n_qubit_tableau(self, num_qubits):
        """The tableau of |0...0>: destabilizers X_q, stabilizers Z_q (reference :133-139)."""
        tableau = np.zeros((2 * num_qubits, 2 * num_qubits + 1), dtype=int)
        tableau[np.arange(2 * num_qubits), np.arange(2 * num_qubits)] = 1
        return tableau

    def stabilizer_tableau(self):
        return self.tableau[self.num_qubits:, :]

    def destabilizer_tableau(self):
        return self.tableau[:self.num_qubits, :]

    # -- gates: whole columns at once -------------------------------------------------------------------------------
    def _x(self, q):
        return self.tableau[:, q]

    def _z(self, q):
        return self.tableau[:, q + self.num_qubits]

    def _apply_hadamard(self, instruction):
        q = value_get(instruction.qubits[0])
        t, n = self.tableau, self.num_qubits
        t[:, -1] ^= t[:, q] & t[:, q + n]
        t[:, [q, q + n]] = t[:, [q + n, q]]

    def _apply_phase(self, instruction):
        q = value_get(instruction.qubits[0])
        t, n = self.tableau, self.num_qubits
        t[:, -1] ^= t[:, q] & t[:, q + n]
        t[:, q + n] ^= t[:, q]

    def _apply_cnot(self, instruction):
        a, b = [value_get(x) for x in instruction.qubits]
        t, n = self.tableau, self.num_qubits
        t[:, -1] ^= t[:, a] & t[:, b + n] & (t[:, b] ^ t[:, a + n] ^ 1)
        t[:, b] ^= t[:, a]
        t[:, a + n] ^= t[:, b + n]

    def _apply_pauli(self, name, q):
        """X, Y, Z conjugate a Pauli row to +-itself: only signs move."""
        t, n = self.tableau, self.num_qubits
        if name == "X":
            t[:, -1] ^= t[:, q + n]
        elif name == "Z":
            t[:, -1] ^= t[:, q]
        else:
            t[:, -1] ^= t[:, q] ^ t[:, q + n]

    # -- rowsum (reference :288-375) ------------------------------------------------------------------------------------
    @staticmethod
    def _g(x1, z1, x2, z2):
        """Power of i picked up when the Pauli (x1 z1) multiplies (x2 z2), element-wise."""
        return np.where((x1 == 0) & (z1 == 0), 0,
                        np.where((x1 == 1) & (z1 == 1), z2 - x2,
                                 np.where((x1 == 1) & (z1 == 0), z2 * (2 * x2 - 1), x2 * (1 - 2 * z2))))

    def _rowsum(self, h, i):
        """row h <- row i * row h (left multiplication), sign from the accumulated powers of i."""
        t, n = self.tableau, self.num_qubits
        acc = int(np.sum(self._g(t[i, :n], t[i, n:2 * n], t[h, :n], t[h, n:2 * n]))) + 2 * int(t[h, -1]) + 2 * int(t[i, -1])
        acc %= 4
        if acc not in (0, 2) and h >= n:
            raise ValueError("An impossible value for the phase_accumulator has occurred")
        # (destabilizer rows, h < n: their products may carry +-i; Aaronson-Gottesman never use those signs.  The
        # reference raises there too, which makes it fail on some valid Clifford programs with MEASURE.)
        t[h, -1] = (acc // 2) % 2
        t[h, :2 * n] ^= t[i, :2 * n]

    def _apply_measurement(self, instruction):
        """MEASURE (reference :427-479): random when some stabilizer anticommutes with Z_q, else determined."""
        q, c = value_get(instruction.qubit), value_get(instruction.classical_reg)
        n = self.num_qubits
        hits = np.nonzero(self.tableau[n:, q] == 1)[0]
        if hits.size:
            p = int(hits[0]) + n
            for row in range(2 * n):
                if row != p and self.tableau[row, q] == 1:
                    self._rowsum(row, p)
            self.tableau[p - n, :] = self.tableau[p, :]
            self.tableau[p, :] = 0
            self.tableau[p, q + n] = 1
            self.tableau[p, -1] = 1 if np.random.random() > 0.5 else 0
            self.classical_memory[c] = self.tableau[p, -1]
        else:
            self.tableau = np.vstack((self.tableau, np.zeros((1, 2 * n + 1), dtype=int)))
            for row in range(n):
                if self.tableau[row, q] == 1:
                    self._rowsum(2 * n, row + n)
            self.classical_memory[c] = self.tableau[-1, -1]
            self.tableau = self.tableau[:2 * n, :]

    # -- state machine --------------------------------------------------------------------------------------------------
    def transition(self, instruction):
        self.pre()
        self._transition(instruction)
        self.post()
        return self.program_counter == len(self.program)

    def pre(self):
        pass

    def post(self):
        pass

    def _transition(self, instruction):
        if isinstance(instruction, Measurement):
            self._apply_measurement(instruction)
            self.program_counter += 1
        elif isinstance(instruction, Gate):
            name = instruction.name
            if name == "H":
                self._apply_hadamard(instruction)
            elif name == "S":
                self._apply_phase(instruction)
            elif name == "CNOT":
                self._apply_cnot(instruction)
            elif name in ("X", "Y", "Z"):
                self._apply_pauli(name, value_get(instruction.qubits[0]))
            elif name == "CZ":
                a, b = instruction.qubits
                target = Gate("H", [], [b])
                self._apply_hadamard(target)
                self._apply_cnot(Gate("CNOT", [], [a, b]))
                self._apply_hadamard(target)
            elif name != "I":
                raise TypeError("We checked for correct gate types previously so the impossible has happened!")
            self.program_counter += 1
        elif isinstance(instruction, Jump):
            self.program_counter = self.find_label(instruction.target)
        elif isinstance(instruction, JumpTarget):
            self.program_counter += 1

    # -- public API -----------------------------------------------------------------------------------------------------
    def _simulate(self, pyquil_program):
        self.load_program(pyquil_program)
        self.tableau = self._n_qubit_tableau(self.num_qubits)
        self.kernel()

    def run(self, pyquil_program, classical_addresses=None, trials=1):
        """Run ``trials`` times; the classical bits at ``classical_addresses`` (default: every address the
        program touches) after each run."""
        self.load_program(pyquil_program)
        if classical_addresses is None:
            classical_addresses = get_classical_addresses_from_program(pyquil_program)
        results = []
        for _ in range(trials):
            self.tableau = self._n_qubit_tableau(self.num_qubits)
            self.kernel()
            results.append([int(b) for b in self.classical_memory[classical_addresses]])
            self.memory_reset()
            self.program_counter = 0
        return results

    def density(self, pyquil_program):
        """rho = prod_i (1 + g_i)/2 over the stabilizer generators, as a dense matrix (reference :234-251)."""
        self._simulate(pyquil_program)
        stabilizers = binary_stabilizer_to_pauli_stabilizer(self.stabilizer_tableau())
        projector = reduce(lambda a, b: a * b, [0.5 * (sI(0) + g) for g in stabilizers])
        if isinstance(projector, PauliTerm):
            projector = PauliSum([projector])
        return tensor_up(projector, self.num_qubits)

    def wavefunction(self, program):
        """The stabilized state as a pyQuil Wavefunction (reference :253-269; phase convention of
        ``project_stabilized_state``)."""
        self._simulate(program)
        stabilizers = binary_stabilizer_to_pauli_stabilizer(self.stabilizer_tableau())
        return Wavefunction(project_stabilized_state(stabilizers, self.num_qubits).flatten())
