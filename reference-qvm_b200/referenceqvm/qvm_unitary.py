"""Unitary QVM: the program's matrix, built by the same gate kernels.

Surface follows reference ``referenceqvm/qvm_unitary.py`` (transition :58-72, unitary :74-97,
expectation :99-112).  ``umat`` is held in HBM as a 2n-qubit vector (row-major, index
(row << n) | col); ``U . umat`` acts on the row index, i.e. the gate lands on vector-qubits a+n.
Starting from the identity, its 2^n columns are 2^n independent states evolved in one sweep.
"""
import os

import numpy as np

from pyquil.quil import Program

from referenceqvm import _engine
from referenceqvm.qam import QAM
from referenceqvm.unitary_generator import resolve_gate, tensor_gates  # noqa: F401


class QVM_Unitary(QAM):
    """Only programs of pure Gates / DefGates are accepted (no MEASURE, no control flow)."""

    def __init__(self, num_qubits=None, program=None, program_counter=None, classical_memory=None,
                 gate_set=None, defgate_set=None, unitary=None):
        super(QVM_Unitary, self).__init__(num_qubits=num_qubits, program=program,
                                          program_counter=program_counter,
                                          classical_memory=classical_memory,
                                          gate_set=gate_set, defgate_set=defgate_set)
        self.all_inst = False
        self._engine = None
        self.device = int(os.environ.get("QVM_B200_DEVICE", "0"))
        if unitary is not None:
            self.umat = unitary

    @property
    def umat(self):
        if self._engine is None:
            return None
        dim = 1 << (self._engine.n // 2)
        return self._engine.amplitudes().reshape(dim, dim)

    @umat.setter
    def umat(self, value):
        if value is None:
            if self._engine is not None:
                self._engine.close()
            self._engine = None
            return
        dense = np.ascontiguousarray(value, dtype=np.complex128)
        n = dense.shape[0].bit_length() - 1
        if self._engine is None or self._engine.n != 2 * n:
            if self._engine is not None:
                self._engine.close()
            self._engine = _engine.Engine(2 * n, device=self.device)
        self._engine.set_amplitudes(dense.reshape(-1))

    def transition(self, instruction):
        if instruction.name in self.gate_set or instruction.name in self.defgate_set:
            matrix, qubits = resolve_gate(self.gate_set, self.defgate_set, instruction)
            _engine.validate_qubits(qubits, self.num_qubits)
            self._engine.queue_matrix(matrix, [q + self.num_qubits for q in qubits])
            self.program_counter += 1
        else:
            raise TypeError("Gate {} is not in the gate set".format(instruction.name))

    def unitary(self, pyquil_program):
        """The unitary of a protoQuil program as a dense complex ndarray."""
        self.load_program(pyquil_program)
        self.umat = np.eye(2 ** self.num_qubits)
        self.kernel()
        return self.umat

    def expectation(self, pyquil_program, operator_programs=None):
        """[<0|U^dagger O U|0> for O in operator_programs], U the unitary of ``pyquil_program``.

        The reference declares this method and raises NotImplementedError (qvm_unitary.py:99-112).  The
        prepared state is the first column of U, which the device already holds: it is copied into an n-qubit
        state and the operators are evaluated there (referenceqvm/_expectation.py)."""
        from referenceqvm import _expectation
        if operator_programs is None:
            operator_programs = [Program()]
        self.load_program(pyquil_program)
        self.umat = np.eye(2 ** self.num_qubits)
        self.kernel()
        n = self.num_qubits
        dim = 1 << n
        column = self._engine.strided(0, dim, dim)              # U[:, 0]: vector index (row << n) | 0
        psi = _engine.Engine(n, device=self.device)
        try:
            psi.set_amplitudes(column)

            def general(program):
                scratch = psi.clone()
                try:
                    defs = dict((dg.name, dg.matrix) for dg in program.defined_gates)
                    for instruction in program:
                        matrix, qubits = resolve_gate(self.gate_set, defs, instruction)
                        _engine.validate_qubits(qubits, n)
                        scratch.queue_matrix(matrix, qubits)
                    return psi.inner_product(scratch)
                finally:
                    scratch.close()

            return [_expectation.evaluate(op, n, psi.pauli_expectation, general) for op in operator_programs]
        finally:
            psi.close()
