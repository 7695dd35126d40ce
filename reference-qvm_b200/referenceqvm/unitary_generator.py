"""Operator builders of the reference, re-expressed as *descriptors* executed by CUDA kernels.

The reference materialises every gate as a 2^n x 2^n scipy-sparse matrix: identity (x) G (x)
identity (``lifted_gate``, unitary_generator.py:45-88) conjugated by a chain of adjacent-SWAP
operators (``permutation_arbitrary`` :153-256, ``two_swap_helper`` :104-150) and multiplies it into
the state (``apply_gate`` :266-323).  That construction is >= 95% of its run time.

Here ``lifted_gate`` / ``apply_gate`` / ``tensor_gates`` keep their names, argument meaning and
error behaviour but return a :class:`LiftedOperator` -- (matrix, qubits, n) -- whose ``.dot()``
runs the strided pair/quad update on the GPU (closed form).  The swap chain has
no counterpart: the kernel addresses the target bits directly.
"""
from numbers import Integral

import numpy as np

from pyquil.quilbase import Addr, Label, Qubit
from pyquil.paulis import PauliSum

from referenceqvm.gates import gate_matrix
from referenceqvm import _engine

# Mirrors the reference flag (unitary_generator.py:42): only the linear-chain mode exists there,
# and qubit connectivity is irrelevant to direct index addressing.
topological_QPU = False


def value_get(param_obj):
    """Raw number / string stored in a pyQuil atom (reference unitary_generator.py:414-428)."""
    if isinstance(param_obj, (float, int, np.integer, np.floating)):
        return param_obj
    for cls, attr in ((Qubit, "index"), (Addr, "address"), (Label, "name")):
        if isinstance(param_obj, cls):
            return getattr(param_obj, attr)
    raise TypeError("Object not recognizable")


def _gate_size(matrix):
    rows = matrix.shape[0]
    if rows == 0 or (rows & (rows - 1)) != 0:
        raise TypeError("Invalid gate size. Must be power of 2! Received {} size".format(matrix.shape))
    return rows.bit_length() - 1


class LiftedOperator(object):
    """A k-qubit matrix addressed to ``qubits`` (first = most significant) of an n-qubit register.

    Stands where the reference hands around an N x N sparse matrix.  ``dot`` applies it on the
    device; ``toarray`` materialises the N x N matrix by pushing the identity through the same
    kernels (columns of the identity are 2^n independent states)."""

    def __init__(self, matrix, qubits, num_qubits):
        self.matrix = np.asarray(matrix)
        self.qubits = tuple(int(q) for q in qubits)
        self.num_qubits = int(num_qubits)
        dim = 1 << self.num_qubits
        self.shape = (dim, dim)

    def _run(self, flat, extra_bits):
        eng = _engine.Engine(self.num_qubits + extra_bits)
        try:
            eng.set_amplitudes(flat)
            eng.queue_matrix(self.matrix, [q + extra_bits for q in self.qubits])
            return eng.amplitudes()
        finally:
            eng.close()

    def dot(self, other):
        if hasattr(other, "toarray"):
            other = other.toarray()
        arr = np.asarray(other)
        dim = self.shape[0]
        if arr.ndim == 1:
            if arr.size != dim:
                raise ValueError("dimension mismatch")
            return self._run(arr, 0)
        if arr.ndim != 2 or arr.shape[0] != dim:
            raise ValueError("dimension mismatch")
        cols = arr.shape[1]
        if cols and (cols & (cols - 1)) == 0:
            # row-major (dim, cols): the row index is the high part of a (n + log2 cols)-qubit index
            return self._run(np.ascontiguousarray(arr).reshape(-1), cols.bit_length() - 1).reshape(dim, cols)
        out = np.empty((dim, cols), dtype=np.complex128)
        for c in range(cols):
            out[:, c] = self._run(np.ascontiguousarray(arr[:, c]), 0)
        return out

    def toarray(self):
        return self.dot(np.eye(self.shape[0], dtype=np.complex128))

    def todense(self):
        return np.asmatrix(self.toarray())

    def __repr__(self):
        return "<LiftedOperator %dx%d on qubits %r of %d>" % (self.matrix.shape + (self.qubits, self.num_qubits))


def lifted_gate(i, matrix, num_qubits):
    """k-qubit ``matrix`` on the adjacent qubits (i+k-1, ..., i): I (x) matrix (x) I_{2^i}.

    Same checks as reference unitary_generator.py:73-80 (TypeError on non power-of-two size,
    ValueError when the block does not fit)."""
    matrix = np.asarray(matrix)
    k = _gate_size(matrix)
    if not (0 <= i < num_qubits + 1 - k):
        raise ValueError("Gate index out of range!")
    return LiftedOperator(matrix, range(i + k - 1, i - 1, -1), num_qubits)


def apply_gate(matrix, args, num_qubits):
    """k-qubit gate on the qubits in the order passed, GATE(args[0], ..., args[k-1]).

    Validation follows reference unitary_generator.py:284-300 and :200-203."""
    if not isinstance(num_qubits, Integral) or num_qubits < 1:
        raise ValueError("Improper number of qubits passed.")
    matrix = np.asarray(matrix)
    if len(matrix.shape) != 2 or matrix.shape[0] != matrix.shape[1]:
        raise TypeError("Gate array must be two-dimensional and square matrix.")
    k = _gate_size(matrix)
    if not (1 <= k <= num_qubits):
        raise TypeError("Invalid gate size. k-qubit gates supported, for k in [1, num_qubits]")
    if not isinstance(args, (list, tuple)):
        args = [args]
    if len(args) == 0:
        raise ValueError("Need at least one qubit index to perform permutation")
    qubits = [value_get(a) for a in args]
    _engine.validate_qubits(qubits, num_qubits)
    if len(qubits) != k:
        raise TypeError("Invalid gate size: %d-qubit matrix applied to %d qubits" % (k, len(qubits)))
    return LiftedOperator(matrix, qubits, num_qubits)


def resolve_gate(gate_set, defgate_set, pyquil_gate):
    """(matrix, qubit indices) of a pyQuil Gate; lookup order and ValueError as in reference
    tensor_gates (unitary_generator.py:340-362)."""
    if pyquil_gate.name in gate_set:
        table = gate_set
    elif pyquil_gate.name in defgate_set:
        table = defgate_set
    else:
        raise ValueError("Instruction (presumed a Gate or DefGate) is not found in standard gate set or "
                         "defined gate set of program!")
    entry = table[pyquil_gate.name]
    if pyquil_gate.params:
        entry = entry(*[value_get(p) for p in pyquil_gate.params])
    return np.asarray(entry), [value_get(q) for q in pyquil_gate.qubits]


def tensor_gates(gate_set, defgate_set, pyquil_gate, num_qubits):
    """The operator of a pyQuil gate instruction on the full register (descriptor form)."""
    matrix, qubits = resolve_gate(gate_set, defgate_set, pyquil_gate)
    return apply_gate(matrix, qubits, num_qubits)


def tensor_up(pauli_terms, num_qubits):
    """Dense matrix of a PauliSum (reference unitary_generator.py:366-411).  Host-side utility of
    the stabilizer QVM only; not on the gate-application path."""
    if not isinstance(pauli_terms, PauliSum):
        raise TypeError("can only tensor PauliSum")
    for term in pauli_terms.terms:
        if term._ops.keys() and max(term._ops.keys()) >= num_qubits:
            raise IndexError("pauli_terms has higher index than qubits")
    dim = 2 ** num_qubits
    total = np.zeros((dim, dim), dtype=complex)
    for term in pauli_terms.terms:
        acc = np.array([[1.0 + 0j]])
        for index in range(num_qubits):
            acc = np.kron(np.asarray(gate_matrix[term[index]], dtype=complex), acc)
        total += acc * term.coefficient
    return total
