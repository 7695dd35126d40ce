"""Expectation values <psi|O|psi> / tr(rho O) on the device.

The reference declares ``expectation(pyquil_program, operator_programs)`` on its QVMs and leaves it
``NotImplementedError`` (qvm_density.py:396-408, qvm_unitary.py:99-112); the operator its docstrings talk
about -- "list of PauliTerms" -- is the one ``tensor_up`` materialises (unitary_generator.py:366-411).  Here an
operator may be a ``PauliTerm``, a ``PauliSum``, or a pyQuil ``Program`` of gates (the pyquil 1.x QVM API):

* Pauli strings never touch the state: P|x> = i^ny (-1)^popcount(x & zmask) |x ^ xmask>, so <psi|P|psi> is ONE
  reduction pass (``qvm_pauli_expectation``) and tr(rho P) a gather of 2^n entries of rho
  (``qvm_density_offdiag``).
* Any other operator program O is applied to a copy of the state and <psi|O psi> is an inner-product pass
  (``qvm_inner_product``).
"""
import numpy as np

from pyquil.paulis import PauliSum, PauliTerm
from pyquil.quil import Program
from pyquil.quilbase import Gate

from referenceqvm.unitary_generator import value_get

_PAULI_NAMES = ("I", "X", "Y", "Z")


def pauli_ops_of_program(program):
    """{qubit: 'X'|'Y'|'Z'} when ``program`` is a product of single-qubit Pauli gates on distinct qubits, else
    None (the general path handles it)."""
    ops = {}
    for instruction in program:
        if not isinstance(instruction, Gate) or instruction.name not in _PAULI_NAMES or instruction.params:
            return None
        if len(instruction.qubits) != 1:
            return None
        q = value_get(instruction.qubits[0])
        if instruction.name == "I":
            continue
        if q in ops:
            return None
        ops[q] = instruction.name
    return ops


def masks(ops, position):
    """(xmask, zmask, number of Y factors) of a Pauli string {qubit: letter}; ``position(qubit)`` -> index bit."""
    xmask = zmask = ny = 0
    for q, letter in ops.items():
        bit = 1 << position(q)
        if letter in ("X", "Y"):
            xmask |= bit
        if letter in ("Z", "Y"):
            zmask |= bit
        ny += letter == "Y"
    return xmask, zmask, ny


def tidy(value):
    """A real number when the imaginary part is rounding noise (Hermitian operators), else the complex value."""
    value = complex(value)
    return value.real if abs(value.imag) <= 1e-13 * max(1.0, abs(value.real)) else value


def evaluate(operator, num_qubits, pauli, general):
    """Dispatch one operator: ``pauli(ops)`` -> value of a Pauli string {qubit: letter}; ``general(program)`` ->
    value of an arbitrary gate program."""
    if isinstance(operator, PauliTerm):
        operator = PauliSum([operator])
    if isinstance(operator, PauliSum):
        total = 0.0
        for term in operator.terms:
            ops = dict(term._ops)
            if ops and max(ops) >= num_qubits:
                raise IndexError("pauli_terms has higher index than qubits")     # as tensor_up (:393-396)
            total = total + term.coefficient * pauli(ops)
        return tidy(total)
    if isinstance(operator, Program):
        ops = pauli_ops_of_program(operator)
        if ops is not None:
            if ops and max(ops) >= num_qubits:
                raise ValueError("operator program acts on qubits outside the prepared register")
            return tidy(pauli(ops))
        return tidy(general(operator))
    raise TypeError("operators must be pyQuil Programs, PauliTerms or PauliSums")
