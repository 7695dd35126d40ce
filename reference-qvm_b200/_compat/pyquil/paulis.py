"""Pauli algebra containers + ``exponentiate`` (pyquil 1.x ``paulis`` reconstruction, from memory).

Only reference code *off* the gate-application path uses this (``tensor_up``, the stabilizer QVM,
and one norm-only test); it is kept small.  Results that depend on it are shim-dependent.
"""
from collections import OrderedDict
from numbers import Number

import numpy as np

from pyquil.gates import CNOT, H, PHASE, RX, RZ, X
from pyquil.quil import Program

PAULI_OPS = ["X", "Y", "Z", "I"]
_PROD = {"ZZ": "I", "YY": "I", "XX": "I", "II": "I", "XY": "Z", "XZ": "Y", "YX": "Z", "YZ": "X",
         "ZX": "Y", "ZY": "X", "IX": "X", "IY": "Y", "IZ": "Z", "ZI": "Z", "YI": "Y", "XI": "X"}
_COEF = {"ZZ": 1.0, "YY": 1.0, "XX": 1.0, "II": 1.0, "XY": 1.0j, "XZ": -1.0j, "YX": -1.0j,
         "YZ": 1.0j, "ZX": 1.0j, "ZY": -1.0j, "IX": 1.0, "IY": 1.0, "IZ": 1.0, "ZI": 1.0,
         "YI": 1.0, "XI": 1.0}


class PauliTerm(object):
    def __init__(self, op, index, coefficient=1.0):
        if op not in PAULI_OPS:
            raise ValueError("%s is not a valid Pauli operator" % op)
        self._ops = OrderedDict()
        if op != "I":
            self._ops[index] = op
        self.coefficient = complex(coefficient)

    def id(self):
        return "".join("%s%s" % (self._ops[q], q) for q in sorted(self._ops))

    def __len__(self):
        return len(self._ops)

    def __getitem__(self, i):
        return self._ops.get(i, "I")

    def __iter__(self):
        for i in self.get_qubits():
            yield i, self[i]

    def get_qubits(self):
        return list(self._ops.keys())

    def copy(self):
        t = PauliTerm("I", 0, self.coefficient)
        t._ops = OrderedDict(self._ops)
        return t

    def __mul__(self, other):
        if isinstance(other, Number):
            t = self.copy()
            t.coefficient *= other
            return t
        if isinstance(other, PauliSum):
            return (PauliSum([self]) * other).simplify()
        t = self.copy()
        t.coefficient *= other.coefficient
        for q, op in other._ops.items():
            mine = t._ops.get(q, "I")
            key = mine + op
            t.coefficient *= _COEF[key]
            new = _PROD[key]
            if new == "I":
                t._ops.pop(q, None)
            else:
                t._ops[q] = new
        return t

    __rmul__ = lambda self, other: self * other

    def __add__(self, other):
        if isinstance(other, Number):
            return self + PauliTerm("I", 0, other)
        if isinstance(other, PauliSum):
            return (PauliSum([self]) + other).simplify()
        return PauliSum([self, other]).simplify()

    __radd__ = lambda self, other: self + other

    def __sub__(self, other):
        return self + (-1.0 * other)

    def __str__(self):
        return "%s*%s" % (self.coefficient, self.id() or "I")


class PauliSum(object):
    def __init__(self, terms):
        if not (isinstance(terms, (list, tuple)) and all(isinstance(t, PauliTerm) for t in terms)):
            raise ValueError("PauliSum's are currently constructed from Sequences of PauliTerms.")
        self.terms = list(terms) if len(terms) else [0.0 * ID()]

    def __len__(self):
        return len(self.terms)

    def __getitem__(self, i):
        return self.terms[i]

    def __iter__(self):
        return iter(self.terms)

    def __add__(self, other):
        if isinstance(other, Number):
            other = PauliTerm("I", 0, other)
        if isinstance(other, PauliTerm):
            other = PauliSum([other])
        return PauliSum(self.terms + other.terms).simplify()

    __radd__ = lambda self, other: self + other

    def __mul__(self, other):
        if isinstance(other, (Number, PauliTerm)):
            others = [other]
        else:
            others = other.terms
        return PauliSum([a * b for a in self.terms for b in others]).simplify()

    __rmul__ = lambda self, other: self * other

    def __sub__(self, other):
        return self + (-1.0 * other)

    def get_qubits(self):
        return sorted(set(q for t in self.terms for q in t.get_qubits()))

    def simplify(self):
        acc = OrderedDict()
        for t in self.terms:
            key = t.id()
            if key in acc:
                acc[key].coefficient += t.coefficient
            else:
                acc[key] = t.copy()
        kept = [t for t in acc.values() if not np.isclose(t.coefficient, 0.0)]
        return PauliSum(kept) if kept else PauliSum([0.0 * ID()])


def ID():
    return PauliTerm("I", 0, 1.0)


def sI(q=0):
    return PauliTerm("I", q)


def sX(q):
    return PauliTerm("X", q)


def sY(q):
    return PauliTerm("Y", q)


def sZ(q):
    return PauliTerm("Z", q)


def exponential_map(term):
    """Return f(param) -> Program implementing exp(-1j * param * term)."""
    if not np.isclose(np.imag(term.coefficient), 0.0):
        raise TypeError("PauliTerm coefficient must be real")
    coeff = term.coefficient.real

    def exp_wrap(param):
        prog = Program()
        if len(term) == 0:
            prog.inst(X(0), PHASE(-param * coeff, 0), X(0), PHASE(-param * coeff, 0))
            return prog
        qubits = sorted(term.get_qubits())
        undo = Program()
        for q in qubits:
            op = term[q]
            if op == "X":
                prog.inst(H(q))
                undo.inst(H(q))
            elif op == "Y":
                prog.inst(RX(np.pi / 2.0, q))
                undo.inst(RX(-np.pi / 2.0, q))
        ladder = [CNOT(a, b) for a, b in zip(qubits[:-1], qubits[1:])]
        prog.inst(ladder)
        prog.inst(RZ(2.0 * coeff * param, qubits[-1]))
        prog.inst(list(reversed(ladder)))
        prog.inst(undo)
        return prog
    return exp_wrap


def exponentiate(term):
    return exponential_map(term)(1.0)
