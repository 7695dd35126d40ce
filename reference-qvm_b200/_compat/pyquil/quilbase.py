"""Instruction containers (pyquil 1.x ``quilbase`` look-alike, containers only)."""
import numpy as np  # noqa: F401  (reference unitary_generator.py:28 star-imports this module and uses ``np``)


class QuilAtom(object):
    def out(self):
        raise NotImplementedError

    def __str__(self):
        return self.out()

    def __repr__(self):
        return "<%s %s>" % (type(self).__name__, self.out())


class Qubit(QuilAtom):
    def __init__(self, index):
        if not (isinstance(index, (int, np.integer)) and index >= 0):
            raise TypeError("Qubit index must be a non-negative int")
        self.index = int(index)

    def out(self):
        return str(self.index)

    def __hash__(self):
        return hash(("q", self.index))

    def __eq__(self, other):
        return isinstance(other, Qubit) and other.index == self.index


class Addr(QuilAtom):
    def __init__(self, value):
        if not (isinstance(value, (int, np.integer)) and value >= 0):
            raise TypeError("Addr value must be a non-negative int")
        self.address = int(value)

    def out(self):
        return "[%d]" % self.address

    def __hash__(self):
        return hash(("a", self.address))

    def __eq__(self, other):
        return isinstance(other, Addr) and other.address == self.address


class Label(QuilAtom):
    def __init__(self, label_name):
        self.name = label_name

    def out(self):
        return "@%s" % self.name

    def __hash__(self):
        return hash(("l", self.name))

    def __eq__(self, other):
        return isinstance(other, Label) and other.name == self.name


def unpack_qubit(q):
    if isinstance(q, Qubit):
        return q
    if isinstance(q, (int, np.integer)):
        return Qubit(int(q))
    raise TypeError("qubit should be an int or Qubit, got %r" % (q,))


def unpack_classical_reg(c):
    if isinstance(c, (list, tuple)):
        if len(c) != 1:
            raise ValueError("classical register list must hold exactly one address")
        c = c[0]
    if isinstance(c, Addr):
        return c
    if isinstance(c, (int, np.integer)):
        return Addr(int(c))
    raise TypeError("classical register should be an int, [int] or Addr")


class AbstractInstruction(object):
    def out(self):
        raise NotImplementedError

    def __str__(self):
        return self.out()

    def __repr__(self):
        return "<%s %s>" % (type(self).__name__, self.out())


class Gate(AbstractInstruction):
    def __init__(self, name, params, qubits):
        if not isinstance(name, str):
            raise TypeError("Gate name must be a string")
        self.name = name
        self.params = list(params)
        self.qubits = [unpack_qubit(q) for q in qubits]

    def get_qubits(self):
        return set(q.index for q in self.qubits)

    def out(self):
        p = "(%s)" % ",".join(repr(x) for x in self.params) if self.params else ""
        return "%s%s %s" % (self.name, p, " ".join(q.out() for q in self.qubits))


class Measurement(AbstractInstruction):
    def __init__(self, qubit, classical_reg):
        self.qubit = unpack_qubit(qubit)
        self.classical_reg = None if classical_reg is None else unpack_classical_reg(classical_reg)

    def get_qubits(self):
        return set([self.qubit.index])

    def out(self):
        if self.classical_reg is None:
            return "MEASURE %s" % self.qubit.out()
        return "MEASURE %s %s" % (self.qubit.out(), self.classical_reg.out())


class SimpleInstruction(AbstractInstruction):
    op = None

    def get_qubits(self):
        return set()

    def out(self):
        return self.op


class Halt(SimpleInstruction):
    op = "HALT"


class Wait(SimpleInstruction):
    op = "WAIT"


class Reset(SimpleInstruction):
    op = "RESET"


class Nop(SimpleInstruction):
    op = "NOP"


class UnaryClassicalInstruction(AbstractInstruction):
    op = None

    def __init__(self, target):
        self.target = unpack_classical_reg(target)

    def get_qubits(self):
        return set()

    def out(self):
        return "%s %s" % (self.op, self.target.out())


class ClassicalTrue(UnaryClassicalInstruction):
    op = "TRUE"


class ClassicalFalse(UnaryClassicalInstruction):
    op = "FALSE"


class ClassicalNot(UnaryClassicalInstruction):
    op = "NOT"


class BinaryClassicalInstruction(AbstractInstruction):
    op = None

    def __init__(self, left, right):
        self.left = unpack_classical_reg(left)
        self.right = unpack_classical_reg(right)

    def get_qubits(self):
        return set()

    def out(self):
        return "%s %s %s" % (self.op, self.left.out(), self.right.out())


class ClassicalAnd(BinaryClassicalInstruction):
    op = "AND"


class ClassicalOr(BinaryClassicalInstruction):
    op = "OR"


class ClassicalMove(BinaryClassicalInstruction):
    op = "MOVE"


class ClassicalExchange(BinaryClassicalInstruction):
    op = "EXCHANGE"


class Jump(AbstractInstruction):
    def __init__(self, target):
        if not isinstance(target, Label):
            raise TypeError("target should be a Label")
        self.target = target

    def get_qubits(self):
        return set()

    def out(self):
        return "JUMP %s" % self.target.out()


class JumpTarget(AbstractInstruction):
    def __init__(self, label):
        if not isinstance(label, Label):
            raise TypeError("label must be a Label")
        self.label = label

    def get_qubits(self):
        return set()

    def out(self):
        return "LABEL %s" % self.label.out()


class JumpConditional(AbstractInstruction):
    op = None

    def __init__(self, target, condition):
        if not isinstance(target, Label):
            raise TypeError("target should be a Label")
        self.target = target
        self.condition = unpack_classical_reg(condition)

    def get_qubits(self):
        return set()

    def out(self):
        return "%s %s %s" % (self.op, self.target.out(), self.condition.out())


class JumpWhen(JumpConditional):
    op = "JUMP-WHEN"


class JumpUnless(JumpConditional):
    op = "JUMP-UNLESS"


class DefGate(AbstractInstruction):
    def __init__(self, name, matrix):
        if not isinstance(name, str):
            raise TypeError("Gate name must be a string")
        matrix = np.asarray(matrix)
        if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
            raise ValueError("Matrix must be square")
        rows = matrix.shape[0]
        if rows == 0 or rows & (rows - 1):
            raise ValueError("Dimension of matrix must be a power of 2")
        self.name = name
        self.matrix = matrix

    def get_constructor(self):
        return lambda *qubits: Gate(name=self.name, params=[], qubits=list(qubits))

    def num_args(self):
        return int(np.log2(self.matrix.shape[0]))

    def out(self):
        return "DEFGATE %s" % self.name
