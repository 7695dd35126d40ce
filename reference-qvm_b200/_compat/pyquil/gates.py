"""Gate / classical-instruction constructors (pyquil 1.x ``gates`` look-alike)."""
from pyquil.quilbase import (ClassicalAnd, ClassicalExchange, ClassicalFalse, ClassicalMove,
                             ClassicalNot, ClassicalOr, ClassicalTrue, Gate, Halt, Measurement,
                             Nop, Reset, Wait, unpack_classical_reg, unpack_qubit)


def _fixed(name, arity):
    def ctor(*qubits):
        if len(qubits) != arity:
            raise TypeError("%s takes %d qubit(s)" % (name, arity))
        return Gate(name=name, params=[], qubits=[unpack_qubit(q) for q in qubits])
    ctor.__name__ = name
    return ctor


def _param(name, arity, nparams=1):
    def ctor(*args):
        if len(args) != arity + nparams:
            raise TypeError("%s takes %d parameter(s) and %d qubit(s)" % (name, nparams, arity))
        return Gate(name=name, params=list(args[:nparams]),
                    qubits=[unpack_qubit(q) for q in args[nparams:]])
    ctor.__name__ = name
    return ctor


I = _fixed("I", 1)
X = _fixed("X", 1)
Y = _fixed("Y", 1)
Z = _fixed("Z", 1)
H = _fixed("H", 1)
S = _fixed("S", 1)
T = _fixed("T", 1)
RX = _param("RX", 1)
RY = _param("RY", 1)
RZ = _param("RZ", 1)
PHASE = _param("PHASE", 1)
CZ = _fixed("CZ", 2)
CNOT = _fixed("CNOT", 2)
CCNOT = _fixed("CCNOT", 3)
CPHASE00 = _param("CPHASE00", 2)
CPHASE01 = _param("CPHASE01", 2)
CPHASE10 = _param("CPHASE10", 2)
CPHASE = _param("CPHASE", 2)
SWAP = _fixed("SWAP", 2)
CSWAP = _fixed("CSWAP", 3)
ISWAP = _fixed("ISWAP", 2)
PSWAP = _param("PSWAP", 2)

WAIT = Wait()
RESET = Reset()
NOP = Nop()
HALT = Halt()


def MEASURE(qubit, classical_reg=None):
    return Measurement(unpack_qubit(qubit),
                       None if classical_reg is None else unpack_classical_reg(classical_reg))


def TRUE(classical_reg):
    return ClassicalTrue(unpack_classical_reg(classical_reg))


def FALSE(classical_reg):
    return ClassicalFalse(unpack_classical_reg(classical_reg))


def NOT(classical_reg):
    return ClassicalNot(unpack_classical_reg(classical_reg))


def AND(classical_reg1, classical_reg2):
    return ClassicalAnd(unpack_classical_reg(classical_reg1), unpack_classical_reg(classical_reg2))


def OR(classical_reg1, classical_reg2):
    return ClassicalOr(unpack_classical_reg(classical_reg1), unpack_classical_reg(classical_reg2))


def MOVE(classical_reg1, classical_reg2):
    return ClassicalMove(unpack_classical_reg(classical_reg1), unpack_classical_reg(classical_reg2))


def EXCHANGE(classical_reg1, classical_reg2):
    return ClassicalExchange(unpack_classical_reg(classical_reg1), unpack_classical_reg(classical_reg2))


STANDARD_GATES = {name: globals()[name] for name in (
    "I X Y Z H S T RX RY RZ PHASE CZ CNOT CCNOT CPHASE00 CPHASE01 CPHASE10 CPHASE "
    "SWAP CSWAP ISWAP PSWAP").split()}
