"""Helpers shared by the tests: golden-fixture loading and (de)serialised Quil programs."""
import json
import os

import numpy as np

import pyquil.gates as pg
from pyquil.quil import Program
from pyquil.quilbase import (ClassicalAnd, ClassicalExchange, ClassicalFalse, ClassicalMove, ClassicalNot,
                             ClassicalOr, ClassicalTrue, Gate, Halt, Jump, JumpTarget, JumpUnless,
                             JumpWhen, Label, Measurement, Nop, Wait)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_UNARY = {"TRUE": ClassicalTrue, "FALSE": ClassicalFalse, "NOT": ClassicalNot}
_BINARY = {"AND": ClassicalAnd, "OR": ClassicalOr, "MOVE": ClassicalMove, "EXCHANGE": ClassicalExchange}


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as fh:
        meta = json.load(fh)
    return meta, np.load(os.path.join(GOLDEN, name + ".npz"))


def program_from_json(blob):
    """Inverse of ``ser_program`` in tests/golden/make_golden.py."""
    prog = Program()
    for d in blob["defgates"]:
        prog.defgate(d["name"], np.array(d["re"]) + 1j * np.array(d["im"]))
    for ins in blob["instructions"]:
        op = ins["op"]
        if op == "gate":
            prog.inst(Gate(ins["name"], ins["params"], ins["qubits"]))
        elif op == "measure":
            prog.inst(Measurement(ins["q"], ins["c"]))
        elif op in _UNARY:
            prog.inst(_UNARY[op](ins["c"]))
        elif op in _BINARY:
            prog.inst(_BINARY[op](ins["l"], ins["r"]))
        elif op == "LABEL":
            prog.inst(JumpTarget(Label(ins["label"])))
        elif op == "JUMP":
            prog.inst(Jump(Label(ins["label"])))
        elif op == "JUMP-WHEN":
            prog.inst(JumpWhen(Label(ins["label"]), ins["c"]))
        elif op == "JUMP-UNLESS":
            prog.inst(JumpUnless(Label(ins["label"]), ins["c"]))
        elif op == "HALT":
            prog.inst(Halt())
        elif op == "NOP":
            prog.inst(Nop())
        elif op == "WAIT":
            prog.inst(Wait())
        else:
            raise ValueError(op)
    return prog


def spec_program(spec):
    """[(name, params, qubits)] -> Program."""
    prog = Program()
    for name, params, qubits in spec:
        prog.inst(Gate(name, list(params), list(qubits)))
    return prog


def oracle_ops(blob_or_program, gate_set=None):
    """Straight-line program -> oracle op list [("gate", matrix, qubits) | ("measure", q, c)].

    Only valid for programs without control flow (the oracle has no program counter)."""
    from referenceqvm.gates import gate_matrix
    gate_set = gate_matrix if gate_set is None else gate_set
    prog = program_from_json(blob_or_program) if isinstance(blob_or_program, dict) else blob_or_program
    defs = {d.name: d.matrix for d in prog.defined_gates}
    ops = []
    for ins in prog:
        if isinstance(ins, Gate):
            m = gate_set[ins.name] if ins.name in gate_set else defs[ins.name]
            if ins.params:
                m = m(*ins.params)
            ops.append(("gate", np.asarray(m, dtype=np.complex128), [q.index for q in ins.qubits]))
        elif isinstance(ins, Measurement):
            ops.append(("measure", ins.qubit.index, ins.classical_reg.address))
        else:
            raise ValueError("control flow not representable as an oracle op list: %r" % (ins,))
    return ops


def straight_line(blob):
    return all(i["op"] in ("gate", "measure") for i in blob["instructions"])
