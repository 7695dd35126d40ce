"""``Program`` container (pyquil 1.x ``quil`` look-alike)."""
import itertools

import numpy as np

from pyquil.quilbase import (AbstractInstruction, Addr, DefGate, Gate, Jump, JumpTarget,
                             JumpUnless, JumpWhen, Label, Measurement, Qubit,
                             UnaryClassicalInstruction, BinaryClassicalInstruction,
                             unpack_classical_reg, unpack_qubit)

_label_counter = itertools.count(1)


def _fresh_label(prefix):
    return Label("%s%d" % (prefix, next(_label_counter)))


class Program(object):
    def __init__(self, *instructions):
        self._defined_gates = []
        self._instructions = []
        self.inst(*instructions)

    # -- container protocol ------------------------------------------------
    @property
    def defined_gates(self):
        return self._defined_gates

    @property
    def instructions(self):
        return list(self._instructions)

    def __iter__(self):
        return iter(self._instructions)

    def __len__(self):
        return len(self._instructions)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return Program(self._instructions[index])
        return self._instructions[index]

    def __add__(self, other):
        p = Program()
        p.inst(self)
        p.inst(other)
        return p

    def __iadd__(self, other):
        self.inst(other)
        return self

    # -- building ----------------------------------------------------------
    def inst(self, *instructions):
        for instruction in instructions:
            if isinstance(instruction, (list, itertools.chain)) or type(instruction).__name__ == "generator":
                self.inst(*instruction)
            elif isinstance(instruction, tuple):
                if len(instruction) == 0:
                    raise ValueError("tuple should have at least one element")
                if len(instruction) == 1:
                    self.inst(instruction[0])
                else:
                    op = instruction[0]
                    if op == "MEASURE":
                        if len(instruction) == 2:
                            self.measure(instruction[1], None)
                        else:
                            self.measure(instruction[1], instruction[2])
                    else:
                        params = []
                        possible_params = instruction[1]
                        rest = instruction[2:]
                        if isinstance(possible_params, list):
                            params = possible_params
                        else:
                            rest = [possible_params] + list(rest)
                        self.gate(op, params, rest)
            elif isinstance(instruction, Program):
                for dg in instruction._defined_gates:
                    if all(dg is not mine for mine in self._defined_gates):
                        self._defined_gates.append(dg)
                self._instructions.extend(instruction._instructions)
            elif isinstance(instruction, DefGate):
                self._defined_gates.append(instruction)
            elif isinstance(instruction, AbstractInstruction):
                self._instructions.append(instruction)
            elif isinstance(instruction, str):
                raise NotImplementedError("the pyquil stand-in has no Quil text parser")
            else:
                raise TypeError("Invalid instruction: %r" % (instruction,))
        return self

    def gate(self, name, params, qubits):
        return self.inst(Gate(name, params, [unpack_qubit(q) for q in qubits]))

    def defgate(self, name, matrix, parameters=None):
        if parameters:
            raise NotImplementedError("parametrised DefGates are not supported")
        return self.inst(DefGate(name, matrix))

    def measure(self, qubit_index, classical_reg):
        return self.inst(Measurement(unpack_qubit(qubit_index),
                                     None if classical_reg is None else unpack_classical_reg(classical_reg)))

    def measure_all(self, *qubit_reg_pairs):
        for q, c in qubit_reg_pairs:
            self.measure(q, c)
        return self

    def while_do(self, classical_reg, q_program):
        start, end = _fresh_label("START"), _fresh_label("END")
        self.inst(JumpTarget(start))
        self.inst(JumpUnless(target=end, condition=unpack_classical_reg(classical_reg)))
        self.inst(q_program)
        self.inst(Jump(start))
        self.inst(JumpTarget(end))
        return self

    def if_then(self, classical_reg, if_program, else_program=None):
        else_program = else_program if else_program is not None else Program()
        then_l, end_l = _fresh_label("THEN"), _fresh_label("END")
        self.inst(JumpWhen(target=then_l, condition=unpack_classical_reg(classical_reg)))
        self.inst(else_program)
        self.inst(Jump(end_l))
        self.inst(JumpTarget(then_l))
        self.inst(if_program)
        self.inst(JumpTarget(end_l))
        return self

    # -- queries -----------------------------------------------------------
    def get_qubits(self):
        qs = set()
        for instr in self._instructions:
            if isinstance(instr, (Gate, Measurement)):
                qs |= instr.get_qubits()
        return qs

    def out(self):
        return "\n".join(i.out() for i in self._defined_gates + self._instructions) + "\n"

    def __str__(self):
        return self.out()


def get_classical_addresses_from_program(program):
    addresses = set()
    for instr in program:
        if isinstance(instr, Measurement) and instr.classical_reg is not None:
            addresses.add(instr.classical_reg.address)
        elif isinstance(instr, UnaryClassicalInstruction):
            addresses.add(instr.target.address)
        elif isinstance(instr, BinaryClassicalInstruction):
            addresses.add(instr.left.address)
            addresses.add(instr.right.address)
    return sorted(addresses)


def merge_programs(prog_list):
    out = Program()
    for p in prog_list:
        out.inst(p)
    return out
