"""Quantum abstract machine: program loading, bit accounting, the fetch/execute loop and the
classical half of the instruction set.

Host-side driver of the gate-application path; behaviour follows reference ``referenceqvm/qam.py``
(load_program :61-107, identify_bits :120-157 incl. the 512-classical-bit floor :104-105 and the
51-qubit cap :152-156, kernel :168-181, find_label :183-201).  The classical / control-flow
transitions the reference spells out twice (qvm_wavefunction.py:164-236 and
qvm_density.py:168-240) live here once, in :meth:`QAM._classical_transition`.
"""
import numpy as np

from pyquil.quil import Program
from pyquil.quilbase import (BinaryClassicalInstruction, ClassicalAnd, ClassicalExchange, ClassicalFalse,
                             ClassicalMove, ClassicalNot, ClassicalOr, ClassicalTrue, Gate, Halt, Jump,
                             JumpConditional, JumpTarget, JumpUnless, JumpWhen, Label, Measurement,
                             UnaryClassicalInstruction)

from referenceqvm.unitary_generator import value_get

Q_LIMIT = 51
CBIT_FLOOR = 512


class QAM(object):
    """State-machine model: classical memory, quantum memory, program, program counter, gate sets."""

    def __init__(self, num_qubits=None, program=None, program_counter=None,
                 classical_memory=None, gate_set=None, defgate_set=None):
        self.num_qubits = num_qubits
        self.classical_memory = classical_memory
        self.program = program
        self.program_counter = program_counter
        self.gate_set = gate_set
        self.defgate_set = defgate_set
        self.all_inst = None

    # -- loading -----------------------------------------------------------------------------
    def load_program(self, pyquil_program):
        if not isinstance(pyquil_program, Program):
            raise TypeError("I can only generate from pyQuil programs")
        if self.all_inst is None:
            raise NotImplementedError("QAM needs to be subclassed in order to load program")

        self.defgate_set = dict((dg.name, dg.matrix) for dg in pyquil_program.defined_gates)

        gates_only = all(isinstance(instr, Gate) and
                         (instr.name in self.gate_set or instr.name in self.defgate_set)
                         for instr in pyquil_program)
        if not gates_only and self.all_inst is False:
            raise TypeError("Some gates used are not allowed in this QAM")

        self.program = pyquil_program
        self.program_counter = 0
        self._allocate_bits()

    def memory_reset(self):
        if self.program is None:
            raise TypeError("Program must be loaded to call reset")
        self._allocate_bits()

    def _allocate_bits(self):
        q_max, c_max = self.identify_bits()
        self.num_qubits = q_max
        self.classical_memory = np.zeros(max(c_max, CBIT_FLOOR)).astype(bool)

    def identify_bits(self):
        """(number of qubits, number of classical bits) the loaded program needs.

        Keeps the reference's accounting quirk: a MEASURE whose qubit raises the qubit count does
        not also raise the classical count in the same step (``elif`` at qam.py:133-136), likewise
        the left/right operands of binary classical instructions (:141-144)."""
        q_top, c_top = -1, -1
        for inst in self.program:
            if isinstance(inst, Gate):                      # (first: nearly every instruction is a gate)
                for q in inst.qubits:
                    v = value_get(q)
                    if v > q_top:
                        q_top = v
            elif isinstance(inst, Measurement):
                if value_get(inst.qubit) > q_top:
                    q_top = value_get(inst.qubit)
                elif value_get(inst.classical_reg) > c_top:
                    c_top = value_get(inst.classical_reg)
            elif isinstance(inst, UnaryClassicalInstruction):
                c_top = max(c_top, value_get(inst.target))
            elif isinstance(inst, BinaryClassicalInstruction):
                if value_get(inst.left) > c_top:
                    c_top = value_get(inst.left)
                elif value_get(inst.right) > c_top:
                    c_top = value_get(inst.right)
        if q_top + 1 > Q_LIMIT:
            raise RuntimeError("Too many qubits. Maximum qubit number supported: {}".format(Q_LIMIT))
        return (q_top + 1, c_top + 1)

    # -- execution ---------------------------------------------------------------------------
    def current_instruction(self):
        return self.program[self.program_counter]

    def kernel(self):
        """Fetch/execute until the counter runs off the program or a transition reports HALT."""
        while self.program_counter < len(self.program):
            if self.transition(self.current_instruction()):
                break

    def find_label(self, label):
        assert isinstance(label, Label)
        for index, action in enumerate(self.program):
            if isinstance(action, JumpTarget) and label == action.label:
                return index
        raise RuntimeError("Improper program - Jump Target not found in the input program!")

    def _classical_transition(self, instruction):
        """Everything that is not a Gate or a MEASURE.  Returns False if the instruction is not a
        supported classical / control-flow instruction."""
        mem = self.classical_memory
        if isinstance(instruction, Jump):
            self.program_counter = self.find_label(instruction.target)
        elif isinstance(instruction, JumpTarget):
            self.program_counter += 1
        elif isinstance(instruction, JumpConditional):
            cond = bool(mem[value_get(instruction.condition)])
            dest = self.find_label(instruction.target)
            if isinstance(instruction, JumpWhen):
                taken = cond
            elif isinstance(instruction, JumpUnless):
                taken = not cond
            else:
                raise TypeError("Invalid JumpConditional")
            self.program_counter = dest if taken else self.program_counter + 1
        elif isinstance(instruction, UnaryClassicalInstruction):
            target = value_get(instruction.target)
            if isinstance(instruction, ClassicalTrue):
                mem[target] = True
            elif isinstance(instruction, ClassicalFalse):
                mem[target] = False
            elif isinstance(instruction, ClassicalNot):
                mem[target] = not mem[target]
            else:
                raise TypeError("Invalid UnaryClassicalInstruction")
            self.program_counter += 1
        elif isinstance(instruction, BinaryClassicalInstruction):
            left, right = value_get(instruction.left), value_get(instruction.right)
            lv, rv = mem[left], mem[right]
            if isinstance(instruction, ClassicalAnd):
                mem[right] = lv & rv
            elif isinstance(instruction, ClassicalOr):
                mem[right] = lv | rv
            elif isinstance(instruction, ClassicalMove):
                mem[right] = lv
            elif isinstance(instruction, ClassicalExchange):
                mem[left], mem[right] = rv, lv
            else:
                raise TypeError("Invalid BinaryClassicalInstruction")
            self.program_counter += 1
        elif isinstance(instruction, Halt):
            self.program_counter = len(self.program)
        else:
            return False
        return True

    @staticmethod
    def _unsupported(instruction):
        return TypeError("Invalid / unsupported instruction type: {}\nCurrently supported: unary/binary "
                         "classical instructions, control flow (if/while/jumps), measurements, and "
                         "gates/defgates.".format(type(instruction)))

    def transition(self, instruction):
        raise NotImplementedError("transition is an abstract class of QAM. Implement in subclass")

    def wavefunction(self, pyquil_program, classical_addresses=None):
        raise NotImplementedError("wavefunction is an abstract class of QAM. Implement in subclass")

    def unitary(self, pyquil_program):
        raise NotImplementedError("unitary is an abstract class of QAM. Implement in subclass")
