"""Compiled programs: the gate prefix of a pyQuil program, classified / fused / planned / encoded ONCE.

The reference re-walks the program and rebuilds every lifted operator on every call
(``qam.py:61-107`` load_program, ``:168-181`` kernel, ``unitary_generator.py:326-363`` per gate).  Here the
first run of a program records the launches its leading run of Gate instructions turns into
(``qvm_record_begin`` / ``qvm_record_end``); later runs of the same Program object replay them with one
library call (``qvm_plan_run``: specialised kernels, captured into a CUDA graph once all of them exist) and
the state machine resumes at the first instruction that is not a Gate (MEASURE, classical, jumps).

The cache lives on the Program object and is validated by a cheap token -- length, identity of the first
and last instruction, number of DefGates -- so appending to a program (``inst``, ``measure``, ``+=``)
invalidates it.  Instruction objects edited in place (not something the pyQuil API offers) would not be
noticed: ``QVM_B200_PLAN_CACHE=0`` or ``qvm.plan_cache = False`` turns the cache off.
"""
import os

from pyquil.quilbase import Gate

ENABLED = os.environ.get("QVM_B200_PLAN_CACHE", "1") != "0"
_ATTR = "_qvm_b200_plans"
MAX_PLANS_PER_PROGRAM = 8


class Entry(object):
    __slots__ = ("token", "k", "compiled", "virt", "num_qubits", "n_cbits", "defgate_set", "gates_only")

    def __init__(self, token, k, compiled, virt, num_qubits, n_cbits, defgate_set, gates_only):
        self.token, self.k, self.compiled, self.virt = token, k, compiled, virt
        self.num_qubits, self.n_cbits, self.defgate_set, self.gates_only = num_qubits, n_cbits, defgate_set, gates_only


def token_of(program):
    n = len(program)
    if n == 0:
        return (0, None, None, len(program.defined_gates))
    return (n, id(program[0]), id(program[n - 1]), len(program.defined_gates))


def lookup(program, key):
    table = getattr(program, _ATTR, None)
    if not table:
        return None
    entry = table.get(key)
    if entry is None:
        return None
    if entry.token != token_of(program):
        del table[key]
        return None
    return entry


def store(program, key, entry):
    try:
        table = program.__dict__.setdefault(_ATTR, {})
    except AttributeError:          # a Program class without a __dict__: no caching
        return
    if len(table) >= MAX_PLANS_PER_PROGRAM:
        table.clear()
    table[key] = entry


def gate_prefix_length(program):
    k = 0
    for instruction in program:
        if not isinstance(instruction, Gate):
            break
        k += 1
    return k


def uses_base_hooks(obj, base, names):
    """True when ``obj``'s class overrides none of ``names`` (a subclass with its own hooks or transitions is
    stepped instruction by instruction, as in the reference)."""
    cls = type(obj)
    return all(getattr(cls, name) is getattr(base, name) for name in names)
