"""Wavefunction QVM whose state lives in HBM.

Public surface and semantics follow reference ``referenceqvm/qvm_wavefunction.py``:
``wavefunction`` (:323-367), ``run`` (:266-285), ``run_and_measure`` (:287-321), ``measurement``
(:96-129), ``transition`` with ``pre``/``post`` hooks (:238-264).  What changed is where the work
happens:

* Gate branch (:158-162): instead of ``unitary = tensor_gates(...); self.wf = unitary.dot(self.wf)``
  the gate is classified and queued; queued gates are fused into one-HBM-pass groups and run by
  the sm_100a tile kernel when the amplitudes are next needed.
* Measurement branch (:147-156): one cooperative kernel (reduce, decide, collapse, renormalise).
  The uniform is drawn on the host from ``np.random.random()`` exactly once per MEASURE, in program
  order, so seeded programs reproduce the reference's classical memory draw for draw.
* ``run`` / ``run_and_measure``: the reference re-simulates the whole program per trial.  When
  that is provably equivalent to sampling one final state (no MEASURE in the program, or a purely
  terminal block of MEASUREs and no control flow) the program is simulated once and the trials
  are drawn by the prefix-sum + binary-search kernel; otherwise the per-trial loop is kept.
  The sampling fast path draws ONE uniform per trial (the reference: one per MEASURE per trial), so
  its results are distributed like the reference's but are not its seeded stream, and ``self.wf``
  afterwards is the un-collapsed final state; draw-for-draw parity with a seeded reference holds on
  the per-trial and batched paths (``sampling_fast_path = False`` forces them).
"""
import gc
import os
import time

import numpy as np

from pyquil.quil import Program  # noqa: F401  (kept for parity with the reference's namespace)
from pyquil.quilbase import (BinaryClassicalInstruction, ClassicalAnd, ClassicalExchange, ClassicalFalse,
                             ClassicalMove, ClassicalNot, ClassicalOr, ClassicalTrue, Gate, Halt, Jump,
                             JumpConditional, JumpTarget, JumpWhen, Measurement, UnaryClassicalInstruction)
from pyquil.wavefunction import Wavefunction

from referenceqvm import _compiled, _engine
from referenceqvm.gates import utility_gates  # noqa: F401
from referenceqvm.qam import QAM
from referenceqvm.unitary_generator import lifted_gate, resolve_gate, tensor_gates, value_get  # noqa: F401

# Above this many qubits ``wavefunction()`` hands back a lazy, sliceable view of the device
# amplitudes instead of one host ndarray (2^31 amplitudes = 32 GiB of host memory).
LAZY_QUBITS = int(os.environ.get("QVM_B200_LAZY_QUBITS", "30"))


def _index_bits(indices, num_qubits=63):
    """(trials, >= num_qubits + 1) uint8 matrix of the little-endian bits of sampled basis-state
    indices; only the bytes that can hold a set bit are unpacked, column ``num_qubits`` is always 0."""
    by = np.ascontiguousarray(indices, dtype="<u8").view(np.uint8).reshape(-1, 8)
    return np.unpackbits(by[:, :num_qubits // 8 + 1], axis=1, bitorder="little")


_frozen_rows = [False]
_GC_FREEZE = os.environ.get("QVM_B200_GC_FREEZE", "1") != "0"


def _thaw_rows():
    """Undo the ``gc.freeze()`` of the previous large result (see ``_rows_to_lists``)."""
    if _frozen_rows[0]:
        _frozen_rows[0] = False
        gc.unfreeze()


_pyrows = [None]


def _pyrows_lib():
    """referenceqvm/_pyrows.so (csrc/qvm_pyrows.c, built by build.py): fills the per-trial lists in C."""
    if _pyrows[0] is None:
        import ctypes
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pyrows.so")
        try:
            lib = ctypes.PyDLL(path)
            lib.qvm_rows_from_indices.restype = ctypes.py_object
            lib.qvm_rows_from_indices.argtypes = [ctypes.c_void_p, ctypes.c_ssize_t, ctypes.c_void_p, ctypes.c_ssize_t]
            lib.qvm_rows_from_int64.restype = ctypes.py_object
            lib.qvm_rows_from_int64.argtypes = [ctypes.c_void_p] + [ctypes.c_ssize_t] * 4
            _pyrows[0] = lib
        except (OSError, AttributeError):
            _pyrows[0] = False           # host-side formatting only: numpy's tolist() does the same, slower
    return _pyrows[0]


def _rows_to_lists(matrix=None, indices=None, sel=None):
    """One Python list per trial (the reference's return type) from an integer matrix, or directly from sampled
    basis-state ``indices`` and the bit position ``sel[c]`` each output column reads (outside 0..63: always 0) --
    without paying the cyclic garbage collector for it.  The C helper takes its rows off the collector's books as it
    builds them; the rest of this paragraph is about numpy's ``tolist()``, the fallback when ``_pyrows.so`` is
    missing.  A million row lists are a million tracked containers:
    built with the collector running, the generational collections they trigger cost a third of the conversion;
    built with it paused, the first allocation after it is switched back on starts a full collection that walks
    all of them (0.15 s per million rows on the GPU box, charged to whoever allocates next).  Rows cannot be part
    of a cycle, so for large results the collector is told to leave them alone: ``gc.freeze()`` moves everything
    allocated so far to the permanent generation -- until the next call into the QVM, which thaws
    (``gc.unfreeze()``) whatever is still alive.  ``QVM_B200_GC_FREEZE=0`` turns this off."""
    _thaw_rows()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        lib = _pyrows_lib()
        if indices is not None:
            idx = np.ascontiguousarray(indices, dtype=np.uint64)
            if lib:
                cols = np.ascontiguousarray(sel, dtype=np.int32)
                out = lib.qvm_rows_from_indices(idx.ctypes.data, idx.shape[0], cols.ctypes.data, cols.shape[0])
            else:
                picked = np.take(_index_bits(idx, 63), [c if 0 <= c < 64 else 0 for c in sel], axis=1)
                picked[:, [i for i, c in enumerate(sel) if not 0 <= c < 64]] = 0
                out = picked.tolist()
        else:
            m = np.asarray(matrix, dtype=np.int64)
            if lib and m.ndim == 2:
                out = lib.qvm_rows_from_int64(m.ctypes.data, m.shape[0], m.shape[1], m.strides[0], m.strides[1])
            else:
                out = m.tolist()
        # rows built by the C helper are untracked (PyObject_GC_UnTrack: a list of ints cannot be in a cycle), so
        # the collector never sees them; only tolist()'s rows need the freeze
        if not lib and _GC_FREEZE and was_enabled and len(out) >= 100000:
            gc.freeze()
            _frozen_rows[0] = True
        return out
    finally:
        if was_enabled:
            gc.enable()


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def _is_swap(matrix, qubits):
    """Exactly the SWAP matrix on two qubits (three scalar probes first: the full comparison costs more
    than everything else the host does for a gate)."""
    return (len(qubits) == 2 and matrix.shape == (4, 4) and matrix[1, 2] == 1 and matrix[1, 1] == 0
            and matrix[0, 0] == 1 and np.array_equal(matrix, _SWAP))


class QVM_Wavefunction(QAM):
    """Supports run(), run_and_measure() and wavefunction()."""

    def __init__(self, num_qubits=None, program=None, program_counter=None,
                 classical_memory=None, gate_set=None, defgate_set=None):
        super(QVM_Wavefunction, self).__init__(num_qubits=num_qubits, program=program,
                                               program_counter=program_counter,
                                               classical_memory=classical_memory,
                                               gate_set=gate_set, defgate_set=defgate_set)
        self.all_inst = True
        self._engine = None
        self.device = int(os.environ.get("QVM_B200_DEVICE", "0"))
        #: set False to force the reference's per-trial re-simulation in run()/run_and_measure()
        self.sampling_fast_path = True
        self.last_timing = {}
        #: compiled programs (referenceqvm/_compiled.py): the leading run of Gate instructions of a Program is
        #: planned and encoded once and replayed on later runs of the same Program object
        self.plan_cache = _compiled.ENABLED
        #: SWAP gates are not executed: they exchange two entries of this program-qubit -> engine-qubit
        #: table, and later instructions are routed through it.  The table is undone (with real SWAPs)
        #: only when amplitudes are read back; MEASURE and sampling just look bits up through it.
        self._virt = []

    # -- device state --------------------------------------------------------------------------
    def _fresh_engine(self, reset=True):
        """|0...0> on the device, re-using the allocation when the register size is unchanged."""
        if self._engine is not None and self._engine.n != self.num_qubits:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine = self._make_engine(self.num_qubits)
        if reset:
            self._engine.reset()
        self._virt = list(range(self.num_qubits))
        return self._engine

    # -- compiled programs ---------------------------------------------------------------------------
    _HOOKS = ("pre", "post", "transition", "_transition", "measurement", "kernel", "load_program")

    def _plan_key(self):
        return ("wavefunction", id(self.gate_set), self.device)

    def _start(self, pyquil_program):
        """load_program + |0...0> + the program's leading run of Gate instructions, leaving the program
        counter on the first instruction that is not a Gate.  The first time a Program object comes by, that
        gate prefix is recorded while it runs; afterwards it is replayed with one library call and neither
        the program nor its gates are looked at again (referenceqvm/_compiled.py)."""
        if not isinstance(pyquil_program, Program):
            raise TypeError("I can only generate from pyQuil programs")
        _thaw_rows()
        cache = self.plan_cache and _compiled.uses_base_hooks(self, QVM_Wavefunction, self._HOOKS)
        entry = _compiled.lookup(pyquil_program, self._plan_key()) if cache else None
        if entry is not None and (self._engine is None or isinstance(self._engine, _engine.Engine)):
            self.defgate_set = entry.defgate_set
            self.program = pyquil_program
            self.num_qubits = entry.num_qubits
            self.classical_memory = np.zeros(entry.n_cbits, dtype=bool)
            eng = self._fresh_engine(reset=False)
            if isinstance(eng, _engine.Engine):
                eng.replay(entry.compiled, reset_first=True)
                self._virt = list(entry.virt)
                self.program_counter = entry.k
                return
            eng.reset()
            self.program_counter = 0
            return
        self.load_program(pyquil_program)
        eng = self._fresh_engine()
        k = _compiled.gate_prefix_length(pyquil_program) if cache and isinstance(eng, _engine.Engine) else 0
        if k == 0:
            return
        eng.begin_recording()
        try:
            while self.program_counter < k:
                self.transition(self.current_instruction())
        except Exception:
            eng.end_recording(keep=False)
            raise
        compiled = eng.end_recording()
        _compiled.store(pyquil_program, self._plan_key(),
                        _compiled.Entry(_compiled.token_of(pyquil_program), k, compiled, list(self._virt), self.num_qubits,
                                        len(self.classical_memory), self.defgate_set, True))

    def _materialise_swaps(self):
        """Bring the engine state back to program-qubit order (real SWAP gates, then identity table)."""
        virt = self._virt
        if virt == list(range(len(virt))):
            return
        # engine qubit virt[q] currently plays program qubit q; sort by transpositions
        holder = {e: q for q, e in enumerate(virt)}          # engine qubit -> program qubit it plays
        for q in range(len(virt)):
            e = virt[q]
            if e == q:
                continue
            other = holder[q]                                 # program qubit currently played by engine qubit q
            self._engine.queue_matrix(_SWAP, [e, q])          # after this engine qubit q plays program qubit q
            virt[q], virt[other] = q, e
            holder[q], holder[e] = q, other
        assert virt == list(range(len(virt)))

    def _make_engine(self, n):
        return _engine.make_engine(n, device=self.device)

    @property
    def wf(self):
        """Host copy of the amplitudes (the reference keeps ``self.wf`` as an ndarray)."""
        if self._engine is None:
            return None
        self._materialise_swaps()
        return self._engine.amplitudes()

    @wf.setter
    def wf(self, amplitudes):
        if amplitudes is None:
            if self._engine is not None:
                self._engine.close()
            self._engine = None
            return
        amplitudes = np.asarray(amplitudes).reshape(-1)
        n = int(amplitudes.size).bit_length() - 1
        if self._engine is None or self._engine.n != n:
            if self._engine is not None:
                self._engine.close()
            self._engine = self._make_engine(n)
        self._engine.set_amplitudes(amplitudes)
        self._virt = list(range(n))

    # -- transitions -----------------------------------------------------------------------------
    def measurement(self, qubit_index, psi=None):
        """Measure ``qubit_index`` of the device state: returns (outcome, p0).

        The reference returns (outcome, collapse operator) and lets the caller multiply
        (qvm_wavefunction.py:96-129); here the kernel collapses in place.  ``psi`` is accepted for
        signature compatibility: a host vector is uploaded first."""
        if psi is not None:
            self.wf = psi
        return self._engine.measure(self._virt[qubit_index], _engine.host_uniforms(self._engine))

    def _transition(self, instruction):
        if isinstance(instruction, Measurement):
            outcome, _ = self.measurement(value_get(instruction.qubit))
            self.classical_memory[value_get(instruction.classical_reg)] = outcome
            self.program_counter += 1
        elif isinstance(instruction, Gate):
            matrix, qubits = resolve_gate(self.gate_set, self.defgate_set, instruction)
            _engine.validate_qubits(qubits, self.num_qubits)
            if _is_swap(matrix, qubits):
                a, b = qubits                                   # a relabel, not a pass over the state
                self._virt[a], self._virt[b] = self._virt[b], self._virt[a]
            else:
                self._engine.queue_matrix(matrix, [self._virt[q] for q in qubits])
            self.program_counter += 1
        elif not self._classical_transition(instruction):
            raise self._unsupported(instruction)

    def kernel(self):
        """QAM.kernel (qam.py:96-100), except that a run of MEASURE instructions on distinct qubits is
        handed to the engine as one request: the same uniforms are drawn in the same order and the same
        sequential decisions are taken, but on the joint distribution of the measured bits -- one pass
        over the state per chunk of qubits instead of two per qubit.  ``transition`` stays single-step."""
        # (a subclass that overrides a hook or a transition sees every MEASURE one by one, as in the reference)
        cls = type(self)
        fused = hasattr(self._engine, "measure_many") and all(
            getattr(cls, name) is getattr(QVM_Wavefunction, name)
            for name in ("pre", "post", "transition", "_transition", "measurement"))
        while self.program_counter < len(self.program):
            instruction = self.current_instruction()
            if fused and isinstance(instruction, Measurement):
                run, seen = [], set()
                j = self.program_counter
                while j < len(self.program) and isinstance(self.program[j], Measurement):
                    q = value_get(self.program[j].qubit)
                    if q in seen:
                        break
                    seen.add(q)
                    run.append(self.program[j])
                    j += 1
                if len(run) > 1:
                    uniforms = [_engine.host_uniforms(self._engine) for _ in run]
                    outcomes = self._engine.measure_many([self._virt[value_get(m.qubit)] for m in run], uniforms)
                    for m, (outcome, _) in zip(run, outcomes):
                        self.classical_memory[value_get(m.classical_reg)] = outcome
                    self.program_counter = j
                    continue
            if self.transition(instruction):
                break

    def transition(self, instruction):
        """One full transition including the (empty) pre/post noise hooks; True when halted."""
        self.pre()
        self._transition(instruction)
        self.post()
        return self.program_counter == len(self.program)

    def pre(self):
        pass

    def post(self):
        pass

    # -- public API --------------------------------------------------------------------------------
    def _address_mask(self, classical_addresses):
        if classical_addresses is None:
            return None
        assert len(set(classical_addresses)) == len(classical_addresses)
        if np.min(classical_addresses) < 0 or np.max(classical_addresses) >= len(self.classical_memory):
            raise RuntimeError("Improper classical memory access outside allocated classical memory range.")
        return np.array(classical_addresses)

    def _simulate(self, pyquil_program, classical_addresses):
        self._start(pyquil_program)
        mask = self._address_mask(classical_addresses)
        self.kernel()
        return mask

    def _run_once(self, pyquil_program, classical_addresses):
        """One full execution; the classical bits only (the reference's run() calls wavefunction() and
        drops the amplitudes, :279-283 -- here they simply stay on the device)."""
        mask = self._simulate(pyquil_program, classical_addresses)
        self._engine.sync()
        memory = self.classical_memory if mask is None else self.classical_memory[mask]
        return [int(x) for x in memory]

    def wavefunction(self, pyquil_program, classical_addresses=None):
        """Simulate a program; returns (Wavefunction, [classical bits at ``classical_addresses``])."""
        mask = self._simulate(pyquil_program, classical_addresses)
        memory = self.classical_memory if mask is None else self.classical_memory[mask]
        results = [int(x) for x in memory]
        self._materialise_swaps()
        if self.num_qubits > LAZY_QUBITS:
            self._engine.flush()
            return Wavefunction(_engine.DeviceAmplitudes(self._engine)), results
        return Wavefunction(self._engine.amplitudes()), results

    def expectation(self, pyquil_program, operator_programs=None):
        """[<psi|O|psi> for O in operator_programs] with |psi> the state ``pyquil_program`` prepares.

        The pyquil QVM call the reference's QVMs declare but leave unimplemented (qvm_unitary.py:99-112,
        qvm_density.py:396-408).  Operators: ``PauliTerm`` / ``PauliSum`` (the operator ``tensor_up`` builds,
        unitary_generator.py:366-411) or gate Programs; default: the identity.  Pauli strings cost one reduction
        pass each and leave the state alone; other programs run on a copy (referenceqvm/_expectation.py)."""
        from referenceqvm import _expectation
        if operator_programs is None:
            operator_programs = [Program()]
        self._simulate(pyquil_program, None)
        eng, virt, n = self._engine, self._virt, self.num_qubits

        def pauli(ops):
            return eng.pauli_expectation({virt[q]: letter for q, letter in ops.items()})

        def general(program):
            if not isinstance(eng, _engine.Engine):
                raise NotImplementedError("operator programs that are not Pauli strings need a single-GPU state")
            scratch = eng.clone()
            try:
                defs = dict((dg.name, dg.matrix) for dg in program.defined_gates)
                for instruction in program:
                    if not isinstance(instruction, Gate):
                        raise TypeError("operator programs may contain gates only")
                    matrix, qubits = resolve_gate(self.gate_set, defs, instruction)
                    _engine.validate_qubits(qubits, n)
                    scratch.queue_matrix(matrix, [virt[q] for q in qubits])
                return eng.inner_product(scratch)
            finally:
                scratch.close()

        return [_expectation.evaluate(op, n, pauli, general) for op in operator_programs]

    def _terminal_measure_block(self, program):
        """Index where a purely terminal run of MEASUREs starts, or None when sampling the final
        state would not be equivalent to re-running the program per trial."""
        instrs = list(program)
        if any(isinstance(i, (Jump, JumpConditional, JumpTarget, Halt)) for i in instrs):
            return None
        first = next((k for k, i in enumerate(instrs) if isinstance(i, Measurement)), None)
        if first is None:
            return len(instrs)
        if not all(isinstance(i, Measurement) for i in instrs[first:]):
            return None
        return first

    def run(self, pyquil_program, classical_addresses=None, trials=1):
        """Run a program ``trials`` times; list of the classical bits at ``classical_addresses``."""
        if not isinstance(pyquil_program, Program):
            raise TypeError("I can only generate from pyQuil programs")
        split = self._terminal_measure_block(pyquil_program) if self.sampling_fast_path else None
        if split is not None and split == len(pyquil_program) and trials > 1:
            # no MEASURE at all: every trial is the same deterministic run
            classical_vals = self._run_once(pyquil_program, classical_addresses)
            return [list(classical_vals) for _ in range(trials)]
        if split is None and trials > 1 and self.sampling_fast_path and self._batchable(pyquil_program):
            if self._has_control_flow(pyquil_program):
                return self._run_batched_flow(pyquil_program, classical_addresses, trials)
            return self._run_batched(pyquil_program, classical_addresses, trials)
        if split is None or trials <= 1:
            return [self._run_once(pyquil_program, classical_addresses) for _ in range(trials)]

        # simulate the measurement-free prefix once, then sample the terminal MEASURE block
        self._start(pyquil_program)
        mask = self._address_mask(classical_addresses)
        while self.program_counter < split:
            self.transition(self.current_instruction())
        indices, _ = self._engine.sample(_engine.host_uniforms(self._engine, trials), physical=True)
        memory = np.repeat(self.classical_memory[None, :], trials, axis=0)
        bits = _index_bits(indices, self.num_qubits)
        for instr in list(pyquil_program)[split:]:
            q, c = value_get(instr.qubit), value_get(instr.classical_reg)
            memory[:, c] = bits[:, self._engine.position(self._virt[q])]
        self.program_counter = len(self.program)
        self.classical_memory = memory[-1].copy()
        picked = memory if mask is None else memory[:, mask]
        return _rows_to_lists(picked.astype(np.int64))

    # -- batched trials: MEASURE in mid-circuit, no control flow ---------------------------------------
    #: a batch holds at most 2^BATCH_BITS amplitudes (1 GiB); one CUDA block measures one trial, so
    #: trials of more than 2^BATCH_MAX_QUBITS amplitudes keep the per-trial loop
    BATCH_BITS = 26
    BATCH_MAX_QUBITS = 20
    BATCH_MAX_TRIAL_BITS = 20

    def _batchable(self, program):
        """run(trials) can execute all trials side by side when every trial walks the same instruction
        sequence: gates, MEASUREs and classical bit operations only (the reference re-simulates per
        trial, :279-283; with jumps the trials may diverge and the per-trial loop is kept)."""
        from referenceqvm import _sharded
        if _sharded.context() is not None and _sharded.context().world > 1:
            return False
        allowed = (Gate, Measurement, UnaryClassicalInstruction, BinaryClassicalInstruction, Jump, JumpConditional,
                   JumpTarget, Halt)
        instrs = list(program)
        if not instrs or not all(isinstance(i, allowed) for i in instrs):
            return False
        n = 1 + max([value_get(q) for i in instrs if isinstance(i, (Gate, Measurement)) for q in i.get_qubits()] or [0])
        return n <= self.BATCH_MAX_QUBITS

    @staticmethod
    def _has_control_flow(program):
        return any(isinstance(i, (Jump, JumpConditional, JumpTarget, Halt)) for i in program)

    def _run_batched(self, pyquil_program, classical_addresses, trials):
        """All trials as independent segments of one device buffer: every gate group is one pass over
        all of them, every MEASURE one kernel with one uniform per trial.  The uniforms are drawn up
        front in the order the per-trial loop would draw them (trial-major, one per MEASURE), so a
        seeded run returns the same bits as the reference's loop."""
        self.load_program(pyquil_program)
        mask = self._address_mask(classical_addresses)
        n = self.num_qubits
        instrs = list(pyquil_program)
        n_meas = sum(1 for i in instrs if isinstance(i, Measurement))
        uniforms = np.random.random_sample(trials * n_meas).reshape(trials, n_meas)
        # at most 2^20 trial segments per batch (one CUDA block each: qvm_measure_batched's limit); more
        # trials simply take more batches
        t_bits = max(0, min((trials - 1).bit_length(), self.BATCH_BITS - n, self.BATCH_MAX_TRIAL_BITS))
        cap = 1 << t_bits
        memory = np.repeat(self.classical_memory[None, :], trials, axis=0)
        eng = _engine.Engine(n + t_bits, device=self.device)
        eng.relabel = False                   # the trial index must stay on the top positions
        last = None
        try:
            for lo in range(0, trials, cap):
                cnt = min(cap, trials - lo)
                eng.pending = []
                eng.state.reset_batched(n)
                virt = list(range(n))
                mem = memory[lo:lo + cnt]
                m = 0
                for instruction in instrs:
                    if isinstance(instruction, Measurement):
                        eng.flush()
                        u = np.full(cap, 0.5)
                        u[:cnt] = uniforms[lo:lo + cnt, m]
                        outcomes, _ = eng.state.measure_batched(n, virt[value_get(instruction.qubit)], u)
                        mem[:, value_get(instruction.classical_reg)] = outcomes[:cnt] != 0
                        m += 1
                    elif isinstance(instruction, Gate):
                        matrix, qubits = resolve_gate(self.gate_set, self.defgate_set, instruction)
                        _engine.validate_qubits(qubits, n)
                        if _is_swap(matrix, qubits):
                            a, b = qubits
                            virt[a], virt[b] = virt[b], virt[a]
                        else:
                            eng.queue_matrix(matrix, [virt[q] for q in qubits])
                    else:
                        self._batched_classical(instruction, mem)
                eng.flush()
                if lo + cnt == trials:          # the reference leaves the last trial's state in self.wf
                    last = (eng.amplitudes((cnt - 1) << n, 1 << n), virt)
        finally:
            eng.close()
        self._fresh_engine()
        if last is not None:
            self._engine.set_amplitudes(last[0])
            self._virt = list(last[1])
        self.program_counter = len(self.program)
        self.classical_memory = memory[-1].copy()
        out = memory if mask is None else memory[:, mask]
        return _rows_to_lists(out.astype(np.int64))

    # -- batched trials under classical control: if_then / while_do / jumps ---------------------------------
    #: a group that runs this many instructions without finishing is taken for a non-terminating program
    BATCH_FLOW_MAX_STEPS = 1 << 22

    def _run_batched_flow(self, pyquil_program, classical_addresses, trials):
        """run() for programs WITH jumps (reference: one full re-simulation per trial, :279-283, control flow at
        :164-231).  Trials start as one batch (one n-qubit segment each, as in ``_run_batched``); a batch walks
        the program with ONE program counter.  At a conditional jump whose condition bit differs between its
        trials the batch splits in two -- the segments are regrouped on the device (``qvm_gather_segments``) --
        and both halves continue independently (loops keep peeling off the trials that leave).  Work is what
        the per-trial loop does, but every gate pass and every MEASURE serves a whole batch.

        Random numbers: trial t uses row t of a (trials x draws) matrix of uniforms, extended by blocks as some
        path needs more draws; the distribution of the results is the reference's, but a path-dependent number
        of draws per trial cannot be laid out like the reference's trial-after-trial stream, so seeded results
        differ from the per-trial loop (``sampling_fast_path = False`` keeps that loop)."""
        self.load_program(pyquil_program)
        mask = self._address_mask(classical_addresses)
        n = self.num_qubits
        instrs = list(pyquil_program)
        labels = {}
        for index, instruction in enumerate(instrs):
            if isinstance(instruction, JumpTarget) and instruction.label.name not in labels:
                labels[instruction.label.name] = index          # first match, as QAM.find_label

        def target(label):
            if label.name not in labels:
                raise RuntimeError("Improper program - Jump Target not found in the input program!")
            return labels[label.name]

        uniforms = [np.random.random_sample((trials, 4))]        # column blocks, drawn as paths ask for them

        def draws(ids, k):
            have = sum(b.shape[1] for b in uniforms)
            while k >= have:
                uniforms.append(np.random.random_sample((trials, 4)))
                have += 4
            for block in uniforms:
                if k < block.shape[1]:
                    return block[ids, k]
                k -= block.shape[1]

        t_bits_cap = max(0, min(self.BATCH_BITS - n, self.BATCH_MAX_TRIAL_BITS))
        memory = np.repeat(self.classical_memory[None, :], trials, axis=0)
        last_state = [None]

        def make_engine(count):
            eng = _engine.Engine(n + max(0, (count - 1).bit_length()), device=self.device)
            eng.relabel = False                   # the trial index must stay on the top positions
            return eng

        def run_group(eng, ids, pc, virt, mem, k, work):
            """Walk one batch until it ends or splits; ``k`` = uniforms its trials have used so far."""
            cap = 1 << (eng.n - n)
            steps = 0
            while pc < len(instrs):
                steps += 1
                if steps > self.BATCH_FLOW_MAX_STEPS:
                    raise RuntimeError("program does not terminate (batched run gave up after %d instructions)" % steps)
                instruction = instrs[pc]
                if isinstance(instruction, Gate):
                    matrix, qubits = resolve_gate(self.gate_set, self.defgate_set, instruction)
                    _engine.validate_qubits(qubits, n)
                    if _is_swap(matrix, qubits):
                        a, b = qubits
                        virt[a], virt[b] = virt[b], virt[a]
                    else:
                        eng.queue_matrix(matrix, [virt[q] for q in qubits])
                    pc += 1
                elif isinstance(instruction, Measurement):
                    eng.flush()
                    u = np.full(cap, 0.5)
                    u[:ids.size] = draws(ids, k)
                    k += 1
                    outcomes, _ = eng.state.measure_batched(n, virt[value_get(instruction.qubit)], u)
                    mem[:, value_get(instruction.classical_reg)] = outcomes[:ids.size] != 0
                    pc += 1
                elif isinstance(instruction, (UnaryClassicalInstruction, BinaryClassicalInstruction)):
                    self._batched_classical(instruction, mem)
                    pc += 1
                elif isinstance(instruction, JumpTarget):
                    pc += 1
                elif isinstance(instruction, Jump):
                    pc = target(instruction.target)
                elif isinstance(instruction, Halt):
                    pc = len(instrs)
                elif isinstance(instruction, JumpConditional):
                    cond = mem[:, value_get(instruction.condition)].astype(bool)
                    taken = cond if isinstance(instruction, JumpWhen) else ~cond
                    dest = target(instruction.target)
                    if taken.all():
                        pc = dest
                    elif not taken.any():
                        pc += 1
                    else:
                        # the batch splits: regroup the segments of each side into an engine of its own
                        eng.flush()
                        for side, next_pc in ((taken, dest), (~taken, pc + 1)):
                            where = np.nonzero(side)[0]
                            sub = make_engine(where.size)
                            sub.state.reset_batched(n)
                            sub.state.gather_segments(eng.state, n, where)
                            work.append((sub, ids[where], next_pc, list(virt), mem[where].copy(), k))
                        eng.close()
                        return
                else:
                    eng.close()
                    raise self._unsupported(instruction)
            eng.flush()
            memory[ids] = mem
            if ids[-1] == trials - 1:                # the reference leaves the last trial's state in self.wf
                last_state[0] = (eng.amplitudes((ids.size - 1) << n, 1 << n), list(virt))
            eng.close()

        chunk = 1 << t_bits_cap
        for lo in range(0, trials, chunk):
            ids = np.arange(lo, min(lo + chunk, trials))
            eng = make_engine(ids.size)
            eng.state.reset_batched(n)
            work = [(eng, ids, 0, list(range(n)), memory[ids].copy(), 0)]
            while work:
                item = work.pop()
                try:
                    run_group(*item, work)
                except Exception:
                    for other in work:
                        other[0].close()
                    raise
        self._fresh_engine()
        if last_state[0] is not None:
            self._engine.set_amplitudes(last_state[0][0])
            self._virt = list(last_state[0][1])
        self.program_counter = len(self.program)
        self.classical_memory = memory[-1].copy()
        out = memory if mask is None else memory[:, mask]
        return _rows_to_lists(out.astype(np.int64))

    @staticmethod
    def _batched_classical(instruction, mem):
        """qam.py's classical transitions on a (trials, addresses) bit matrix."""
        if isinstance(instruction, UnaryClassicalInstruction):
            t = value_get(instruction.target)
            if isinstance(instruction, ClassicalTrue):
                mem[:, t] = True
            elif isinstance(instruction, ClassicalFalse):
                mem[:, t] = False
            elif isinstance(instruction, ClassicalNot):
                mem[:, t] = ~mem[:, t]
            else:
                raise TypeError("Invalid UnaryClassicalInstruction")
        else:
            left, right = value_get(instruction.left), value_get(instruction.right)
            lv, rv = mem[:, left].copy(), mem[:, right].copy()
            if isinstance(instruction, ClassicalAnd):
                mem[:, right] = lv & rv
            elif isinstance(instruction, ClassicalOr):
                mem[:, right] = lv | rv
            elif isinstance(instruction, ClassicalMove):
                mem[:, right] = lv
            elif isinstance(instruction, ClassicalExchange):
                mem[:, left], mem[:, right] = rv, lv
            else:
                raise TypeError("Invalid BinaryClassicalInstruction")

    def run_and_measure(self, pyquil_program, qubits=None, trials=1):
        """Simulate, then measure ``qubits`` (in the given order) ``trials`` times.

        Qubits beyond the program's register read 0 (reference :316-317)."""
        if qubits is None:
            qubits = []
        elif type(qubits) is not list:
            raise TypeError("Must pass in qubit indices as list")
        if not isinstance(pyquil_program, Program):
            raise TypeError("I can only generate from pyQuil programs")

        deterministic = not any(isinstance(i, Measurement) for i in pyquil_program)
        if self.sampling_fast_path and deterministic and trials > 1 and len(qubits) > 0:
            t0 = time.perf_counter()
            self._simulate(pyquil_program, None)
            t1 = time.perf_counter()
            # the launches are queued: the host draws its uniforms while the device works (a deterministic program
            # draws nothing before this point, so the stream of np.random is where the reference's would be)
            uniforms = _engine.host_uniforms(self._engine, trials)
            t1b = time.perf_counter()
            self._engine.sync()
            t2 = time.perf_counter()
            indices, _ = self._engine.sample(uniforms, physical=True)
            t3 = time.perf_counter()
            # physical bit position each requested qubit reads; -1 (always 0) for qubits beyond the register
            sel = [self._engine.position(self._virt[q]) if 0 <= q < self.num_qubits else -1 for q in qubits]
            out = _rows_to_lists(indices=indices, sel=sel)
            #: where the last call's wall time went (bench.py reports it as e2e.breakdown)
            self.last_timing = {"program_walk_s": t1 - t0, "plan_launch_sync_s": t2 - t1b, "host_uniforms_s": t1b - t1,
                                "cdf_and_draws_s": t3 - t2,
                                "bits_to_lists_s": time.perf_counter() - t3}
            return out

        results = []
        for _ in range(trials):
            self._simulate(pyquil_program, None)
            inside = [q for q in qubits if q < self.num_qubits]
            if hasattr(self._engine, "measure_many") and len(set(inside)) == len(inside) > 1:
                # one uniform per measured qubit, in the order the loop below would draw them
                uniforms = [_engine.host_uniforms(self._engine) for _ in inside]
                outcomes = iter(self._engine.measure_many([self._virt[q] for q in inside], uniforms))
                trial_results = [next(outcomes)[0] if q < self.num_qubits else 0 for q in qubits]
                results.append(trial_results)
                continue
            trial_results = []
            for qubit_index in qubits:
                if qubit_index < self.num_qubits:
                    measured_val, _ = self._engine.measure(self._virt[qubit_index], _engine.host_uniforms(self._engine))
                else:
                    measured_val = 0
                trial_results.append(measured_val)
            results.append(trial_results)
        return results
