"""Density-matrix QVM: rho is a 2n-qubit vector in HBM driven by the same kernels as psi.

Public surface follows reference ``referenceqvm/qvm_density.py``: ``NoiseModel`` (:57-86),
``QVM_Density.measurement`` (:120-143), ``transition`` with ``_pre`` / ``_post`` noise hooks
(:242-348, :363-371), ``apply_kraus`` (:350-361), ``density`` (:373-394), ``run`` (:410-432),
``run_and_measure`` (:434-452).

Layout: rho is flattened row-major, vector index (i << n) | j  <->  rho[i, j].
``rho <- U rho U^dagger`` is U on vector-qubits (a+n) followed by conj(U) on (a); a channel
``rho <- sum_k K_k rho K_k^dagger`` on qubit q is ONE 4x4 non-unitary matrix
``S = sum_k K_k (x) conj(K_k)`` on vector-qubits (q+n, q).  All of these are queued through the
fusion planner, so a gate plus the noise on every qubit after it is typically one or two passes
over HBM instead of the reference's 2 n #Kraus sparse products.
"""
import os

import numpy as np

from pyquil.quil import Program
from pyquil.quilbase import Gate, Measurement

from referenceqvm import _compiled, _engine, _native
from referenceqvm.gates import noise_gates, utility_gates  # noqa: F401
from referenceqvm.qam import QAM
from referenceqvm.unitary_generator import lifted_gate, resolve_gate, tensor_gates, value_get  # noqa: F401

INFINITY = float("inf")
"Used for infinite coherence times."


def sparse_trace(sparse_matrix):
    """Trace of a matrix-like (reference qvm_density.py:51-54)."""
    if isinstance(sparse_matrix, DeviceDensity):
        return np.sum(sparse_matrix.diagonal())
    return np.sum(sparse_matrix.diagonal())


class NoiseModel(object):
    """T1 / T2 / readout / Pauli-channel parameters (reference qvm_density.py:57-86)."""

    def __init__(self, T1=30e-6, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-09, ro_fidelity=0.95,
                 depolarizing=None, bitflip=None, phaseflip=None, bitphaseflip=None):
        self.T1 = T1
        self.T2 = T2
        self.gate_time_1q = gate_time_1q
        self.gate_time_2q = gate_time_2q
        self.gate_time_3q = float(200E-9)
        self.gate_time_4q = float(200E-9)
        self.gate_time_5q = float(200E-9)
        self.ro_fidelity = ro_fidelity
        self.depolarizing = depolarizing
        self.bitflip = bitflip
        self.phaseflip = phaseflip
        self.bitphaseflip = bitphaseflip


_channel_pieces = {}       # (n, Kraus operator bytes) -> expanded ops of the channel on (n, 0)


def kraus_superoperator(operators):
    """S = sum_k K_k (x) conj(K_k): the channel as one matrix on vector-qubits (q+n.., q..)."""
    ops = [np.asarray(k, dtype=np.complex128) for k in operators]
    return sum(np.kron(k, np.conj(k)) for k in ops)


class DeviceDensity(object):
    """rho resident on the device(s), quacking enough like the reference's ``csc_matrix`` result
    (``.todense()``, ``.toarray()``, ``.diagonal()``, ``[i, j]``, ``.shape``).

    Two flavours.  ``density()`` and ``apply_kraus()`` hand out SNAPSHOTS that own an engine of their own (the
    reference returns an independent matrix from every call, qvm_density.py:391-394): later calls on the QVM
    do not change them.  The ``_density`` attribute is a live VIEW of the QVM's engine (the reference's
    attribute is the QVM's own matrix too): it flushes queued gates before every read, follows the QVM as it
    runs, and raises once the QVM has replaced that engine's state."""

    def __init__(self, engine, num_qubits, live=False):
        self._engine = engine             # Engine / ShardedEngine holding the 2n-qubit vector
        self._live_state = engine.state if live else None
        self.num_qubits = num_qubits
        dim = 1 << num_qubits
        self.shape = (dim, dim)
        self.dtype = np.dtype(np.complex128)

    @property
    def _state(self):                     # (single-GPU engines: the raw DeviceState; kept for `_density = rho`)
        return self._engine.state

    def close(self):
        """Release a snapshot's device memory now (otherwise when the object is collected)."""
        if self._live_state is None and self._engine is not None and not getattr(self, "_adopted", False):
            try:
                self._engine.close()
            except Exception:             # interpreter shutdown
                pass
        self._engine = None

    def __del__(self):
        self.close()

    def _check(self):
        if self._live_state is not None and self._engine.state is not self._live_state:
            raise RuntimeError("this view of the QVM's density matrix is stale: the QVM has since replaced its "
                               "state (density() returns independent snapshots)")

    def toarray(self):
        self._check()
        return self._engine.amplitudes().reshape(self.shape)

    def todense(self):
        return np.asmatrix(self.toarray())

    def tocsc(self):
        import scipy.sparse as sps
        return sps.csc_matrix(self.toarray())

    def diagonal(self):
        self._check()
        dim = self.shape[0]
        eng = self._engine
        if isinstance(eng, _engine.Engine) and eng.layout_is_identity():
            return eng.strided(0, dim + 1, dim)
        return eng.density_diag(self.num_qubits).astype(np.complex128)

    def trace(self):
        return np.sum(self.diagonal())

    def __getitem__(self, key):
        i, j = key
        dim = self.shape[0]
        i, j = int(i) % dim, int(j) % dim
        self._check()
        return self._engine.probe([i * dim + j])[0]

    def __array__(self, dtype=None, copy=None):
        a = self.toarray()
        return a if dtype is None else a.astype(dtype)


class QVM_Density(QAM):
    """Supports density(), run(), run_and_measure(), apply_kraus()."""

    def __init__(self, num_qubits=None, program=None, program_counter=None, classical_memory=None,
                 gate_set=None, defgate_set=None, density=None, noise_model=None):
        super(QVM_Density, self).__init__(num_qubits=num_qubits, program=program,
                                          program_counter=program_counter,
                                          classical_memory=classical_memory,
                                          gate_set=gate_set, defgate_set=defgate_set)
        self.noise_model = noise_model
        self.all_inst = True
        self._engine = None
        self.device = int(os.environ.get("QVM_B200_DEVICE", "0"))
        #: compiled programs (referenceqvm/_compiled.py): gate prefix + its noise channels planned once per Program
        self.plan_cache = _compiled.ENABLED
        if density is not None:
            self._density = density

    # -- device state ------------------------------------------------------------------------------
    def _engine_for(self, n):
        if self._engine is not None and self._engine.n != 2 * n:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            # one GPU, or -- when sharding is enabled in this process -- the 2n-qubit vector split by its top
            # positions, i.e. by rows of rho (reference: one csc_matrix, qvm_density.py:390-391)
            self._engine = _engine.make_engine(2 * n, device=self.device)
            if isinstance(self._engine, _engine.Engine):
                # single GPU: 2n <= 28 positions in practice; relabelling saves no pass there (profiles/README.md)
                self._engine.relabel = False
        return self._engine

    @property
    def _density(self):
        if self._engine is None:
            return None
        return DeviceDensity(self._engine, self._engine.n // 2, live=True)

    @_density.setter
    def _density(self, value):
        if value is None:
            if self._engine is not None:
                self._engine.close()
            self._engine = None
            return
        if isinstance(value, DeviceDensity):
            eng = self._engine_for(value.num_qubits)
            if value._engine is eng:
                return
            if isinstance(eng, _engine.Engine) and isinstance(value._engine, _engine.Engine):
                value._engine.restore_layout()
                eng.adopt(value._engine.state)          # takes the buffer over (no copy)
                value._adopted = True
            else:
                eng.set_amplitudes(value.toarray().reshape(-1))
            return
        dense = value.toarray() if hasattr(value, "toarray") else np.asarray(value)
        if dense.ndim != 2 or dense.shape[0] != dense.shape[1] or dense.shape[0] & (dense.shape[0] - 1):
            raise ValueError("density must be a square matrix of power-of-two dimension")
        n = dense.shape[0].bit_length() - 1
        self._engine_for(n).set_amplitudes(np.ascontiguousarray(dense, dtype=np.complex128).reshape(-1))

    # -- building blocks -----------------------------------------------------------------------------
    def _queue_conjugation(self, matrix, qubits):
        """rho <- M rho M^dagger : M on the row qubits, conj(M) on the column qubits."""
        n = self.num_qubits
        matrix = np.asarray(matrix, dtype=np.complex128)
        _engine.validate_qubits(qubits, n)
        self._engine.queue_matrix(matrix, [q + n for q in qubits])
        self._engine.queue_matrix(np.conj(matrix), list(qubits))

    def _queue_channel(self, operators, index, superoperator=None):
        n = self.num_qubits
        if not (0 <= index < n):
            raise ValueError("Gate index out of range!")
        if superoperator is None:
            superoperator = kraus_superoperator(operators)
        self._engine.queue_matrix(superoperator, [index + n, index])

    def _queue_channel_everywhere(self, operators):
        """The same channel on every qubit (the noise model after an instruction): one superoperator,
        classified and expanded once on (n, 0) -- qubit i is the same op with every position + i --
        and remembered per (Kraus operators, n): a noise model applies the same few channels after
        every instruction."""
        n = self.num_qubits
        if hasattr(self._engine, "queue_shifted"):
            key = (n,) + tuple(np.asarray(k, dtype=np.complex128).tobytes() for k in operators)
            pieces = _channel_pieces.get(key)
            if pieces is None:
                op = _engine.compile_gate(kraus_superoperator(operators), [n, 0])
                pieces = () if op.kind == _engine.nat.OP_NOP else _engine.expand_op(op)
                if len(_channel_pieces) > 4096:
                    _channel_pieces.clear()
                _channel_pieces[key] = pieces
            for i in range(n):
                self._engine.queue_shifted(pieces, i)
            return
        sup = kraus_superoperator(operators)
        for i in range(n):
            self._queue_channel(operators, i, sup)

    def measurement(self, qubit_index):
        """MEASURE on rho: p0 = tr(P0 rho); one uniform draw; rho <- P_b rho P_b / p_b.

        Returns (outcome, p0).  (Reference :120-143 returns the collapse operator instead.)"""
        n = self.num_qubits
        p0 = self._engine.density_prob_zero(n, qubit_index)
        if _engine.shared_uniform(self._engine, _engine.host_uniforms(self._engine)) < p0:
            outcome, scale = 0, 1.0 / np.sqrt(p0)
        else:
            outcome, scale = 1, 1.0 / np.sqrt(1 - p0)
        proj = np.zeros((2, 2), dtype=np.complex128)
        proj[outcome, outcome] = scale
        self._queue_conjugation(proj, [qubit_index])
        return outcome, p0

    def _transition(self, instruction):
        if isinstance(instruction, Measurement):
            outcome, _ = self.measurement(value_get(instruction.qubit))
            self.classical_memory[value_get(instruction.classical_reg)] = outcome
            self.program_counter += 1
        elif isinstance(instruction, Gate):
            matrix, qubits = resolve_gate(self.gate_set, self.defgate_set, instruction)
            self._queue_conjugation(matrix, qubits)
            self.program_counter += 1
        elif not self._classical_transition(instruction):
            raise self._unsupported(instruction)

    def _pre(self, instruction):
        """Readout-noise hook before MEASURE.

        Reproduces the reference arithmetic (:254-265) as written: for every qubit and every Kraus
        operator K of bitphase_flip(1 - ro_fidelity) it does ``rho += K rho`` -- i.e. (I + K) on
        the row index.  That is not a physical channel ("reference defects"); it is
        kept so results stay identical to the reference on the same inputs."""
        if self.noise_model is None or not isinstance(instruction, Measurement):
            return
        n = self.num_qubits
        ro_prob = 1 - self.noise_model.ro_fidelity
        for iqubit in range(n):
            for kop in noise_gates['bitphase_flip'](ro_prob):
                self._engine.queue_matrix(np.eye(2) + np.asarray(kop, dtype=np.complex128), [iqubit + n])

    def _post(self, instruction):
        """Noise after every instruction, on every qubit (reference :267-348)."""
        nm = self.noise_model
        if nm is None:
            return
        n = self.num_qubits

        def decay(t_gate_1q, t_gate_2q, t_coherence):
            span = len(instruction.get_qubits())
            t_gate = t_gate_1q if span == 1 else t_gate_2q
            return 1.0 - np.exp(-t_gate / t_coherence)

        if nm.T1 != INFINITY and isinstance(instruction, Gate):
            p = decay(nm.gate_time_1q, nm.gate_time_2q, nm.T1)
            if np.isclose(p, 1.0):
                # the reference leaves `relaxation_ops` unbound here (:292-296) and fails
                raise UnboundLocalError("relaxation probability is 1: the reference cannot apply this channel")
            ops = noise_gates['relaxation'](p)
            self._queue_channel_everywhere(ops)

        if nm.T2 != INFINITY:
            p = decay(nm.gate_time_1q, nm.gate_time_2q, nm.T2)
            if not np.isclose(p, 1.0):
                ops = noise_gates['dephasing'](p)
                self._queue_channel_everywhere(ops)

        for attr, key, label in (("depolarizing", "depolarizing", "depolarizing"),
                                 ("bitflip", "bit_flip", "bitflip"),
                                 ("phaseflip", "phase_flip", "phaseflip")):
            value = getattr(nm, attr)
            if value is not None:
                if not isinstance(value, float):
                    raise TypeError("%s noise value should be None or a float" % label)
                ops = noise_gates[key](value)
                self._queue_channel_everywhere(ops)

        if nm.bitphaseflip is not None:
            # the reference type-checks `phaseflip` here (:341), kept for identical behaviour
            if not isinstance(nm.phaseflip, float):
                raise TypeError("bitphaseflip noise value should be None or a float")
            ops = noise_gates['bitphase_flip'](nm.bitphaseflip)
            self._queue_channel_everywhere(ops)

    def apply_kraus(self, operators, index):
        """sum_k L_k rho L_k^dagger with L_k = K_k lifted to ``index``; returns the new density and
        leaves ``self._density`` untouched, like the reference (:350-361)."""
        n = self.num_qubits
        if self._engine is None or self._engine.n != 2 * n:
            raise ValueError("no density of %r qubits is loaded" % (n,))
        if not (0 <= index < n):
            raise ValueError("Gate index out of range!")
        scratch = self._engine.clone()
        scratch.queue_matrix(kraus_superoperator(operators), [index + n, index])
        scratch.flush()
        return DeviceDensity(scratch, n)

    def transition(self, instruction):
        self._pre(instruction)
        self._transition(instruction)
        self._post(instruction)

    # -- public API ------------------------------------------------------------------------------------
    def load_program(self, pyquil_program):
        """As QAM.load_program; additionally keeps a user-assigned ``_density`` when its size matches
        the program (the reference's tests assign ``_density``, then load_program + kernel)."""
        super(QVM_Density, self).load_program(pyquil_program)
        if self._engine is not None and self._engine.n != 2 * self.num_qubits:
            self._engine.close()
            self._engine = None

    def density(self, pyquil_program):
        """Density matrix of a program (starts from |0..0><0..0|)."""
        self._simulate_density(pyquil_program)
        return DeviceDensity(self._engine.clone(), self.num_qubits)             # an independent snapshot

    _HOOKS = ("_pre", "_post", "transition", "_transition", "measurement", "kernel", "load_program", "apply_kraus")

    def _plan_key(self):
        nm = self.noise_model
        noise = None if nm is None else (nm.T1, nm.T2, nm.gate_time_1q, nm.gate_time_2q, nm.ro_fidelity, nm.depolarizing,
                                         nm.bitflip, nm.phaseflip, nm.bitphaseflip)
        return ("density", id(self.gate_set), self.device, noise)

    def _simulate_density(self, pyquil_program):
        """load_program, rho = |0..0><0..0|, run.  The program's leading run of Gate instructions -- with the
        noise channels that follow each of them -- is recorded the first time a Program object comes by and
        replayed afterwards (referenceqvm/_compiled.py); MEASURE and classical instructions always step."""
        if not isinstance(pyquil_program, Program):
            raise TypeError("I can only generate from pyQuil programs")
        cache = (self.plan_cache and _compiled.uses_base_hooks(self, QVM_Density, self._HOOKS)
                 and (self._engine is None or isinstance(self._engine, _engine.Engine)))
        entry = _compiled.lookup(pyquil_program, self._plan_key()) if cache else None
        if entry is not None and not isinstance(self._engine_for(entry.num_qubits), _engine.Engine):
            entry = None
        if entry is not None:
            self.defgate_set = entry.defgate_set
            self.program = pyquil_program
            self.num_qubits = entry.num_qubits
            self.classical_memory = np.zeros(entry.n_cbits, dtype=bool)
            self._engine_for(self.num_qubits).replay(entry.compiled, reset_first=True)
            self.program_counter = entry.k
            self.kernel()
            return
        self.load_program(pyquil_program)
        eng = self._engine_for(self.num_qubits)
        eng.reset()
        k = _compiled.gate_prefix_length(pyquil_program) if cache and isinstance(eng, _engine.Engine) else 0
        if k:
            eng.begin_recording()
            try:
                while self.program_counter < k:
                    self.transition(self.current_instruction())
            except Exception:
                eng.end_recording(keep=False)
                raise
            compiled = eng.end_recording()
            _compiled.store(pyquil_program, self._plan_key(),
                            _compiled.Entry(_compiled.token_of(pyquil_program), k, compiled, None, self.num_qubits,
                                            len(self.classical_memory), self.defgate_set, True))
        self.kernel()

    def expectation(self, pyquil_program, operator_programs=None):
        """[tr(rho O) for O in operator_programs] with rho the (noisy) state ``pyquil_program`` prepares.

        The reference declares this method and raises NotImplementedError (qvm_density.py:396-408).  Operators:
        ``PauliTerm`` / ``PauliSum`` (``tensor_up``'s operator, unitary_generator.py:366-411) or Programs of
        Pauli gates; each Pauli string is a gather of 2^n entries of rho (referenceqvm/_expectation.py)."""
        from referenceqvm import _expectation
        if operator_programs is None:
            operator_programs = [Program()]
        self._simulate_density(pyquil_program)
        n, eng = self.num_qubits, self._engine

        def pauli(ops):
            return eng.density_pauli_expectation(n, ops)

        def general(program):
            # tr(rho O) = tr(O rho): O on the row qubits of a copy, then the trace of the copy
            scratch = eng.clone()
            try:
                defs = dict((dg.name, dg.matrix) for dg in program.defined_gates)
                for instruction in program:
                    if not isinstance(instruction, Gate):
                        raise TypeError("operator programs may contain gates only")
                    matrix, qubits = resolve_gate(self.gate_set, defs, instruction)
                    _engine.validate_qubits(qubits, n)
                    scratch.queue_matrix(matrix, [q + n for q in qubits])
                return complex(np.sum(scratch.density_pauli_expectation(n, {})))
            finally:
                scratch.close()

        return [_expectation.evaluate(op, n, pauli, general) for op in operator_programs]

    def run(self, pyquil_program, classical_addresses=None, trials=1):
        """Re-simulates per trial and returns the classical memory slices (numpy bool arrays)."""
        results = []
        for _ in range(trials):
            self._simulate_density(pyquil_program)
            results.append(self.classical_memory[classical_addresses])
        return list(results)

    def run_and_measure(self, pyquil_program, trials=1):
        """Simulate once, then draw ``trials`` basis states from diag(rho) by inverse-CDF sampling.

        Uses the uniforms numpy's ``np.random.choice(p=)`` would consume (reference :445-447), so a
        seeded call returns the reference's draws; bitstrings are big-endian (:449-451)."""
        self._simulate_density(pyquil_program)
        n = self.num_qubits
        uniforms = _engine.host_uniforms(self._engine, trials)
        indices, total = self._engine.sample(uniforms, mode=1, n_rho=n)
        if abs(total - 1.0) > 1e-8:
            raise ValueError("probabilities do not sum to 1")
        shifts = np.arange(n - 1, -1, -1, dtype=np.uint64)
        bits = (indices[:, None] >> shifts[None, :]) & np.uint64(1)
        from referenceqvm.qvm_wavefunction import _rows_to_lists
        return _rows_to_lists(bits.astype(np.int64))
