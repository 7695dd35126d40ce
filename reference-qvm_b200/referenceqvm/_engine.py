"""Execution engine shared by the wavefunction / density / unitary QVMs.

Owns the HBM-resident state (a ``DeviceState``), a queue of classified-but-not-yet-applied gates and
the logical->physical qubit layout.  Gates are queued as the state machine walks the program and
flushed -- through the fusion planner, one tile-kernel launch per group -- whenever something needs
the amplitudes (MEASURE, readback, sampling, a classical branch is never such a reason).
"""
import os

import numpy as np

from referenceqvm import _native as nat
from referenceqvm import _planner

_TILE_BITS = int(os.environ.get("QVM_B200_TILE_BITS", "11"))     # 11 / 4: best gate-apps/s in the round-1 A/B (profiles/)
_LOW_BITS = int(os.environ.get("QVM_B200_LOW_BITS", "4"))
#: in-tile qubit relabelling (``_planner.run_relabelled``): passes hand the qubits the next group needs
#: over to the low positions, so the layout (logical qubit -> physical position) changes as gates run
_RELABEL = os.environ.get("QVM_B200_RELABEL", "1") != "0"
_SWAP_MATRIX = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)

_classify_cache = {}
_walked_ahead = set()       # signatures of gate streams whose kernels were requested ahead of their run (Engine._compile_ahead)


def compile_gate(matrix, qubits):
    """(matrix, qubits) -> ClassifiedOp, memoised on the matrix bytes and the qubit tuple."""
    m = np.ascontiguousarray(matrix, dtype=np.complex128)
    key = (m.shape[0], m.tobytes(), tuple(qubits))
    op = _classify_cache.get(key)
    if op is None:
        op = nat.classify(m, qubits)
        if len(_classify_cache) > 65536:
            _classify_cache.clear()
        _classify_cache[key] = op
    return op


def expand_op(op):
    """Cheaper equivalents of a classified op, in execution order.

    A two-target diagonal d = (d00, d01, d10, d11) without zero entries factorises exactly into
    diag(d00, d01) on the low target, diag(1, d10/d00) on the high target and the doubly controlled
    scalar d11 d00 / (d01 d10).  Those are the kernel's cheapest op shapes (a per-thread scalar, or
    8 multiplies on one register slot), whereas a general two-target diagonal may need the generic
    table-driven handler.  Dephasing superoperators diag(1, f, f, 1) on (q+n, q) -- n of them after
    every instruction of a noisy density run (reference qvm_density.py:298-313) -- are the main user."""
    if op.kind != nat.OP_DIAG or len(op.targets) != 2:
        return (op,)
    d00, d01, d10, d11 = op.mat
    if d00 == 0 or d01 == 0 or d10 == 0 or d11 == 0:
        return (op,)
    hi, lo = op.targets
    out = [nat.ClassifiedOp(nat.OP_DIAG, (lo,), op.ctrl_mask, np.array([d00, d01], dtype=np.complex128))]
    r = d10 / d00
    if r != 1:
        out.append(nat.ClassifiedOp(nat.OP_DIAG, (hi,), op.ctrl_mask, np.array([1.0, r], dtype=np.complex128)))
    c = d11 * d00 / (d01 * d10)
    if c != 1:
        out.append(nat.ClassifiedOp(nat.OP_DIAG, (), op.ctrl_mask | (1 << hi) | (1 << lo),
                                    np.array([c], dtype=np.complex128)))
    return tuple(out)


class DiagonalMerger(object):
    """Peephole over a pending op list: an uncontrolled one-qubit diagonal (or a controlled scalar)
    is multiplied into the previous one on the same qubit (same control set) when no dense gate on
    those qubits was queued in between -- everything in between then commutes with it.  Noisy density
    runs queue the same dephasing factors on every qubit after every instruction; this keeps one op
    per qubit per stretch instead of one per instruction.  Factors are multiplied on the host (one
    rounding per merge, ~1e-16)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._d1 = {}        # target position -> index in pending of the last mergeable diag(d0, d1)
        self._cc = {}        # control mask -> index in pending of the last mergeable controlled scalar

    def absorb(self, pending, op):
        """True when ``op`` was merged into ``pending`` (nothing to append)."""
        if op.kind == nat.OP_DENSE:
            for t in op.targets:
                self._d1.pop(t, None)
                for mask in [m for m in self._cc if (m >> t) & 1]:
                    del self._cc[mask]
            return False
        if op.kind != nat.OP_DIAG:
            return False
        if len(op.targets) == 1 and not op.ctrl_mask:
            table, key = self._d1, op.targets[0]
        elif not op.targets and op.ctrl_mask:
            table, key = self._cc, op.ctrl_mask
        else:
            return False
        i = table.get(key)
        if i is None:
            table[key] = len(pending)
            return False
        prev = pending[i]
        pending[i] = nat.ClassifiedOp(prev.kind, prev.targets, prev.ctrl_mask, prev.mat * op.mat)
        return True


def validate_qubits(qubits, n_positions):
    """Index checks of the reference's permutation_arbitrary (unitary_generator.py:200-203);
    duplicates are rejected too (the reference's duplicate check :205-208 is dead code and a
    duplicate there ends in an unrelated exception)."""
    seen = set()
    for q in qubits:
        if not (0 <= q < n_positions):
            raise ValueError("Permutation SWAP index not valid")
        if q in seen:
            raise ValueError("Duplicate qubit %d found" % q)
        seen.add(q)


def host_uniforms(engine, count=None):
    """The state machine's draws from numpy's global RNG (one per MEASURE, ``trials`` per sampling call --
    reference qvm_wavefunction.py:120, qvm_density.py:131, :445).  On a sharded engine only rank 0 draws: the
    engine broadcasts rank 0's numbers anyway, and the other ranks' streams stay untouched."""
    ctx = getattr(engine, "ctx", None)
    if ctx is not None and ctx.rank != 0:
        return 0.0 if count is None else np.zeros(count)
    return np.random.random() if count is None else np.random.random_sample(count)


def shared_uniform(engine, r):
    """Rank 0's value of a host-drawn uniform, on every rank (identity on a single-GPU engine)."""
    if hasattr(engine, "_broadcast_floats"):
        return float(engine._broadcast_floats([r])[0])
    return r


def _signed_sum(values, zmask):
    """sum_i (-1)^popcount(i & zmask) values[i]."""
    i = np.arange(values.size, dtype=np.uint64) & np.uint64(zmask)
    parity = np.zeros(values.size, dtype=np.uint64)
    while np.any(i):
        parity ^= i & np.uint64(1)
        i >>= np.uint64(1)
    return complex(np.sum(np.where(parity == 1, -values, values)))


def sample_distribution(probabilities, uniforms, device=0):
    """Inverse-CDF draws (numpy's ``searchsorted(cumsum, u * total, 'right')``) from a plain array of
    2^k probabilities with the device sampler (``qvm_probs_prepare`` mode 2): returns (indices, total)."""
    p = np.ascontiguousarray(probabilities, dtype=np.float64).reshape(-1)
    k = int(p.size).bit_length() - 1
    if p.size < 2 or (1 << k) != p.size:
        raise ValueError("need a power-of-two number (>= 2) of probabilities")
    scratch = nat.DeviceState(k - 1, device)
    try:
        scratch.set_state(p.view(np.complex128))
        total = scratch.probs_prepare(2, 0)
        return scratch.sample(uniforms), total
    finally:
        scratch.close()


class CompiledFlush(object):
    """What ``Engine.end_recording`` returns: the launches of the recorded flushes (a library plan) and the
    qubit layout they leave behind."""
    __slots__ = ("plan", "pos_of", "n", "gates", "groups")

    def __init__(self, plan, pos_of, n, gates, groups):
        self.plan, self.pos_of, self.n, self.gates, self.groups = plan, pos_of, n, gates, groups


class Engine(object):
    """Single-GPU engine: every position is local."""

    def __init__(self, n_positions, device=0, tile_bits=None, low_bits=None, _state=None):
        self.n = int(n_positions)
        self.device = device
        self.tile_bits = _TILE_BITS if tile_bits is None else tile_bits
        self.low_bits = _LOW_BITS if low_bits is None else low_bits
        self.state = nat.DeviceState(self.n, device) if _state is None else _state
        self.pending = []
        self._merger = DiagonalMerger()
        self.gates_queued = 0
        self.groups_launched = 0
        self.relabel = _RELABEL
        #: logical qubit -> physical position (bit of the amplitude index in HBM)
        self.pos_of = list(range(self.n))

    @classmethod
    def from_state(cls, device_state, tile_bits=None, low_bits=None):
        """Engine around an existing DeviceState (takes ownership)."""
        return cls(device_state.n_local, device_state.device, tile_bits, low_bits, _state=device_state)

    # -- lifetime ----------------------------------------------------------------------------
    def close(self):
        self.pending = []
        self._merger.reset()
        if self.state is not None:
            self.state.close()
            self.state = None

    def adopt(self, device_state):
        """Replace the state buffer (same shape) -- used by the density QVM's ``_density`` setter."""
        assert device_state.n_local == self.n
        self.pending = []
        self._merger.reset()
        self.pos_of = list(range(self.n))
        self.state.close()
        self.state = device_state

    # -- state -------------------------------------------------------------------------------
    def reset(self):
        self.pending = []
        self._merger.reset()
        self.pos_of = list(range(self.n))
        self.state.reset_zero(True)

    def set_amplitudes(self, amplitudes):
        self.pending = []
        self._merger.reset()
        a = np.asarray(amplitudes).reshape(-1)
        if a.size != (1 << self.n):
            raise ValueError("state of %d amplitudes does not match %d positions" % (a.size, self.n))
        self.pos_of = list(range(self.n))
        self.state.set_state(a)

    def layout_is_identity(self):
        return all(p == q for q, p in enumerate(self.pos_of))

    def position(self, qubit):
        """Physical position (index bit) of a logical qubit: for bit look-ups in ``sample(physical=True)``."""
        return self.pos_of[qubit]

    def amplitudes(self, offset=0, count=None):
        """Amplitudes in LOGICAL index order (read through the layout; HBM is not rearranged)."""
        self.flush()
        if self.layout_is_identity():
            return self.state.get_state(offset, count)
        if count is None:
            count = (1 << self.n) - offset
        return self.state.get_layout(offset, 1, count, self.pos_of)

    def strided(self, offset, stride, count):
        self.flush()
        if self.layout_is_identity():
            return self.state.get_strided(offset, stride, count)
        return self.state.get_layout(offset, stride, count, self.pos_of)

    def probe(self, logical_indices):
        """Amplitudes at a handful of LOGICAL basis-state indices, read through the layout (no amplitude
        moves): same meaning as ``ShardedEngine.probe``."""
        self.flush()
        idx = np.asarray(logical_indices, dtype=np.int64).reshape(-1)
        out = np.zeros(idx.size, dtype=np.complex128)
        for j, logical in enumerate(idx):
            phys = 0
            for q in range(self.n):
                phys |= ((int(logical) >> q) & 1) << self.pos_of[q]
            out[j] = self.state.get_state(phys, 1)[0]
        return out

    def restore_layout(self):
        """Put every qubit back on its own position (SWAP passes) -- for the few consumers that address
        the raw buffer: clones, the density-matrix trace / diagonal kernels."""
        self.flush()
        if self.layout_is_identity():
            return
        pos = list(self.pos_of)
        at = {p: q for q, p in enumerate(pos)}
        swaps = []
        for q in range(self.n):
            p = pos[q]
            if p != q:
                other = at[q]                   # the qubit sitting on q's home position
                swaps.append(compile_gate(_SWAP_MATRIX, [p, q]))
                pos[q], pos[other] = q, p
                at[q], at[p] = q, other
        for group in _planner.plan(swaps, self.n, self.tile_bits, self.low_bits):
            self.state.apply_group(group.tile, group.ops)
            self.groups_launched += 1
        self.pos_of = list(range(self.n))

    def norm2(self):
        self.flush()
        return self.state.norm2()

    def clone_state(self):
        self.restore_layout()
        return self.state.clone()

    def clone(self):
        """An independent engine holding a copy of this state (identity layout)."""
        dup = Engine.from_state(self.clone_state(), self.tile_bits, self.low_bits)
        dup.relabel = self.relabel
        return dup

    # -- gates -------------------------------------------------------------------------------
    def queue_matrix(self, matrix, qubits):
        """Queue a k-qubit matrix on logical qubits (first = most significant, reference order)."""
        qubits = [int(q) for q in qubits]
        validate_qubits(qubits, self.n)
        self.queue(compile_gate(matrix, qubits))

    def queue(self, op):
        self.gates_queued += 1
        if op.kind != nat.OP_NOP:
            for piece in expand_op(op):
                if not self._merger.absorb(self.pending, piece):
                    self.pending.append(piece)

    def queue_shifted(self, pieces, shift):
        """Queue one gate given as already expanded pieces (``expand_op``) whose positions are all to be
        moved up by ``shift`` -- the same channel on qubit after qubit without re-classifying it."""
        self.gates_queued += 1
        for piece in pieces:
            if shift:
                piece = nat.ClassifiedOp(piece.kind, tuple(t + shift for t in piece.targets), piece.ctrl_mask << shift,
                                         piece.mat)
            if not self._merger.absorb(self.pending, piece):
                self.pending.append(piece)

    #: Ahead of a long flush on a large state the planner is walked once WITHOUT the device: every group's
    #: specialised kernel goes to the background NVRTC threads at once (cubins via the disk cache), so that the
    #: groups launched a second or two into the run already find theirs instead of being interpreted -- the first
    #: call of a never-seen program then runs mostly on specialised kernels.  Costs one extra planner walk on the
    #: host (~0.1 s for 1300 gates); skipped when the identity layout is not where the flush starts.
    AHEAD_MIN_QUBITS = int(os.environ.get("QVM_B200_AHEAD_MIN_QUBITS", "29"))
    AHEAD_MIN_OPS = 300
    #: > 0: the structures of the passes that start within this many seconds of the run are compiled last
    #: (``qvm_jit_defer``).  Measured on the 33-qubit cold call: no gain (2151 ms against 2034 ms without; the late
    #: compilations then compete with the sampling tail for the host cores), so it is off by default.
    AHEAD_COMPILE_S = float(os.environ.get("QVM_B200_AHEAD_COMPILE_S", "0"))

    def _compile_ahead(self, ops):
        """Returns the planner's recorded group choices (``_planner.Choices``, ready to replay) when it walked."""
        if (self.n < self.AHEAD_MIN_QUBITS or len(ops) < self.AHEAD_MIN_OPS or not self.relabel
                or not self.layout_is_identity() or getattr(self.state, "n_global", 0)):
            return None
        signature = hash((self.n, self.tile_bits, self.low_bits,
                          tuple((o.kind, o.targets, o.ctrl_mask) for o in ops)))
        if signature in _walked_ahead:
            return None              # this gate stream has been through here before: its kernels exist or are on their way
        if len(_walked_ahead) > 4096:
            _walked_ahead.clear()
        _walked_ahead.add(signature)
        choices = _planner.Choices()
        try:
            from referenceqvm import precompile
            if os.environ.get("QVM_B200_JIT", "1") != "1" or not nat.jit_cache_dir():
                return None
            # the run starts right behind this walk; launches wait for a kernel that is on its way for as long as the
            # GPU has earlier passes queued (qvm_capi.cu, launch_rec).  Optionally (AHEAD_COMPILE_S) the kernels of the
            # first passes -- which go through the interpreter whatever happens -- are compiled last.
            nat.jit_defer(int(self.AHEAD_COMPILE_S / (2.0 ** self.n * 32.0 / 2.7e12)))
            try:
                precompile.walk_pending(ops, self.n, self.tile_bits, self.low_bits, wait=False, choices=choices)
            finally:
                nat.jit_defer(0)
        except Exception:            # an optimisation only: the run itself reports real errors
            return None
        return choices.replay()

    def flush(self):
        self._merger.reset()
        if not self.pending:
            return
        ops, self.pending = self.pending, []
        choices = self._compile_ahead(ops)      # (the walk's group choices: the run need not search again)
        if self.relabel and self.n > self.tile_bits:
            self.groups_launched += _planner.run_relabelled(ops, self.n, self.pos_of, self._launch_relabelled,
                                                            self.tile_bits, self.low_bits, choices=choices)
            return
        if not self.layout_is_identity():
            ops = [o.remapped(self.pos_of) for o in ops]
        for group in _planner.plan(ops, self.n, self.tile_bits, self.low_bits):
            self.state.apply_group(group.tile, group.ops)
            self.groups_launched += 1

    # -- compiled programs (referenceqvm/_compiled.py) -----------------------------------------------
    def begin_recording(self):
        """From a freshly reset state: keep every launch of the flushes that follow for replay."""
        if self.pending or not self.layout_is_identity():
            raise RuntimeError("recording starts from a reset engine")
        self.state.record_begin()
        self._rec_mark = (self.gates_queued, self.groups_launched)

    def end_recording(self, keep=True):
        try:
            if keep:
                self.flush()
        except Exception:
            self.state.record_end(keep=False)
            raise
        plan = self.state.record_end(keep=keep)
        if plan is None:
            return None
        return CompiledFlush(plan, list(self.pos_of), self.n, self.gates_queued - self._rec_mark[0],
                             self.groups_launched - self._rec_mark[1])

    def replay(self, compiled, reset_first=True):
        """Reset (optionally) and run a compiled gate stream: one library call."""
        if compiled.n != self.n:
            raise ValueError("compiled program is for %d positions, engine has %d" % (compiled.n, self.n))
        self.pending = []
        self._merger.reset()
        self.state.plan_run(compiled.plan, reset_first)
        self.pos_of = list(compiled.pos_of)
        self.gates_queued += compiled.gates
        self.groups_launched += compiled.groups

    def _launch_relabelled(self, tile_positions, ops, bring_low):
        return self.state.apply_group(tile_positions, ops, bring_low=bring_low if bring_low else [],
                                      n_low=self.low_bits)

    # -- measurement -------------------------------------------------------------------------
    def measure(self, qubit, r):
        """MEASURE with a host-drawn uniform r: outcome 0 iff r < p0; collapse + renormalise."""
        self.flush()
        return self.state.measure(self.pos_of[int(qubit)], float(r))

    #: a run of MEASUREs is fused into joint-distribution passes from this many qubits on (below, the
    #: state is a few KiB and the single-qubit kernel's launch latency is all there is)
    MEASURE_MANY_MIN_QUBITS = 10

    def measure_many(self, qubits, uniforms):
        """A run of MEASUREs on distinct qubits, in order, each with its own host-drawn uniform: the
        reference's sequence of single measurements (qvm_wavefunction.py:147-156: p0 on the current,
        already collapsed state; outcome 0 iff r < p0; scale 1/sqrt(p0) or 1/sqrt(1 - p0)) replayed on the
        joint distribution of the measured bits, which costs one pass per chunk of qubits instead of two
        per qubit.  Returns [(outcome, p0), ...]."""
        qubits = [int(q) for q in qubits]
        if len(set(qubits)) != len(qubits):
            raise ValueError("measure_many needs distinct qubits")
        if self.n < self.MEASURE_MANY_MIN_QUBITS or len(qubits) < 2:
            return [self.measure(q, r) for q, r in zip(qubits, uniforms)]
        self.flush()
        positions = [self.pos_of[q] for q in qubits]
        lanes = np.arange(32)[None, :]
        results = []
        # the projection so far: applied ONCE, after the last chunk; the chunks in between read only the amplitudes
        # that are still consistent with it (qvm_measure_hist, filter mode)
        pending = None
        cum_norm = 1.0
        i = 0
        while i < len(positions):
            chunk, high = [], []
            while i + len(chunk) < len(positions):
                p = positions[i + len(chunk)]
                if p >= 5:
                    if len(high) == 5:
                        break
                    high.append(p)
                chunk.append(p)
            hist = self.state.measure_hist(high, pending, apply=False)
            bins = np.arange(1 << len(high))[:, None]
            consistent = np.ones(hist.shape, dtype=bool)
            norm_factor, mask, value = 1.0, 0, 0
            for j, p in enumerate(chunk):
                bit = ((bins >> high.index(p)) & 1) if p >= 5 else ((lanes >> p) & 1)
                bit = np.broadcast_to(bit, hist.shape)
                p0 = norm_factor * float(hist[consistent & (bit == 0)].sum())
                outcome = 0 if float(uniforms[i + j]) < p0 else 1
                norm_factor *= 1.0 / (p0 if outcome == 0 else 1.0 - p0)
                consistent = consistent & (bit == outcome)
                mask |= 1 << p
                value |= outcome << p
                results.append((outcome, p0))
            cum_norm *= norm_factor
            if pending is not None:
                mask, value = mask | pending[0], value | pending[1]
            pending = (mask, value, float(np.sqrt(cum_norm)))
            i += len(chunk)
        self.state.measure_hist([], pending, want_hist=False)
        return results

    def prob_zero(self, qubit):
        self.flush()
        return self.state.prob_zero(self.pos_of[int(qubit)])

    # -- expectation values (referenceqvm/_expectation.py) -----------------------------------------------
    def pauli_expectation(self, ops):
        """<psi|P|psi> for the Pauli string ``ops`` = {logical qubit: 'X'|'Y'|'Z'}: one reduction pass."""
        from referenceqvm._expectation import masks
        self.flush()
        xmask, zmask, ny = masks(ops, lambda q: self.pos_of[q])
        return self.state.pauli_expectation(xmask, zmask) * (1j) ** ny

    def density_pauli_expectation(self, n_rho, ops):
        """tr(rho P) for the vectorised rho this engine holds: i^ny sum_i (-1)^popcount(i & zmask) rho[i, i ^ xmask]."""
        from referenceqvm._expectation import masks
        self.flush()
        xmask, zmask, ny = masks(ops, lambda q: q)
        vals = self.state.density_offdiag(n_rho, self.pos_of, xmask)
        return _signed_sum(vals, zmask) * (1j) ** ny

    def inner_product(self, other):
        """<self|other> of two engines of equal shape (both brought to the identity layout)."""
        self.restore_layout()
        other.restore_layout()
        return self.state.inner_product(other.state)

    def density_diag(self, n_rho):
        """diag(rho) (2^n doubles) of the vectorised rho this engine holds, read through the layout."""
        self.flush()
        return self.state.density_diag(n_rho, self.pos_of)

    def density_prob_zero(self, n_rho, qubit):
        self.flush()
        if self.layout_is_identity():
            return self.state.density_prob_zero(n_rho, int(qubit))
        diag = self.density_diag(n_rho)
        return float(diag[((np.arange(diag.size) >> int(qubit)) & 1) == 0].sum())

    def sample(self, uniforms, mode=0, n_rho=0, physical=False):
        """Inverse-CDF draws of basis-state indices from |psi|^2 (mode 0) or diag(rho) (mode 1).

        Draws walk the CDF in HBM order.  With ``physical=True`` the raw indices come back (bit
        ``position(q)`` is qubit q); otherwise they are converted to logical indices."""
        self.flush()
        if mode == 1 and not self.layout_is_identity():
            return sample_distribution(self.density_diag(n_rho), uniforms, self.device)
        total = self.state.probs_prepare(mode, n_rho)
        indices = self.state.sample(uniforms)
        if not physical and not self.layout_is_identity():
            phys = np.asarray(indices, dtype=np.uint64)
            indices = np.zeros_like(phys)
            for q, p in enumerate(self.pos_of):
                indices |= ((phys >> np.uint64(p)) & np.uint64(1)) << np.uint64(q)
        return indices, total

    def sync(self):
        self.flush()
        self.state.sync()

    def stats(self):
        s = self.state.stats()
        s["gates_queued"] = self.gates_queued
        s["groups_launched"] = self.groups_launched
        return s


def make_engine(n_positions, device=0):
    """Engine for an ``n_positions``-bit state: sharded over the process group when
    ``_sharded.enable()`` was called in this process (one rank per GPU), single-GPU otherwise."""
    from referenceqvm import _sharded
    ctx = _sharded.context()
    if ctx is not None and ctx.world > 1:
        return _sharded.ShardedEngine(n_positions, ctx)
    return Engine(n_positions, device=device)


class DeviceAmplitudes(object):
    """Lazy, sliceable view of amplitudes that stay in HBM (used when 2^n * 16 B is too large to
    hand back as one ndarray; reference returns ``self.wf`` itself, qvm_wavefunction.py:367)."""

    def __init__(self, engine):
        self._engine = engine
        self.shape = (1 << engine.n,)
        self.dtype = np.dtype(np.complex128)
        self.ndim = 1
        self.size = self.shape[0]

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        n = self.shape[0]
        if isinstance(key, slice):
            start, stop, step = key.indices(n)
            if step == 1:
                return self._engine.amplitudes(start, max(stop - start, 0))
            if step > 1:
                cnt = max((stop - start + step - 1) // step, 0)
                return self._engine.strided(start, step, cnt)
            raise IndexError("negative steps are not supported")
        idx = int(key)
        if idx < 0:
            idx += n
        if not (0 <= idx < n):
            raise IndexError(idx)
        return self._engine.amplitudes(idx, 1)[0]

    def __array__(self, dtype=None, copy=None):
        a = self._engine.amplitudes()
        return a if dtype is None else a.astype(dtype)
