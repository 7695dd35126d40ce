"""Sharded engine: one rank per GPU, amplitudes split by the top log2(G) qubits.

The reference has no distributed path.  Here each rank holds 2^(n-g) contiguous
amplitudes (g = log2 world size); physical positions >= n-g are *global*: their value is a bit of
the rank.  A host-side logical->physical qubit layout tracks remaps.

* Diagonal gates and controls on global positions need no communication: the kernels read the bit
  from the shard index (``qvm_set_shard``).
* A dense target on a global position blocks that op (and what depends on it); everything else keeps
  running in fused passes.  When nothing more can run, ONE exchange epoch swaps all wanted global qubits
  with local victims (Belady: farthest next dense use).  Three implementations, best first:

  - ``push`` (default): out of place.  Every rank reads its shard once and stores each amplitude where
    it will live -- its own alternate buffer or a peer's, over NVLink peer memory (CUDA IPC) -- with
    write-only traffic.  Whenever a fused gate pass precedes the epoch, that pass's own last round does
    the storing (``qvm_apply_group_push``): the exchange then costs no pass of its own.  One
    stream-ordered barrier per epoch.  Needs a second buffer of the shard's size.
  - ``p2p``: in place, one swap kernel per peer over peer memory (``qvm_exchange_block``); used when the
    alternate buffer does not fit (36 qubits on 8 GPUs: 128 GiB shards).
  - ``staged``: pack / send / recv / unpack through two staging buffers with torch.distributed
    (gloo in the CPU tests; ``QVM_B200_SWAP=nccl`` for A/B runs).

* MEASURE = local partial p0 -> all-reduce -> local collapse.  Sampling = local two-level CDF,
  all-gather of the shard totals, each rank resolves the shots that fall in its range.

All ranks run the same program (SPMD) and make identical planning decisions; random draws are
broadcast from rank 0.  Collectives go through a small ``comm`` object: ``TorchComm``
(torch.distributed: NCCL on GPUs, gloo on CPU) in production; the single-GPU test suite plugs in an
in-process communicator that runs the ranks as threads on one device (tests/_thread_comm.py).
"""
import os
import threading
import warnings

import numpy as np

from referenceqvm import _native as nat
from referenceqvm import _planner
from referenceqvm._engine import _LOW_BITS, _RELABEL, _TILE_BITS, DiagonalMerger, compile_gate, expand_op, validate_qubits

_tls = threading.local()       # per-thread context (in-process ranks of the test suite)
_default = None                # context of the process (set from the main thread)

MAX_PUSH_BITS = 3              # kMaxPushBits of the library: the push exchange serves up to 8 ranks


class TorchComm(object):
    """Collectives of the sharded engine on torch.distributed.  Tensors are made by the shard backend
    (CUDA tensors on the shard's stream for NCCL, CPU tensors for gloo)."""

    same_process = False

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._make = None
        self._token = None

    def bind(self, make_tensor):
        """``make_tensor(values, dtype=None)`` -> tensor on the device the backend computes on."""
        self._make = make_tensor
        self._token = make_tensor([0.0])

    def _global_rank(self, group_rank):
        return group_rank if self.group is None else self.dist.get_global_rank(self.group, group_rank)

    def allreduce(self, values, op="sum"):
        t = self._make(np.ascontiguousarray(values, dtype=np.float64))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN if op == "min" else self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allreduce_int64(self, values):
        t = self._make(np.ascontiguousarray(values, dtype=np.int64), dtype=self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allgather(self, values):
        """The same 1-D float64 array from every rank, concatenated in rank order."""
        mine = self._make(np.ascontiguousarray(values, dtype=np.float64))
        out = self._make(np.zeros(mine.numel() * self.world))
        self.dist.all_gather(list(out.split(mine.numel())), mine, group=self.group)
        return out.cpu().numpy()

    def allgather_bytes(self, raw):
        mine = self._make(np.frombuffer(bytes(raw), dtype=np.uint8).copy(), dtype=self.torch.uint8)
        parts = [self.torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine, group=self.group)
        return [bytes(p.cpu().numpy().tobytes()) for p in parts]

    def broadcast(self, values, src=0):
        t = self._make(np.ascontiguousarray(values, dtype=np.float64))
        self.dist.broadcast(t, src=self._global_rank(src), group=self.group)
        return t.cpu().numpy()

    def stream_barrier(self, shard=None):
        """Cross-rank barrier ordered on the shard's stream (a one-element all-reduce): kernels queued
        after it start only when every rank's kernels queued before it have finished.  The host does
        not wait."""
        self.dist.all_reduce(self._token, group=self.group)

    def sendrecv(self, send, recv, partner, reverse):
        dist = self.dist
        ops = [dist.P2POp(dist.isend, send, self._global_rank(partner), self.group),
               dist.P2POp(dist.irecv, recv, self._global_rank(partner), self.group)]
        if reverse:                 # fixed order on both sides of a pair keeps gloo from deadlocking
            ops.reverse()
        return dist.batch_isend_irecv(ops)


class ShardContext(object):
    def __init__(self, comm, device, backend_factory=None, engine_options=None):
        self.comm, self.device = comm, device
        #: shard backend / constructor options of the engines ``_engine.make_engine`` creates under this context
        #: (tests: the numpy backend with small tiles; production: CudaShard with the defaults)
        self.backend_factory = backend_factory
        self.engine_options = dict(engine_options or {})
        self.rank, self.world = comm.rank, comm.world
        self.group = getattr(comm, "group", None)
        self.n_global = int(self.world).bit_length() - 1
        if (1 << self.n_global) != self.world:
            raise ValueError("world size %d is not a power of two" % self.world)


def enable(device=None, group=None, comm=None, backend_factory=None, engine_options=None):
    """Shard every engine created from now on (by this thread; from the main thread: by the process)
    across the ranks of ``comm`` (default: a TorchComm on ``group`` / WORLD)."""
    global _default
    ctx = ShardContext(comm if comm is not None else TorchComm(group), 0 if device is None else device,
                       backend_factory, engine_options)
    _tls.ctx = ctx
    if threading.current_thread() is threading.main_thread():
        _default = ctx
    return ctx


def disable():
    global _default
    _tls.ctx = None
    if threading.current_thread() is threading.main_thread():
        _default = None


def context():
    ctx = getattr(_tls, "ctx", None)
    return ctx if ctx is not None else _default


class CudaShard(object):
    """Local backend of the product path: one DeviceState + torch CUDA tensors for the collectives."""

    def __init__(self, n_local, ctx):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.n_local = n_local
        torch.cuda.set_device(ctx.device)
        self.state = nat.DeviceState(n_local, ctx.device)
        self.state.set_shard(ctx.n_global, ctx.rank)
        # One dedicated stream carries the library's kernels AND is torch's current stream, so NCCL
        # operations (which order themselves against the current stream) see the kernels in order.
        # (The legacy default stream has handle 0, which qvm_set_stream reads as "own stream".)
        self.stream = torch.cuda.Stream(device=ctx.device)
        torch.cuda.set_stream(self.stream)
        self.state.set_stream(self.stream.cuda_stream)
        self.device = torch.device("cuda", ctx.device)
        self.peers = None           # rank -> (pointer of its buffer 0, of its buffer 1 or 0)
        self.peer_error = None
        self._ipc = False

    def close(self):
        self.close_peers()
        self.state.close()

    def alloc(self, n_amps):
        return self.torch.empty(2 * n_amps, dtype=self.torch.float64, device=self.device)

    def scalar_tensor(self, values, dtype=None):
        return self.torch.tensor(values, dtype=dtype or self.torch.float64, device=self.device)

    def apply_groups(self, ops, tile_bits, low_bits):
        n = 0
        for group in _planner.plan(ops, self.n_local, tile_bits, low_bits):
            self.state.apply_group(group.tile, group.ops)
            n += 1
        return n

    def apply_group(self, tile_positions, ops, bring_low, n_low, push=None):
        """One pass with in-tile relabelling; returns the new position of every tile position."""
        return self.state.apply_group(tile_positions, ops, bring_low=bring_low, n_low=n_low, push=push)

    def alloc_alt(self):
        """Second buffer of the shard's size (push exchange); False when it does not fit."""
        try:
            return bool(self.state.alt_alloc())
        except Exception as exc:
            self.peer_error = "%s: %s" % (type(exc).__name__, exc)
            return False

    def open_peers(self, comm, with_alt):
        """Make every peer's shard addressable from this rank: CUDA IPC handles between processes (the
        loads / stores then travel over NVLink), plain device pointers between in-process ranks.  With
        ``with_alt`` the alternate buffers too (push exchange).  Returns False -- and keeps the reason in
        ``peer_error`` -- when that is not possible; the caller falls back to the staged exchange.
        Every rank takes part in every collective of this method whatever fails locally."""
        me = self.ctx.rank
        self.close_peers()
        if comm.same_process:
            ptrs = comm.allgather_object((self.state.device_ptr(), self.state.alt_ptr() if with_alt else 0))
            self.peers = {r: ptrs[r] for r in range(comm.world) if r != me}
            self._ipc = False
            return True
        blank = b"\0" * 64
        try:
            cur = self.state.ipc_export()
            alt = self.state.ipc_export_alt() if with_alt else blank
        except Exception as exc:            # reported by the engine (as a warning), never swallowed
            self.peer_error = "%s: %s" % (type(exc).__name__, exc)
            cur = alt = blank
        all_cur = comm.allgather_bytes(cur)
        all_alt = comm.allgather_bytes(alt) if with_alt else None
        if self.peer_error is not None or blank in all_cur or (with_alt and blank in all_alt):
            self.peer_error = self.peer_error or "a peer could not export its shard"
            return False
        self.peers = {}
        self._ipc = True
        try:
            for r in range(comm.world):
                if r != me:
                    self.peers[r] = (self.state.ipc_open(all_cur[r]), 0)
                    if with_alt:
                        self.peers[r] = (self.peers[r][0], self.state.ipc_open(all_alt[r]))
        except Exception as exc:
            self.peer_error = "%s: %s" % (type(exc).__name__, exc)
            return False
        return True

    def close_peers(self):
        if self.peers and self._ipc:
            for ptrs in self.peers.values():
                for ptr in ptrs:
                    if ptr:
                        try:
                            self.state.ipc_close(ptr)
                        except Exception:
                            pass
        self.peers = None

    def drop_alt(self):
        self.state.alt_free()

    def exchange_half(self, bit, mine_value, partner, off, cnt):
        self.state.exchange_half(bit, mine_value, self.peers[partner][0], off, cnt)

    def exchange_block(self, local_bits, mine_value, peer_value, partner, off, cnt):
        self.state.exchange_block(local_bits, mine_value, peer_value, self.peers[partner][0], off, cnt)

    def exchange_push(self, local_bits, mine_value, dst_ptrs):
        self.state.exchange_push(local_bits, mine_value, dst_ptrs)

    def pack_half(self, bit, value, off, cnt, buf):
        self.state.pack_half(bit, value, off, cnt, buf.data_ptr())

    def unpack_half(self, bit, value, off, cnt, buf):
        self.state.unpack_half(bit, value, off, cnt, buf.data_ptr())

    def reset_zero(self, set_amp0):
        self.state.reset_zero(set_amp0)

    def set_state(self, amps):
        self.state.set_state(amps)

    def get_state(self, offset=0, count=None):
        return self.state.get_state(offset, count)

    def norm2(self):
        return self.state.norm2()

    def prob_zero(self, position):
        return self.state.prob_zero(position)

    def measure_hist(self, high_positions, collapse=None, want_hist=True, apply=True):
        return self.state.measure_hist(high_positions, collapse, want_hist, apply)

    def density_diag(self, n_rho, pos_of):
        return self.state.density_diag(n_rho, pos_of)

    def density_offdiag(self, n_rho, pos_of, col_xor):
        return self.state.density_offdiag(n_rho, pos_of, col_xor)

    def pauli_expectation(self, xmask, zmask):
        return self.state.pauli_expectation(xmask, zmask)

    def copy_from(self, other):
        self.state.copy_from(other.state)

    def collapse(self, position, outcome, scale):
        self.state.collapse(position, outcome, scale)

    def probs_prepare(self):
        return self.state.probs_prepare(0, 0)

    def sample(self, uniforms):
        return self.state.sample(uniforms)

    def sync(self):
        self.state.sync()

    def stats(self):
        return self.state.stats()

    def record_event(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        return e


class ShardedEngine(object):
    """Same surface as ``_engine.Engine`` for a state spread over all ranks."""

    CHUNK_AMPS = 1 << 26          # 1 GiB per staging buffer

    def __init__(self, n_positions, ctx=None, tile_bits=None, low_bits=None, backend_factory=None,
                 chunk_amps=None, exchange=None):
        self.ctx = ctx or context()
        if self.ctx is None:
            raise RuntimeError("sharding is not enabled")
        self.comm = self.ctx.comm
        self.n = int(n_positions)
        self.g = self.ctx.n_global
        self.n_local = self.n - self.g
        if self.n_local < 1:
            raise ValueError("%d qubits cannot be sharded over %d ranks" % (self.n, self.ctx.world))
        opts = self.ctx.engine_options
        tile_bits = opts.get("tile_bits") if tile_bits is None else tile_bits
        low_bits = opts.get("low_bits") if low_bits is None else low_bits
        exchange = exchange or opts.get("exchange")
        self.tile_bits = _TILE_BITS if tile_bits is None else tile_bits
        self.low_bits = _LOW_BITS if low_bits is None else low_bits
        self.shard = (backend_factory or self.ctx.backend_factory or CudaShard)(self.n_local, self.ctx)
        if hasattr(self.comm, "bind"):
            self.comm.bind(self.shard.scalar_tensor)
        self.chunk_amps = min(chunk_amps or self.CHUNK_AMPS, 1 << (self.n_local - 1))
        self._send = self._recv = None
        # how global<->local exchanges run (module docstring): "push" > "p2p" > "staged", each needing
        # what the next one does not; every rank must agree, so the choice is all-reduced
        want = exchange or os.environ.get("QVM_B200_SWAP", "push")
        want = {"nccl": "staged"}.get(want, want)
        self.exchange = "staged"
        self.exchange_note = None
        if want in ("push", "p2p") and hasattr(self.shard, "open_peers"):
            with_alt = want == "push" and self.g <= MAX_PUSH_BITS
            if with_alt:
                have = self.shard.alloc_alt()
                if not self._all_agree(have):
                    # the second buffer does not fit somewhere (36 qubits on 8 GPUs): in-place peer-memory swaps
                    self.exchange_note = "no room for the push exchange's second buffer (%d bytes per rank): " \
                                         "in-place peer-memory exchange" % (16 << self.n_local)
                    if have:
                        self.shard.drop_alt()
                    self.shard.peer_error = None
                    with_alt = False
            ok = bool(self.shard.open_peers(self.comm, with_alt))
            if self._all_agree(ok):
                self.exchange = "push" if with_alt else "p2p"
            else:
                self.exchange_note = "peer memory unavailable (%s): staged send/recv exchange" % (
                    self.shard.peer_error or "a peer could not map it")
            if self.exchange_note:
                warnings.warn("referenceqvm sharded engine, rank %d: %s" % (self.ctx.rank, self.exchange_note))
        self.p2p = self.exchange != "staged"       # (kept: bench / older callers read it)
        self._buf = 0                              # which of the two buffers is current (push mode; same on all ranks)
        self.pos_of = list(range(self.n))         # logical qubit -> physical position
        #: in-tile relabelling of local qubits (``_planner.run_relabelled``) when the backend offers it
        self.relabel = _RELABEL and hasattr(self.shard, "apply_group")
        self.pending = []                         # ClassifiedOps in LOGICAL coordinates
        self._merger = DiagonalMerger()
        self.gates_queued = 0
        self.groups_launched = 0
        self.swaps = 0
        self.fused_swaps = 0
        self.swap_bytes = 0
        #: when a dict {"batch": [], "swap": [], "fused": []}: (start, end) device events around every
        #: kernel batch, every stand-alone exchange and every fused pass+exchange are appended (bench.py
        #: reads them after a sync)
        self.timing = None

    def _all_agree(self, flag):
        return bool(self.comm.allreduce([1.0 if flag else 0.0], op="min")[0] > 0.5)

    # -- lifetime / state ------------------------------------------------------------------------
    @property
    def state(self):
        return self.shard.state

    def close(self):
        self.pending = []
        self._merger.reset()
        self._send = self._recv = None
        if self.shard is not None:
            self.shard.close()
            self.shard = None

    def reset(self):
        self.pending = []
        self._merger.reset()
        self.pos_of = list(range(self.n))
        self.shard.reset_zero(self.ctx.rank == 0)

    def set_amplitudes(self, amplitudes):
        """Every rank passes the FULL vector (tests / small states); each keeps its shard."""
        a = np.asarray(amplitudes).reshape(-1)
        if a.size != (1 << self.n):
            raise ValueError("state size mismatch")
        self.pending = []
        self._merger.reset()
        self.pos_of = list(range(self.n))
        lo = self.ctx.rank << self.n_local
        self.shard.set_state(a[lo:lo + (1 << self.n_local)])

    def local_amplitudes(self):
        self.flush()
        return self.shard.get_state()

    def amplitudes(self, offset=0, count=None):
        """Amplitudes [offset, offset+count) in LOGICAL order, on every rank.  Each rank fills the
        part of the range it owns and the pieces are summed (x + 0 is exact)."""
        self.flush()
        self.restore_layout()
        total = 1 << self.n
        count = total - offset if count is None else int(count)
        if offset < 0 or count < 0 or offset + count > total:
            raise ValueError("range outside the state")
        out = np.zeros(count, dtype=np.complex128)
        lo = self.ctx.rank << self.n_local
        a, b = max(offset, lo), min(offset + count, lo + (1 << self.n_local))
        if b > a:
            out[a - offset:b - offset] = self.shard.get_state(a - lo, b - a)
        return self.comm.allreduce(out.view(np.float64)).view(np.complex128)

    def strided(self, offset, stride, count):
        self.flush()
        self.restore_layout()
        idx = offset + stride * np.arange(int(count), dtype=np.int64)
        out = np.zeros(int(count), dtype=np.complex128)
        mine = np.nonzero((idx >> self.n_local) == self.ctx.rank)[0]
        if mine.size:
            first = int(idx[mine[0]]) - (self.ctx.rank << self.n_local)
            out[mine] = self.shard.state.get_strided(first, stride, mine.size)
        return self.comm.allreduce(out.view(np.float64)).view(np.complex128)

    def probe(self, logical_indices):
        """Amplitudes at a handful of LOGICAL basis-state indices, identical on every rank, read through
        the layout without moving an amplitude (``amplitudes()`` would restore the layout first): the
        cross-check between runs on different numbers of GPUs (bench.py ``check.probe``)."""
        self.flush()
        idx = np.asarray(logical_indices, dtype=np.int64).reshape(-1)
        out = np.zeros(idx.size, dtype=np.complex128)
        for j, logical in enumerate(idx):
            phys = 0
            for q in range(self.n):
                phys |= ((int(logical) >> q) & 1) << self.pos_of[q]
            if (phys >> self.n_local) == self.ctx.rank:
                out[j] = self.shard.get_state(phys & ((1 << self.n_local) - 1), 1)[0]
        return self.comm.allreduce(out.view(np.float64)).view(np.complex128)

    def norm2(self):
        self.flush()
        return self._allreduce_sum(self.shard.norm2())

    def clone(self):
        """An independent engine holding a copy of this state (same layout): density snapshots, apply_kraus."""
        self.flush()
        dup = ShardedEngine(self.n, self.ctx, self.tile_bits, self.low_bits, backend_factory=type(self.shard),
                            exchange=self.exchange)
        dup.shard.copy_from(self.shard)
        dup.pos_of = list(self.pos_of)
        dup.relabel = self.relabel
        return dup

    # -- expectation values ------------------------------------------------------------------------------------------
    def ensure_local(self, qubits):
        """Bring the listed logical qubits onto local positions (one exchange epoch when any of them is global)."""
        self.flush()
        wanted = [q for q in qubits if self.pos_of[q] >= self.n_local]
        if wanted:
            self._bring_local(wanted, [], [])

    def pauli_expectation(self, ops):
        """<psi|P|psi> for a Pauli string {logical qubit: letter}: X / Y factors are made local first, every shard
        reduces its part (``qvm_pauli_expectation``; Z factors on global positions are signs of the shard), the
        parts are summed."""
        from referenceqvm._expectation import masks
        self.ensure_local([q for q, letter in ops.items() if letter in ("X", "Y")])
        xmask, zmask, ny = masks(ops, lambda q: self.pos_of[q])
        part = self.shard.pauli_expectation(xmask, zmask)
        total = self.comm.allreduce([part.real, part.imag])
        return complex(total[0], total[1]) * (1j) ** ny

    def density_pauli_expectation(self, n_rho, ops):
        from referenceqvm._engine import _signed_sum
        from referenceqvm._expectation import masks
        self.flush()
        xmask, zmask, ny = masks(ops, lambda q: q)
        vals = self.shard.density_offdiag(n_rho, self.pos_of, xmask)
        vals = self.comm.allreduce(vals.view(np.float64)).view(np.complex128)
        return _signed_sum(vals, zmask) * (1j) ** ny

    # -- density matrices: the state is a vectorised rho of n_rho qubits (rows = the high n_rho vector qubits, so
    # the global positions are row bits) ----------------------------------------------------------------------------
    def density_diag(self, n_rho):
        """diag(rho), identical on every rank: each shard gathers the entries it owns through the layout
        (``qvm_density_diag``), the pieces add up exactly (x + 0)."""
        self.flush()
        return self.comm.allreduce(self.shard.density_diag(n_rho, self.pos_of))

    def density_prob_zero(self, n_rho, qubit):
        diag = self.density_diag(n_rho)
        return float(diag[((np.arange(diag.size) >> int(qubit)) & 1) == 0].sum())

    # -- gates -------------------------------------------------------------------------------------
    def queue_matrix(self, matrix, qubits):
        qubits = [int(q) for q in qubits]
        validate_qubits(qubits, self.n)
        self.queue(compile_gate(matrix, qubits))

    def queue(self, op):
        self.gates_queued += 1
        if op.kind != nat.OP_NOP:
            for piece in expand_op(op):
                if not self._merger.absorb(self.pending, piece):
                    self.pending.append(piece)

    def flush(self):
        """Run every queued gate.  Ops whose dense targets are all local run in fused batches; an op
        with a dense target on a global position blocks (with everything that depends on it) until an
        exchange epoch brings all currently wanted global qubits in -- one multi-bit swap per epoch
        instead of one swap (and one broken batch) per gate.  In push mode the batch's last pass does
        the epoch's exchange itself."""
        self._merger.reset()
        if not self.pending:
            return
        ops, self.pending = self.pending, []
        self._compile_ahead(ops)
        deps = _planner.dependencies(ops)
        n = len(ops)
        done = [False] * n
        remaining = n
        while remaining:
            batch, in_batch, wanted = [], set(), []
            for i in range(n):
                if done[i]:
                    continue
                if any(not (done[d] or d in in_batch) for d in deps[i]):
                    continue
                op = ops[i]
                if op.kind == nat.OP_DENSE:
                    need = [t for t in op.targets if self.pos_of[t] >= self.n_local]
                    if need:
                        wanted.extend(t for t in need if t not in wanted)
                        continue
                batch.append(i)
                in_batch.add(i)
            exchange_follows = bool(wanted) and len(batch) < remaining
            fused = {}

            def plan_push(leftover, tile_positions, batch=batch, wanted=wanted, fused=fused):
                """Called by the planner right before the batch's LAST launch: choose the epoch's victims
                (knowing which ops will have run) and hand the launch its push."""
                keep = set(leftover)
                after = list(done)
                for j, i in enumerate(batch):
                    if j not in keep:
                        after[i] = True
                victims = self._choose_victims(wanted, ops, after, avoid=set(tile_positions))
                fused["pairs"] = self._push_pairs(wanted, victims)
                return self._push_args(fused["pairs"])

            use_fused = (exchange_follows and self.exchange == "push" and batch and self.relabel
                         and self.n_local > self.tile_bits and os.environ.get("QVM_B200_FUSED_PUSH", "1") != "0")
            # a thin last group is not worth a pass of its own when an exchange follows: its ops stay
            # pending and merge with what the exchange unblocks (they are logical ops, the layout is looked
            # up when they finally run)
            deferred = self._run_logical([ops[i] for i in batch], defer=exchange_follows,
                                         on_last=plan_push if use_fused else None)
            if deferred:
                keep = set(deferred)              # positions within `batch`
                batch = [i for j, i in enumerate(batch) if j not in keep]
            for i in batch:
                done[i] = True
            remaining -= len(batch)
            if remaining:
                if not wanted:
                    raise RuntimeError("sharded scheduler made no progress")
                if "pairs" in fused:
                    self._push_finish(fused["pairs"], fused_pass=True)
                else:
                    self._bring_local(wanted, ops, done)

    def _compile_ahead(self, ops):
        """First sight of a long gate stream on large shards (as ``Engine._compile_ahead``): walk the schedule
        without a device so NVRTC builds every group's kernel on the host cores while the first passes run on the
        interpreter; launches then wait for kernels that are on their way only while the GPU has passes queued."""
        from referenceqvm import _engine
        if (self.n_local < 26 or len(ops) < _engine.Engine.AHEAD_MIN_OPS or not self.relabel
                or getattr(self.shard, "dry", False) or not hasattr(self.shard, "state")
                or self.pos_of != list(range(self.n)) or self.exchange == "staged"
                or os.environ.get("QVM_B200_JIT", "1") != "1"):
            return
        signature = hash((self.n, self.ctx.world, self.tile_bits, self.low_bits, self.exchange,
                          tuple((o.kind, o.targets, o.ctrl_mask) for o in ops)))
        if signature in _engine._walked_ahead:
            return
        if len(_engine._walked_ahead) > 4096:
            _engine._walked_ahead.clear()
        _engine._walked_ahead.add(signature)
        try:
            from referenceqvm import precompile
            if not nat.jit_cache_dir():
                return
            # the processes of one node (torchrun: LOCAL_WORLD_SIZE of them) share the disk cache and walk the same
            # schedule: each compiles every LOCAL_WORLD_SIZE-th structure and picks the others' cubins up from disk
            stride, offset = 1, 0
            if not getattr(self.comm, "same_process", False):
                try:
                    stride, offset = int(os.environ["LOCAL_WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
                except (KeyError, ValueError):
                    stride, offset = 1, 0
            if not 0 <= offset < stride:
                stride, offset = 1, 0
            nat.jit_share(stride, offset)
            nat.jit_defer(int(_engine.Engine.AHEAD_COMPILE_S / (2.0 ** self.n_local * 32.0 / 2.7e12)) // stride)
            try:
                precompile.walk_sharded(ops, self.n, self.ctx.world, self.tile_bits, self.low_bits, self.exchange,
                                        wait=False)
            finally:
                nat.jit_defer(0)
                nat.jit_share(1, 0)
        except Exception:            # an optimisation only: the run itself reports real errors
            pass

    def _choose_victims(self, wanted, ops, done, avoid=()):
        """Local logical qubits to trade for the global ones in ``wanted``: those whose next dense use
        lies farthest ahead (Belady), lowest positions kept, positions in ``avoid`` (the tile of the pass
        that will carry the exchange) only as a last resort."""
        nxt = {}
        for i in range(len(ops) - 1, -1, -1):
            if not done[i] and ops[i].kind == nat.OP_DENSE:
                for t in ops[i].targets:
                    nxt[t] = i
        busy = set(wanted)
        for i, op in enumerate(ops):          # targets that must stay local next to the wanted ones
            if not done[i] and op.kind == nat.OP_DENSE and any(t in busy for t in op.targets):
                busy.update(op.targets)
        victims = []
        for q in wanted:
            v = self._pick_victim(nxt, busy, len(ops), avoid)
            victims.append(v)
            busy.add(v)
        return victims

    def _bring_local(self, wanted, ops, done):
        """Swap the logical qubits in ``wanted`` (all on global positions) with local victims."""
        victims = self._choose_victims(wanted, ops, done)
        if self.exchange == "push":
            self._swap_push(wanted, victims)
        elif self.exchange == "p2p" and len(wanted) > 1:
            self._swap_many(wanted, victims)
        else:
            for q, v in zip(wanted, victims):
                self._swap_global_local(q, v)

    # -- push exchange ---------------------------------------------------------------------------------
    def _push_pairs(self, globals_, victims):
        """[(victim qubit, global qubit)] sorted by the victims' positions (the library wants them ascending)."""
        return sorted(zip(victims, globals_), key=lambda vg: self.pos_of[vg[0]])

    def _push_args(self, pairs):
        """(victim positions, my global-bit value, destination pointers by victim value) for the library."""
        vpos = [self.pos_of[v] for v, _ in pairs]
        gbit = [self.pos_of[g] - self.n_local for _, g in pairs]
        rank = self.ctx.rank
        mine = sum(((rank >> gb) & 1) << i for i, gb in enumerate(gbit))
        base_rank = rank
        for gb in gbit:
            base_rank &= ~(1 << gb)
        alt = 1 - self._buf
        dst = []
        for value in range(1 << len(pairs)):
            peer = base_rank | sum(((value >> i) & 1) << gb for i, gb in enumerate(gbit))
            dst.append(0 if peer == rank else self.shard.peers[peer][alt])
        return vpos, mine, dst

    def _push_finish(self, pairs, fused_pass=False):
        """After the pushing pass: ONE stream-ordered barrier (every rank's stores have landed before
        anyone reads its new current buffer; the previous epoch's barrier already keeps writers away from
        a buffer that is still being read), then the bookkeeping."""
        self.comm.stream_barrier(self.shard)
        self._buf ^= 1
        m = len(pairs)
        for v, g in pairs:
            self.pos_of[g], self.pos_of[v] = self.pos_of[v], self.pos_of[g]
        self.swaps += 1
        self.fused_swaps += 1 if fused_pass else 0
        self.swap_bytes += 16 * (1 << (self.n_local - m)) * ((1 << m) - 1)

    def _swap_push(self, globals_, victims):
        e0 = self.shard.record_event() if self.timing is not None else None
        pairs = self._push_pairs(globals_, victims)
        vpos, mine, dst = self._push_args(pairs)
        self.shard.exchange_push(vpos, mine, dst)
        self._push_finish(pairs)
        if e0 is not None:
            self.timing["swap"].append((e0, self.shard.record_event()))

    # -- in-place peer-memory exchange -------------------------------------------------------------------
    def _swap_many(self, globals_, victims):
        """One exchange for m (global qubit, local victim) pairs: with every peer P of my 2^m-rank
        subgroup I trade my sub-block "victim bits == P's global bits" for P's sub-block "victim bits
        == my global bits" -- (2^m - 1) / 2^m of the shard moves, once."""
        m = len(globals_)
        pairs = self._push_pairs(globals_, victims)
        vpos = [self.pos_of[v] for v, _ in pairs]
        gbit = [self.pos_of[g] - self.n_local for _, g in pairs]
        rank = self.ctx.rank
        mine = sum(((rank >> gb) & 1) << i for i, gb in enumerate(gbit))
        base_rank = rank
        for gb in gbit:
            base_rank &= ~(1 << gb)
        block = 1 << (self.n_local - m)
        e0 = self.shard.record_event() if self.timing is not None else None
        self.shard.sync()          # (also makes a pending lazy reset real before a peer reads this shard)
        self.comm.stream_barrier(self.shard)
        for step in range(1, 1 << m):
            # XOR schedule: at every step the 2^m ranks of the subgroup pair up perfectly, so no GPU is
            # the target of two peers at once (all NVLink ports stay evenly loaded)
            value = mine ^ step
            peer = base_rank | sum(((value >> i) & 1) << gb for i, gb in enumerate(gbit))
            lo, cnt = (0, block // 2) if rank < peer else (block // 2, block - block // 2)
            self.shard.exchange_block(vpos, value, mine, peer, lo, cnt)
        self.comm.stream_barrier(self.shard)
        for v, g in pairs:
            self.pos_of[g], self.pos_of[v] = self.pos_of[v], self.pos_of[g]
        self.swaps += 1
        self.swap_bytes += 16 * block * ((1 << m) - 1)
        if e0 is not None:
            self.timing["swap"].append((e0, self.shard.record_event()))

    #: tile size of a pass that carries an exchange: it is bound by NVLink, so it takes everything that is left of
    #: its batch as soon as that fits one larger tile (0, the default: same tile as every other pass -- on the
    #: benchmark circuits the carrying pass is thin anyway: everything else is waiting for the exchange)
    PUSH_TILE_BITS = int(os.environ.get("QVM_B200_PUSH_TILE_BITS", "0"))
    #: a batch's last group is deferred across the following exchange when it holds fewer ops than this
    DEFER_TAIL_BELOW = int(os.environ.get("QVM_B200_DEFER_TAIL", "16"))

    def _run_logical(self, batch, defer=False, on_last=None):
        """Ops in LOGICAL coordinates whose dense targets are all local right now.  Returns the indices
        (into ``batch``) of the ops that were NOT run (a thin last group, only when ``defer``).
        ``on_last(leftover, tile_positions)`` -> push arguments for the batch's last launch (fused
        exchange)."""
        if not batch:
            return []
        if not (self.relabel and self.n_local > self.tile_bits):
            self._run_batch([op.remapped(self.pos_of) for op in batch])
            return []
        leftover = []
        e0 = self.shard.record_event() if self.timing is not None else None

        def launch(tile, ops, bring, push=None):
            if push is None:
                return self.shard.apply_group(tile, ops, bring if bring else [], self.low_bits)
            f0 = self.shard.record_event() if self.timing is not None else None
            out = self.shard.apply_group(tile, ops, bring if bring else [], self.low_bits, push=push)
            if f0 is not None:
                self.timing["fused"].append((f0, self.shard.record_event()))
            return out

        self.groups_launched += _planner.run_relabelled(
            batch, self.n_local, self.pos_of, launch, self.tile_bits, self.low_bits,
            defer_tail_below=self.DEFER_TAIL_BELOW if defer else 0, leftover=leftover, on_last=on_last,
            last_tile_bits=min(self.PUSH_TILE_BITS, self.n_local - self.g - 2),      # (room for the victims outside the tile)
            # a batch whose last pass carries the exchange is planned with the shallow group search: the deep one
            # (for shards of 2^32 amplitudes) leaves one- and two-round last passes, and those push at 520 GB/s
            # instead of 710 (33 qubits on 2 GPUs: 1494 gate-apps/s against 1636, profiles/r3a_*)
            search_depth=min(2, _planner.search_depth_for(self.n_local)) if on_last is not None else None)
        if e0 is not None:
            self.timing["batch"].append((e0, self.shard.record_event()))
        return leftover

    def _run_batch(self, batch):
        """Ops in PHYSICAL coordinates, fixed layout."""
        if batch:
            e0 = self.shard.record_event() if self.timing is not None else None
            self.groups_launched += self.shard.apply_groups(batch, self.tile_bits, self.low_bits)
            if e0 is not None:
                self.timing["batch"].append((e0, self.shard.record_event()))

    def _pick_victim(self, next_use, busy, horizon, avoid=()):
        """Local logical qubit to push out: the one whose next dense use is farthest (Belady).
        The lowest local positions are kept (they are every tile's coalescing bits)."""
        best, best_key = None, None
        for q in range(self.n):
            p = self.pos_of[q]
            if p >= self.n_local or q in busy:
                continue
            key = (0, next_use.get(q, horizon + 1), p)     # farthest use first, then highest position
            if p in avoid:
                key = (-1, next_use.get(q, horizon + 1), p)  # inside the carrying pass's tile: would unfuse the push
            if p < min(self.low_bits, self.n_local - 1):
                key = (-2, 0, p)                            # avoid unless nothing else is left
            if best_key is None or key > best_key:
                best, best_key = q, key
        if best is None:
            raise RuntimeError("no local qubit available for a global<->local swap")
        return best

    # -- the staged exchange -----------------------------------------------------------------------------
    def _buffers(self):
        if self._send is None:
            self._send = [self.shard.alloc(self.chunk_amps) for _ in range(2)]
            self._recv = [self.shard.alloc(self.chunk_amps) for _ in range(2)]
        return self._send, self._recv

    def _swap_global_local(self, q_global, q_local):
        """Swap the physical positions of logical qubits q_global (on a global position) and q_local."""
        pg, pl = self.pos_of[q_global], self.pos_of[q_local]
        assert pg >= self.n_local > pl
        if self.exchange == "push":
            self._swap_push([q_global], [q_local])
            return
        gbit = pg - self.n_local
        mine = (self.ctx.rank >> gbit) & 1
        partner = self.ctx.rank ^ (1 << gbit)
        e0 = self.shard.record_event() if self.timing is not None else None
        half = 1 << (self.n_local - 1)
        if self.exchange == "p2p":
            # stream-ordered cross-rank barrier (a tiny all-reduce), the swap kernel on my half of the
            # element range, barrier again: nobody touches a shard its partner is still working on
            self.shard.sync()      # (also makes a pending lazy reset real before a peer reads this shard)
            self.comm.stream_barrier(self.shard)
            lo = 0 if mine == 0 else half // 2
            cnt = half // 2 if mine == 0 else half - half // 2
            self.shard.exchange_half(pl, 1 - mine, partner, lo, cnt)
            self.comm.stream_barrier(self.shard)
            self.pos_of[q_global], self.pos_of[q_local] = pl, pg
            self.swaps += 1
            self.swap_bytes += 16 * half
            if e0 is not None:
                self.timing["swap"].append((e0, self.shard.record_event()))
            return
        send, recv = self._buffers()
        moving = 1 - mine           # the half whose local bit differs from my global bit changes owner
        pending = None              # (requests, offset, count, buffer index) of the chunk in flight
        k = 0
        for off in range(0, half, self.chunk_amps):
            cnt = min(self.chunk_amps, half - off)
            b = k & 1
            self.shard.pack_half(pl, moving, off, cnt, send[b])
            reqs = self.comm.sendrecv(send[b][:2 * cnt], recv[b][:2 * cnt], partner, reverse=bool(mine))
            if pending is not None:
                self._finish_chunk(pl, moving, pending)
            pending = (reqs, off, cnt, b)
            k += 1
        if pending is not None:
            self._finish_chunk(pl, moving, pending)
        self.pos_of[q_global], self.pos_of[q_local] = pl, pg
        self.swaps += 1
        self.swap_bytes += 16 * half
        if e0 is not None:
            self.timing["swap"].append((e0, self.shard.record_event()))

    def _finish_chunk(self, pl, moving, pending):
        reqs, off, cnt, b = pending
        for r in reqs:
            r.wait()
        self.shard.unpack_half(pl, moving, off, cnt, self._recv[b])

    def restore_layout(self):
        """Bring every logical qubit back to its own position (before a full readback)."""
        self.flush()
        # 1. global positions: swap each misplaced global qubit through a local slot
        for p in range(self.n - 1, self.n_local - 1, -1):
            holder = self.pos_of.index(p)
            if holder == p:
                continue
            # logical qubit p currently sits at some other position
            cur = self.pos_of[p]
            if cur >= self.n_local:
                # both global: route through a local victim
                victim = self._pick_victim({}, {p, holder}, 0)
                self._swap_global_local(p, victim)          # p becomes local
                cur = self.pos_of[p]
            self._swap_global_local(holder, p)              # holder (global at position p) <-> p (local)
        # 2. local permutation: plain SWAP gates on physical positions
        swap = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
        batch = []
        for q in range(self.n_local):
            p = self.pos_of[q]
            if p == q:
                continue
            other = self.pos_of.index(q)
            batch.append(compile_gate(swap, [p, q]))
            self.pos_of[q], self.pos_of[other] = q, p
        self._run_batch(batch)
        assert self.pos_of == list(range(self.n))

    # -- reductions ------------------------------------------------------------------------------------
    def _allreduce_sum(self, value):
        return float(self.comm.allreduce([float(value)])[0])

    def _broadcast_floats(self, values):
        return self.comm.broadcast(np.asarray(values, dtype=np.float64))

    def measure(self, qubit, r):
        """MEASURE: partial p0 on every rank, all-reduce, same decision everywhere, local collapse."""
        self.flush()
        r = float(self._broadcast_floats([r])[0])
        pos = self.pos_of[int(qubit)]
        p0 = self._allreduce_sum(self.shard.prob_zero(pos))
        outcome = 0 if r < p0 else 1
        scale = 1.0 / np.sqrt(p0 if outcome == 0 else 1.0 - p0)
        self.shard.collapse(pos, outcome, scale)
        return outcome, p0

    def measure_many(self, qubits, uniforms):
        """``Engine.measure_many`` for a sharded state: every rank builds the joint distribution of its own
        shard (local high positions x the 5 lowest index bits), the histograms are all-gathered -- a measured
        GLOBAL position is simply a bit of the rank -- and every rank replays the same sequential decisions
        on the same numbers (rank 0's uniforms), then applies its own part of the projection (a rank whose
        global bits lost keeps nothing).  Needs a backend with ``measure_hist``."""
        qubits = [int(q) for q in qubits]
        if len(set(qubits)) != len(qubits):
            raise ValueError("measure_many needs distinct qubits")
        if self.n_local < 10 or len(qubits) < 2 or not hasattr(self.shard, "measure_hist"):
            return [self.measure(q, r) for q, r in zip(qubits, uniforms)]
        self.flush()
        uniforms = self._broadcast_floats(np.asarray(uniforms, dtype=np.float64))
        rank, world, nl = self.ctx.rank, self.ctx.world, self.n_local
        positions = [self.pos_of[q] for q in qubits]
        lanes = np.arange(32)[None, None, :]
        ranks = np.arange(world)[:, None, None]
        results = []
        pending = None
        cum_norm, alive = 1.0, True
        i = 0
        while i < len(positions):
            chunk, high = [], []
            while i + len(chunk) < len(positions):
                p = positions[i + len(chunk)]
                if 5 <= p < nl:
                    if len(high) == 5:
                        break
                    high.append(p)
                chunk.append(p)
            local = self.shard.measure_hist(high, pending, apply=False)
            hist = self.comm.allgather(local.reshape(-1)).reshape((world,) + local.shape)
            bins = np.arange(local.shape[0])[None, :, None]
            consistent = np.ones(hist.shape, dtype=bool)
            norm_factor, mask, value = 1.0, 0, 0
            for j, p in enumerate(chunk):
                if p >= nl:
                    bit = (ranks >> (p - nl)) & 1
                elif p >= 5:
                    bit = (bins >> high.index(p)) & 1
                else:
                    bit = (lanes >> p) & 1
                bit = np.broadcast_to(bit, hist.shape)
                p0 = norm_factor * float(hist[consistent & (bit == 0)].sum())
                outcome = 0 if float(uniforms[i + j]) < p0 else 1
                norm_factor *= 1.0 / (p0 if outcome == 0 else 1.0 - p0)
                consistent = consistent & (bit == outcome)
                if p >= nl:
                    alive = alive and ((rank >> (p - nl)) & 1) == outcome
                else:
                    mask |= 1 << p
                    value |= outcome << p
                results.append((outcome, p0))
            # the projection so far is applied once, after the last chunk; until then the histograms are taken over
            # the amplitudes still consistent with it (qvm_measure_hist, filter mode: only those are read)
            cum_norm *= norm_factor
            if pending is not None:
                mask, value = mask | pending[0], value | pending[1]
            pending = (mask, value, float(np.sqrt(cum_norm)) if alive else 0.0)
            i += len(chunk)
        self.shard.measure_hist([], pending, want_hist=False)
        return results

    def prob_zero(self, qubit):
        self.flush()
        return self._allreduce_sum(self.shard.prob_zero(self.pos_of[int(qubit)]))

    def position(self, qubit):
        """Physical position (index bit) of a logical qubit: for bit look-ups in ``sample(physical=True)``."""
        return self.pos_of[qubit]

    def sample(self, uniforms, mode=0, n_rho=0, physical=False, logical_order=False):
        """Inverse-CDF draws of LOGICAL basis-state indices (``physical=True``: of HBM indices, bit
        ``position(q)`` being qubit q), identical on every rank.

        The CDF is walked in HBM order (rank-major, then the shard's physical index) and the drawn
        indices are translated through the layout -- the same distribution as a walk in logical order
        without moving a single amplitude.  ``logical_order=True`` restores the layout first (exchanges
        and SWAP passes), so that the draws equal ``searchsorted`` on the logical CDF uniform by uniform."""
        if mode == 1:
            # diag(rho) is small (2^n_rho doubles): every rank gathers it and runs the same device sampler on it
            from referenceqvm._engine import sample_distribution
            u = self._broadcast_floats(uniforms)
            diag = self.density_diag(n_rho)
            if hasattr(self.shard, "sample_distribution"):      # (the numpy backend of the CPU tests)
                return self.shard.sample_distribution(diag, u)
            return sample_distribution(diag, u, self.ctx.device)
        self.flush()
        if logical_order:
            self.restore_layout()
        u = self._broadcast_floats(uniforms)
        local_total = self.shard.probs_prepare()
        totals = self.comm.allgather([local_total])
        cum = np.cumsum(totals)
        grand = cum[-1]
        target = u * grand
        # owner[i] = number of shards whose cumulative total is <= target[i]: a comparison per shard boundary
        # (np.searchsorted spends ~35 ns per needle whatever the length of the array: 36 ms for 10^6 shots)
        owner = np.zeros(u.size, dtype=np.int16)
        above = np.empty(u.size, dtype=bool)
        for boundary in cum[:-1]:
            np.greater_equal(target, boundary, out=above)
            np.add(owner, above.view(np.int8), out=owner)
        sel = np.nonzero(owner == self.ctx.rank)[0]
        phys = np.zeros(u.size, dtype=np.int64)
        if sel.size:
            start = cum[self.ctx.rank] - totals[self.ctx.rank]
            lu = np.clip((target[sel] - start) / totals[self.ctx.rank], 0.0, np.nextafter(1.0, 0.0))
            local_idx = self.shard.sample(lu).astype(np.int64)
            phys[sel] = (np.int64(self.ctx.rank) << np.int64(self.n_local)) | local_idx
        phys = self.comm.allreduce_int64(phys).astype(np.uint64)
        if physical:                # the caller looks bits up through position(): no per-qubit pass over the shots
            return phys, grand
        logical = np.zeros_like(phys)
        for q in range(self.n):
            logical |= ((phys >> np.uint64(self.pos_of[q])) & np.uint64(1)) << np.uint64(q)
        return logical, grand

    def sync(self):
        self.flush()
        self.shard.sync()

    def stats(self):
        s = self.shard.stats()
        s.update(gates_queued=self.gates_queued, groups_launched=self.groups_launched, swaps=self.swaps,
                 fused_swaps=self.fused_swaps, swap_bytes=self.swap_bytes, exchange=self.exchange)
        return s
