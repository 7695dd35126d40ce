"""Ahead-of-time compilation of a program's specialised kernels (no GPU needed).

    python -m referenceqvm.precompile --qubits 33 --depth 40 --seed 1234 --gpus 1 2 4 8

A fused group runs as an NVRTC-specialised kernel once its *structure* has been compiled
(csrc/qvm_jit.h); until then the interpreted kernel serves it at about half the speed.  This tool walks
the same host path a run takes -- classification, diagonal merging, the fusion planner with in-tile
relabelling, for several GPUs the sharded scheduler with its exchange epochs -- but hands every group to
``qvm_spec_precompile`` instead of a device: the cubins land in the library's disk cache
(``<package>/_spec_cache``, or ``QVM_B200_CACHE_DIR``) and the first run of that program in any later
process starts on specialised kernels.  ``__graft_entry__.build()`` does this for the benchmark circuit.

Nothing here touches amplitudes: the "shard" below only plans.
"""
import argparse
import os
import sys
import time

import numpy as np

from referenceqvm import _engine, _native as nat, _planner, _sharded


class _DryComm(object):
    """Rank 0 of a world of G that never talks: every collective returns this rank's own contribution (all
    ranks of a real run take identical planning decisions)."""
    same_process = True

    def __init__(self, world):
        self.rank, self.world, self.group = 0, world, None

    def allreduce(self, values, op="sum"):
        return np.array(values, dtype=np.float64)

    def allreduce_int64(self, values):
        return np.array(values, dtype=np.int64)

    def allgather(self, values):
        return np.tile(np.array(values, dtype=np.float64).reshape(-1), self.world)

    def allgather_object(self, obj):
        return [obj] * self.world

    def broadcast(self, values, src=0):
        return np.array(values, dtype=np.float64)

    def stream_barrier(self, shard=None):
        pass


class _DryShard(object):
    """Plans like ``_sharded.CudaShard`` launches, compiles instead of running."""
    dry = True

    def __init__(self, n_local, ctx):
        self.n_local, self.ctx = n_local, ctx
        self.peers = None
        self.peer_error = None
        self.compiled = self.skipped = self.groups = 0

    def close(self):
        pass

    def scalar_tensor(self, values, dtype=None):
        return np.asarray(values)

    def alloc_alt(self):
        return True

    def drop_alt(self):
        pass

    def open_peers(self, comm, with_alt):
        self.peers = {r: (1, 1) for r in range(comm.world) if r != self.ctx.rank}
        return True

    def close_peers(self):
        self.peers = None

    def reset_zero(self, set_amp0):
        pass

    def apply_group(self, tile_positions, ops, bring_low, n_low, push=None):
        status, new_pos = nat.spec_precompile(self.n_local, tile_positions, ops, self.ctx.n_global, self.ctx.rank,
                                              bring_low or [], n_low)
        self.groups += 1
        if status == 0:
            self.compiled += 1
        else:
            self.skipped += 1
        return new_pos

    def apply_groups(self, ops, tile_bits, low_bits):
        n = 0
        for group in _planner.plan(ops, self.n_local, tile_bits, low_bits):
            self.apply_group(group.tile, group.ops, [], 0)
            n += 1
        return n

    def exchange_push(self, local_bits, mine_value, dst_ptrs):
        pass

    def exchange_block(self, local_bits, mine_value, peer_value, partner, off, cnt):
        pass

    def exchange_half(self, bit, mine_value, partner, off, cnt):
        pass

    def sync(self):
        pass

    def record_event(self):
        return None

    def stats(self):
        return {}


def walk_pending(pending, n_qubits, tile_bits, low_bits, wait=True, choices=None):
    """Single-GPU planner walk over an already merged op list (what ``Engine.flush`` is about to run): every group
    goes to ``qvm_spec_precompile``; with ``wait=False`` the compilations are left running on the library's
    background threads (``Engine.flush`` calls this ahead of a long run on a large state, see there).  ``choices``
    (a recording ``_planner.Choices``) receives the planner's group choices for the run to replay."""
    shard = _DryShard(n_qubits, _sharded.ShardContext(_DryComm(1), 0))
    pos_of = list(range(n_qubits))
    if _engine._RELABEL and n_qubits > tile_bits:
        _planner.run_relabelled(pending, n_qubits, pos_of,
                                lambda tile, gops, bring: shard.apply_group(tile, gops, bring if bring else [], low_bits),
                                tile_bits, low_bits, choices=choices)
    else:
        shard.apply_groups(pending, tile_bits, low_bits)
    if wait:
        nat.jit_wait()
    return shard


def walk_sharded(pending, n_qubits, gpus, tile_bits, low_bits, exchange, wait=True):
    """The sharded scheduler's walk (exchange epochs, victims, fused pushes) over an already merged op list, from
    the identity layout: what ``ShardedEngine.flush`` is about to run on every rank.  Group structures do not depend
    on the rank (the shard index is a launch parameter), so each rank can walk for itself."""
    ctx = _sharded.ShardContext(_DryComm(gpus), 0)
    eng = _sharded.ShardedEngine(n_qubits, ctx, tile_bits=tile_bits, low_bits=low_bits, backend_factory=_DryShard,
                                 exchange=exchange)
    eng.pending = list(pending)
    eng.flush()
    if wait:
        nat.jit_wait()
    return eng.shard


def precompile_ops(ops, n_qubits, gpus=1, tile_bits=None, low_bits=None, exchange=None):
    """Compile the specialised kernels a run of ``ops`` (ClassifiedOps on logical qubits, in program order) on
    ``gpus`` GPUs would use.  ``exchange``: the exchange mode the run will use ("push" unless the shard's second
    buffer does not fit in HBM, then "p2p"): victim choice, hence the later group structures, depend on it.
    Returns {"groups", "compiled", "left_to_interpreter", "seconds"}."""
    t0 = time.perf_counter()
    tile_bits = _engine._TILE_BITS if tile_bits is None else tile_bits
    low_bits = _engine._LOW_BITS if low_bits is None else low_bits
    if gpus == 1:
        pending, merger = [], _engine.DiagonalMerger()
        for op in ops:
            if op.kind != nat.OP_NOP:
                for piece in _engine.expand_op(op):
                    if not merger.absorb(pending, piece):
                        pending.append(piece)
        shard = walk_pending(pending, n_qubits, tile_bits, low_bits, wait=False)
    else:
        ctx = _sharded.ShardContext(_DryComm(gpus), 0)
        eng = _sharded.ShardedEngine(n_qubits, ctx, tile_bits=tile_bits, low_bits=low_bits, backend_factory=_DryShard,
                                     exchange=exchange or ("push" if ctx.n_global <= _sharded.MAX_PUSH_BITS else "p2p"))
        shard = eng.shard
        for op in ops:
            eng.queue(op)
        eng.flush()
    nat.jit_wait()                    # the cubins are built on the library's background threads
    return {"groups": shard.groups, "compiled": shard.compiled, "left_to_interpreter": shard.skipped,
            "seconds": time.perf_counter() - t0}


def precompile_spec(spec, n_qubits, gpus=1, gate_set=None, exchange=None):
    """``spec`` = [(gate name, params, qubits)] as produced by workloads.py."""
    from referenceqvm.gates import gate_matrix
    gate_set = gate_set or gate_matrix
    ops = []
    for name, params, qubits in spec:
        m = gate_set[name]
        ops.append(_engine.compile_gate(np.asarray(m(*params) if params else m), list(qubits)))
    return precompile_ops(ops, n_qubits, gpus, exchange=exchange)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=33)
    ap.add_argument("--depth", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--gpus", type=int, nargs="+", default=[1])
    args = ap.parse_args()
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if here not in sys.path:
        sys.path.insert(0, here)
    from workloads import random_circuit_spec
    spec = random_circuit_spec(args.qubits, args.depth, args.seed)
    for g in args.gpus:
        print(g, precompile_spec(spec, args.qubits, g), nat.jit_cache_dir())


if __name__ == "__main__":
    main()
