"""Dry run of the host-side planning for a sharded random circuit: how many kernel groups and
global<->local swaps the engine would issue (no device, no process group)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "reference-qvm_b200")):
    sys.path.insert(0, p)

import referenceqvm  # noqa: E402,F401
from workloads import random_circuit_spec  # noqa: E402
from referenceqvm import _engine, _planner, _sharded, gates as G  # noqa: E402


class NullShard(object):
    def __init__(self, n_local, ctx):
        self.n_local = n_local
        self.ops_per_group = []

    def apply_groups(self, ops, tile_bits, low_bits):
        groups = _planner.plan(ops, self.n_local, tile_bits, low_bits)
        self.ops_per_group += [len(g.ops) for g in groups]
        return len(groups)

    def apply_group(self, tile, ops, bring_low, n_low):
        """Relabelling pass, modelled: every request is honoured."""
        self.ops_per_group.append(len(ops))
        tile = list(tile)
        new_pos = list(tile)
        incoming = [p for p in bring_low if p >= n_low]
        victims = [p for p in range(n_low - 1, -1, -1) if p not in bring_low][:len(incoming)]
        for a, b in zip(incoming, victims):
            new_pos[tile.index(a)], new_pos[tile.index(b)] = b, a
        return new_pos

    def alloc(self, n):
        return None

    def pack_half(self, *a):
        pass

    unpack_half = pack_half

    def record_event(self):
        return None


def main():
    n, depth, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    spec = random_circuit_spec(n, depth, 1234)
    ops = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        ops.append(_engine.compile_gate(m(*params) if params else m, qubits))
    if world == 1:
        t0 = time.time()
        groups = _planner.plan(ops, n, _engine._TILE_BITS, _engine._LOW_BITS)
        print("fixed layout: groups", len(groups), "ops", len(ops), "plan s", time.time() - t0)
        shard = NullShard(n, None)
        passes = _planner.run_relabelled(ops, n, list(range(n)), lambda t, o, b: shard.apply_group(t, o, b or [], _engine._LOW_BITS),
                                         _engine._TILE_BITS, _engine._LOW_BITS)
        print("relabelling: groups", passes)
        return
    ctx = _sharded.ShardContext(None, 0, world, 0)
    eng = _sharded.ShardedEngine(n, ctx, backend_factory=NullShard)
    eng._swap_global_local_real = eng._swap_global_local

    def fake_swap(qg, ql):
        pg, pl = eng.pos_of[qg], eng.pos_of[ql]
        eng.pos_of[qg], eng.pos_of[ql] = pl, pg
        eng.swaps += 1
    eng._swap_global_local = fake_swap
    eng.p2p = True
    moved = [0.0]

    def fake_many(gl, vi):
        for g, v in zip(gl, vi):
            eng.pos_of[g], eng.pos_of[v] = eng.pos_of[v], eng.pos_of[g]
        eng.swaps += 1
        moved[0] += 1.0 - 0.5 ** len(gl)
    eng._swap_many = fake_many
    real_fake = fake_swap

    def fake_swap2(qg, ql):
        real_fake(qg, ql)
        moved[0] += 0.5
    eng._swap_global_local = fake_swap2
    t0 = time.time()
    for op in ops:
        eng.queue(op)
    eng.flush()
    print("world", world, "groups", eng.groups_launched, "exchanges", eng.swaps, "shards moved per GPU %.2f" % moved[0],
          "ops", len(ops), "plan s %.3f" % (time.time() - t0))


if __name__ == "__main__":
    main()
