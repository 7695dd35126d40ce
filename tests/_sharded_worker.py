"""Worker of tests/test_sharded_gloo.py: one rank of a gloo process group driving the sharded
engine's host logic on the numpy shard backend, diffed against the oracle on every rank."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, os.path.join(REPO, "reference-qvm_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

import referenceqvm  # noqa: E402,F401
from _numpy_shard import NumpyShard  # noqa: E402
from oracle import qvm_oracle as orc  # noqa: E402
from referenceqvm import _sharded, gates as G  # noqa: E402


def oracle_ops(spec):
    out = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        out.append(("gate", np.asarray(m(*params) if params else m, dtype=complex), qubits))
    return out


def main():
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["QVM_TEST_PORT"],
                            rank=int(os.environ["QVM_TEST_RANK"]), world_size=int(os.environ["QVM_TEST_WORLD"]))
    ctx = _sharded.enable(device=0)
    rank, world = ctx.rank, ctx.world
    checked = 0
    for n, depth, seed, chunk in ((8, 6, 1, 4), (9, 8, 2, 1 << 20), (10, 5, 1234, 16), (7, 10, 5, 1)):
        spec = orc.random_circuit_spec(n, depth, seed)
        eng = _sharded.ShardedEngine(n, ctx, tile_bits=5, low_bits=2, backend_factory=NumpyShard, chunk_amps=chunk)
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
        # shard-local view in the CURRENT layout: permute the oracle's state the same way
        eng.flush()
        assert eng.swaps > 0, "circuit never touched a global qubit densely"
        assert abs(eng.norm2() - 1.0) < 1e-12
        p0 = eng.prob_zero(n - 1)
        assert abs(p0 - orc.fast_prob_zero(want, n - 1, n)) < 1e-12
        got = eng.amplitudes()
        assert eng.pos_of == list(range(n))
        err = float(np.max(np.abs(got - want)))
        assert err < 1e-12, (n, depth, seed, err)
        part = eng.amplitudes(3, 17)
        assert np.array_equal(part, got[3:20])

        # MEASURE on a global and a local logical qubit: same uniforms as the oracle
        psi = want
        for q, r in ((n - 1, 0.3), (0, 0.8), (n // 2, 0.5)):
            outcome, p0 = eng.measure(q, r if rank == 0 else -1.0)        # rank 0's draw wins
            o_want, p_want, psi = orc.fast_measure(psi, q, n, r)
            assert outcome == o_want and abs(p0 - p_want) < 1e-12
        err = float(np.max(np.abs(eng.amplitudes() - psi)))
        assert err < 1e-12, ("after measure", err)

        # sampling: identical indices on all ranks, equal to the oracle's inverse CDF (ties aside)
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        u = np.random.RandomState(seed).random_sample(257)
        probs = np.abs(want) ** 2
        # default: the CDF is walked in HBM order and indices are translated through the layout
        idx, total = eng.sample(u if rank == 0 else np.zeros(257))
        assert abs(total - 1.0) < 1e-12
        logical = np.arange(1 << n, dtype=np.int64)
        phys = np.zeros_like(logical)
        for q, p in enumerate(eng.pos_of):
            phys |= ((logical >> q) & 1) << p
        probs_hbm = np.empty_like(probs)
        probs_hbm[phys] = probs
        logical_of_phys = np.empty_like(logical)
        logical_of_phys[phys] = logical
        ref = logical_of_phys[np.asarray(orc.inverse_cdf_sample(probs_hbm, u), dtype=np.int64)]
        assert np.array_equal(idx.astype(np.int64), ref)
        # physical=True: raw HBM indices; bit position(q) is qubit q (what run_and_measure unpacks)
        raw, _ = eng.sample(u if rank == 0 else np.zeros(257), physical=True)
        rebuilt = np.zeros(raw.size, dtype=np.int64)
        for q in range(n):
            rebuilt |= ((raw.astype(np.int64) >> eng.position(q)) & 1) << q
        assert np.array_equal(rebuilt, ref)
        # on request the engine restores the logical layout first: the CDF order is the reference's
        idx, total = eng.sample(u if rank == 0 else np.zeros(257), logical_order=True)
        assert eng.pos_of == list(range(n))
        ref = orc.inverse_cdf_sample(probs, u)
        assert np.array_equal(idx.astype(np.int64), np.asarray(ref, dtype=np.int64))
        eng.close()
        checked += 1
    # runs of MEASUREs on a sharded state (local shards of >= 10 positions): joint distribution per shard,
    # all-gathered; measured global positions are rank bits; every rank takes the same decisions
    for n, depth, seed in ((12, 5, 3), (13, 4, 8)):
        spec = orc.random_circuit_spec(n, depth, seed)
        eng = _sharded.ShardedEngine(n, ctx, tile_bits=5, low_bits=2, backend_factory=NumpyShard)
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
        rs = np.random.RandomState(seed)
        order = [int(q) for q in rs.permutation(n)]
        uniforms = rs.random_sample(n)
        eng.flush()
        launches = eng.shard.launches
        got = eng.measure_many(order, uniforms if rank == 0 else np.zeros(n))      # rank 0's draws win
        assert any(eng.pos_of[q] >= eng.n_local for q in order)
        assert eng.shard.launches - launches <= (n + 4) // 5 + 2                     # chunks, not 2 passes per qubit
        psi = want
        for (outcome, p0), q, r in zip(got, order, uniforms):
            o_want, p_want, psi = orc.fast_measure(psi, q, n, r)
            assert outcome == o_want and abs(p0 - p_want) < 1e-12
        err = float(np.max(np.abs(eng.amplitudes() - psi)))
        assert err < 1e-12, ("after measure_many", err)
        eng.close()
        checked += 1
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d/%d ok (%d circuits)" % (rank, world, checked))


if __name__ == "__main__":
    main()
