"""Host logic of the PUSH exchange on CPU: the real ``ShardedEngine`` (scheduler, victim choice, fused
last pass + exchange, buffer parity, layout tracking) with the ranks as threads of this process
(tests/_thread_comm.py) on the numpy shard backend; every result diffed against the oracle."""
import numpy as np
import pytest

from _numpy_shard import NumpyShard
from _thread_comm import run_ranks
from oracle import qvm_oracle as orc
from referenceqvm import _sharded, gates as G


def oracle_ops(spec):
    out = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        out.append(("gate", np.asarray(m(*params) if params else m, dtype=complex), qubits))
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_push_exchange_host_logic(world):
    seen = {}

    def rank_main(comm):
        ctx = _sharded.enable(device=0, comm=comm)
        try:
            for n, depth, seed in ((10, 6, 1), (11, 8, 2), (12, 5, 1234)):
                spec = orc.random_circuit_spec(n, depth, seed)
                eng = _sharded.ShardedEngine(n, ctx, tile_bits=5, low_bits=2, backend_factory=NumpyShard, exchange="push")
                assert eng.exchange == "push"
                eng.reset()
                for name, params, qubits in spec:
                    m = G.gate_matrix[name]
                    eng.queue_matrix(m(*params) if params else m, qubits)
                want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
                eng.flush()
                assert eng.swaps > 0
                assert abs(eng.norm2() - 1.0) < 1e-12
                probe_idx = [0, 1, 5, (1 << n) - 1, 1 << (n - 1), 77]
                assert np.max(np.abs(eng.probe(probe_idx) - want[probe_idx])) < 1e-12
                got = eng.amplitudes()
                assert np.max(np.abs(got - want)) < 1e-12, (n, depth, seed)
                if comm.rank == 0:
                    seen[(world, n)] = (eng.swaps, eng.fused_swaps, eng.shard.fused_pushes)
                psi = want
                for q, r in ((n - 1, 0.3), (0, 0.8)):
                    outcome, p0 = eng.measure(q, r if comm.rank == 0 else -1.0)
                    o_want, p_want, psi = orc.fast_measure(psi, q, n, r)
                    assert outcome == o_want and abs(p0 - p_want) < 1e-12
                assert np.max(np.abs(eng.amplitudes() - psi)) < 1e-12
                eng.close()
        finally:
            _sharded.disable()

    run_ranks(world, rank_main)
    # the fused path was exercised: some epochs rode on a gate pass
    assert sum(v[1] for v in seen.values()) > 0, seen
    assert all(v[1] == v[2] for v in seen.values()), seen
