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


def density_oracle(prog, n, p1q=None, p2q=None):
    from _programs import oracle_ops as prog_ops
    vec = np.zeros(4 ** n, dtype=complex)
    vec[0] = 1.0
    for _, m, q in prog_ops(prog):
        vec = orc.fast_density_gate(vec, m, q, n)
        if p1q is not None:
            p = p1q if len(q) == 1 else p2q
            for i in range(n):
                vec = orc.fast_apply_kraus(vec, G.dephasing(p), i, n)
    return vec


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_density_qvm_host_logic(world):
    """QVM_Density on a sharded engine (rho split by rows; numpy backend, thread ranks): density(), the layout-aware
    diagonal, MEASURE on rho, sampling from diag(rho), apply_kraus -- all against the oracle."""
    from _programs import spec_program
    from referenceqvm.api import QVMConnection
    from referenceqvm.qvm_density import INFINITY, NoiseModel

    def rank_main(comm):
        _sharded.enable(device=0, comm=comm, backend_factory=NumpyShard,
                        engine_options={"tile_bits": 5, "low_bits": 2, "exchange": "push"})
        try:
            n = 5
            spec = [("H", [], [q]) for q in range(n)] + orc.random_circuit_spec(n, 3, 1234)
            prog = spec_program(spec)
            nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
            qvm = QVMConnection(type_trans="density", noise_model=nm)
            rho = qvm.density(prog)
            assert isinstance(qvm._engine, _sharded.ShardedEngine) and qvm._engine.swaps > 0
            p1, p2 = 1.0 - np.exp(-50e-9 / 30e-6), 1.0 - np.exp(-150e-9 / 30e-6)
            want = orc.vec_to_rho(density_oracle(prog, n, p1, p2), n)
            got = np.asarray(rho.todense())
            assert np.max(np.abs(got - want)) < 1e-12
            assert np.max(np.abs(rho.diagonal() - np.diag(want))) < 1e-12
            assert abs(rho[3, 5] - want[3, 5]) < 1e-12
            # the live view, read through the (non-identity) layout
            assert np.max(np.abs(qvm._density.diagonal().real - np.diag(want).real)) < 1e-12
            # sampling from diag(rho): numpy's choice() stream (rank 0's uniforms)
            np.random.seed(17)
            shots = qvm.run_and_measure(prog, trials=200)
            u = np.random.RandomState(17).random_sample(200)
            ref = orc.inverse_cdf_sample(np.diag(want).real, u)
            assert [int("".join(map(str, s)), 2) for s in shots] == [int(i) for i in ref]
            # MEASURE on rho (noise-free QVM): p0 and the collapse follow the oracle with the same uniform
            plain = QVMConnection(type_trans="density")
            mprog = spec_program(spec)
            mprog.measure(2, [0]).measure(4, [1])
            np.random.seed(5)
            plain.density(mprog)
            rs = np.random.RandomState(5)
            vec = density_oracle(spec_program(spec), n)
            outcomes = []
            for q in (2, 4):
                p0 = orc.fast_density_prob_zero(vec, q, n)
                o = 0 if rs.random_sample() < p0 else 1
                outcomes.append(o)
                vec = orc.fast_density_collapse(vec, q, o, p0 if o == 0 else 1 - p0, n)
            assert [int(b) for b in plain.classical_memory[[0, 1]]] == outcomes
            assert np.max(np.abs(np.asarray(plain._density.todense()) - orc.vec_to_rho(vec, n))) < 1e-12
            # apply_kraus leaves the QVM's rho alone and returns the channel's result
            before = np.asarray(plain._density.todense())
            out = plain.apply_kraus(G.bit_flip(0.2), 1)
            assert np.max(np.abs(np.asarray(out.todense()) - orc.vec_to_rho(
                orc.fast_apply_kraus(orc.rho_to_vec(before), G.bit_flip(0.2), 1, n), n))) < 1e-12
            assert np.max(np.abs(np.asarray(plain._density.todense()) - before)) == 0
        finally:
            _sharded.disable()

    run_ranks(world, rank_main)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_expectation_host_logic(world):
    """Pauli expectation values on sharded engines (numpy backend, thread ranks): X / Y factors on global qubits
    trigger an exchange, Z factors on global qubits are signs of the shard; density: gather through the layout."""
    from _programs import spec_program
    from pyquil.paulis import PauliSum, PauliTerm
    from referenceqvm.api import QVMConnection

    def rank_main(comm):
        _sharded.enable(device=0, comm=comm, backend_factory=NumpyShard,
                        engine_options={"tile_bits": 5, "low_bits": 2, "exchange": "push"})
        try:
            n = 9
            spec = orc.random_circuit_spec(n, 4, 3)
            prog = spec_program(spec)
            psi, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
            ops = [PauliTerm("Z", n - 1) * PauliTerm("Z", 0), PauliTerm("X", n - 1) * PauliTerm("Y", n - 2, 0.5),
                   PauliSum([PauliTerm("Y", 3, 2.0), PauliTerm("X", n - 1) * PauliTerm("Z", n - 2) * PauliTerm("X", 1)]),
                   spec_program([("X", [], [n - 1]), ("Z", [], [2])]), spec_program([])]
            terms = [[(1.0, {n - 1: "Z", 0: "Z"})], [(0.5, {n - 1: "X", n - 2: "Y"})],
                     [(2.0, {3: "Y"}), (1.0, {n - 1: "X", n - 2: "Z", 1: "X"})], [(1.0, {n - 1: "X", 2: "Z"})], [(1.0, {})]]
            got = QVMConnection().expectation(prog, ops)
            for g, t in zip(got, terms):
                assert abs(g - orc.expectation_state(psi, t, n)) < 1e-12
            nd = 5
            dspec = orc.random_circuit_spec(nd, 3, 6)
            rho = orc.vec_to_rho(density_oracle(spec_program(dspec), nd), nd)
            dops = [PauliTerm("Z", nd - 1) * PauliTerm("X", 0), PauliTerm("Y", nd - 1) * PauliTerm("Y", nd - 2, 0.25),
                    spec_program([("X", [], [nd - 1])])]
            dterms = [[(1.0, {nd - 1: "Z", 0: "X"})], [(0.25, {nd - 1: "Y", nd - 2: "Y"})], [(1.0, {nd - 1: "X"})]]
            got = QVMConnection(type_trans="density").expectation(spec_program(dspec), dops)
            for g, t in zip(got, dterms):
                assert abs(g - orc.expectation_density(rho, t, nd)) < 1e-12
        finally:
            _sharded.disable()

    run_ranks(world, rank_main)
