"""The sharded CUDA path on ONE GPU: the real ``ShardedEngine`` + ``CudaShard`` (libqvm_b200 kernels, push /
in-place exchange kernels over peer pointers, fused pass + exchange, all-gathered MEASURE runs, sharded
sampling, the public API) with G = 2 / 4 / 8 ranks running as threads of this process on device 0
(tests/_thread_comm.py).  Every shard is its own ``qvm_t`` on its own stream; "peer memory" is the
other shards' device pointers.  Amplitudes, MEASURE outcomes and sampled indices are diffed against
the oracle.  (tests/test_gpu_sharded.py runs the same checks with one process per GPU over NCCL / CUDA
IPC when the box has several GPUs.)"""
import numpy as np
import pytest

from _programs import spec_program
from _thread_comm import run_ranks
from oracle import qvm_oracle as orc
from referenceqvm import _sharded, gates as G
from referenceqvm.api import QVMConnection

pytestmark = pytest.mark.gpu


def oracle_ops(spec):
    out = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        out.append(("gate", np.asarray(m(*params) if params else m, dtype=complex), qubits))
    return out


def queue_spec(eng, spec):
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        eng.queue_matrix(m(*params) if params else m, qubits)


@pytest.mark.parametrize("exchange", ["push", "p2p"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_engine_on_one_gpu(world, exchange):
    facts = {}

    def rank_main(comm):
        rank = comm.rank
        ctx = _sharded.enable(device=0, comm=comm)
        try:
            for case, (n, depth, seed, jit) in enumerate(((14, 6, 1, 1), (16, 8, 2, 2), (20, 5, 1234, 2), (17, 10, 5, 1))):
                spec = orc.random_circuit_spec(n, depth, seed)
                eng = _sharded.ShardedEngine(n, ctx, exchange=exchange)
                assert eng.exchange == exchange, eng.exchange_note
                eng.shard.state.set_jit(jit)
                eng.reset()
                queue_spec(eng, spec)
                want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
                eng.flush()
                assert eng.swaps > 0
                assert abs(eng.norm2() - 1.0) < 1e-12
                probe_idx = [0, 1, 5, (1 << n) - 1, 1 << (n - 1), 12345]
                assert np.max(np.abs(eng.probe(probe_idx) - want[probe_idx])) < 1e-12
                got = eng.amplitudes()
                err = float(np.max(np.abs(got - want)))
                assert err < 1e-12, (n, depth, seed, err)
                st = eng.stats()
                # (in-process ranks ask for the same structure at the same moment: one of them compiles it, the others
                # run that pass interpreted -- so the specialised launches are counted over all ranks)
                jit_all = int(comm.allreduce([float(st["jit_launches"])])[0])
                if rank == 0:
                    facts[case] = (eng.swaps, eng.fused_swaps, st["fused_push_launches"], st["push_launches"], jit_all)
                psi = want
                # (uniforms off the round values: a qubit left in an exact |+> has p0 = 0.5 +- 1 ulp, and r = 0.5 would
                # make the outcome a matter of summation order)
                for q, r in ((n - 1, 0.3), (0, 0.8), (n // 2, 0.4375)):
                    outcome, p0 = eng.measure(q, r if rank == 0 else -1.0)
                    o_want, p_want, psi = orc.fast_measure(psi, q, n, r)
                    assert outcome == o_want and abs(p0 - p_want) < 1e-12, (case, rank, q, r, outcome, o_want, p0, p_want)
                err = float(np.max(np.abs(eng.amplitudes() - psi)))
                assert err < 1e-12, ("after measure", err)
                eng.close()

            # a run of MEASUREs on the sharded state (measure_hist per shard, all-gathered): all qubits, shuffled
            for n, depth, seed in ((18, 5, 7), (16, 6, 9)):
                spec = orc.random_circuit_spec(n, depth, seed)
                eng = _sharded.ShardedEngine(n, ctx, exchange=exchange)
                eng.reset()
                queue_spec(eng, spec)
                psi, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
                rs = np.random.RandomState(seed)
                order = [int(q) for q in rs.permutation(n)]
                uniforms = rs.random_sample(n)
                got = eng.measure_many(order, uniforms if rank == 0 else np.zeros(n))
                for (outcome, p0), q, r in zip(got, order, uniforms):
                    o_want, p_want, psi = orc.fast_measure(psi, q, n, r)
                    assert outcome == o_want and abs(p0 - p_want) < 1e-12
                err = float(np.max(np.abs(eng.amplitudes() - psi)))
                assert err < 1e-12, ("after measure_many", err)
                eng.close()

            # sampling through the layout, and the public API (SPMD: QVMConnection on every rank)
            n = 15
            spec = orc.random_circuit_spec(n, 6, 7)
            want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
            eng = _sharded.ShardedEngine(n, ctx, exchange=exchange)
            eng.reset()
            queue_spec(eng, spec)
            u = np.random.RandomState(11).random_sample(20000)
            idx, total = eng.sample(u if rank == 0 else np.zeros(u.size))
            assert abs(total - 1.0) < 1e-12
            logical = np.arange(1 << n, dtype=np.int64)
            phys = np.zeros_like(logical)
            for q, p in enumerate(eng.pos_of):
                phys |= ((logical >> q) & 1) << p
            probs_hbm = np.empty(1 << n)
            probs_hbm[phys] = np.abs(want) ** 2
            logical_of_phys = np.empty_like(logical)
            logical_of_phys[phys] = logical
            ref = logical_of_phys[np.asarray(orc.inverse_cdf_sample(probs_hbm, u), dtype=np.int64)]
            assert np.mean(idx.astype(np.int64) == ref) > 0.999
            eng.close()

            qvm = QVMConnection()
            wf, _ = qvm.wavefunction(spec_program(spec))
            assert isinstance(qvm._engine, _sharded.ShardedEngine)
            assert float(np.max(np.abs(wf.amplitudes - want))) < 1e-12
            shots = np.asarray(qvm.run_and_measure(spec_program(spec), list(range(n)), trials=4000))
            assert shots.shape == (4000, n)
            # marginals of the sampled bits against the exact ones (4000 shots: 5 sigma ~ 0.04)
            probs = np.abs(want) ** 2
            for q in range(n):
                p1 = float(probs[((np.arange(1 << n) >> q) & 1) == 1].sum())
                assert abs(shots[:, q].mean() - p1) < 0.05
            qvm._engine.close()
        finally:
            _sharded.disable()

    run_ranks(world, rank_main)
    if exchange == "push":
        # every epoch went through the push path, and some of them rode on a gate pass (fused stores)
        assert all(f[2] + f[3] >= f[0] for f in facts.values()), facts
        assert sum(f[1] for f in facts.values()) > 0 and sum(f[2] for f in facts.values()) > 0, facts
        assert facts[1][4] > 0 or facts[2][4] > 0, facts          # specialised kernels carried pushes too


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_density_qvm_on_one_gpu(world):
    """QVM_Density with rho sharded by rows over G in-process ranks (CUDA kernels, push exchange): density() under
    T2 dephasing on every qubit, diagonal / element reads through the layout, seeded sampling from diag(rho)
    (numpy's choice() stream), MEASURE on rho, apply_kraus -- against the oracle ("Density" row)."""
    from _programs import oracle_ops as prog_ops
    from referenceqvm.qvm_density import INFINITY, NoiseModel

    def oracle_vec(prog, n, p1q=None, p2q=None):
        vec = np.zeros(4 ** n, dtype=complex)
        vec[0] = 1.0
        for _, m, q in prog_ops(prog):
            vec = orc.fast_density_gate(vec, m, q, n)
            if p1q is not None:
                p = p1q if len(q) == 1 else p2q
                for i in range(n):
                    vec = orc.fast_apply_kraus(vec, G.dephasing(p), i, n)
        return vec

    def rank_main(comm):
        _sharded.enable(device=0, comm=comm)
        try:
            n = 7
            spec = [("H", [], [q]) for q in range(n)] + orc.random_circuit_spec(n, 3, 1234)
            prog = spec_program(spec)
            nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
            qvm = QVMConnection(type_trans="density", noise_model=nm)
            rho = qvm.density(prog)
            assert isinstance(qvm._engine, _sharded.ShardedEngine) and qvm._engine.swaps > 0
            p1, p2 = 1.0 - np.exp(-50e-9 / 30e-6), 1.0 - np.exp(-150e-9 / 30e-6)
            want = orc.vec_to_rho(oracle_vec(prog, n, p1, p2), n)
            assert np.max(np.abs(np.asarray(rho.todense()) - want)) < 1e-12
            assert np.max(np.abs(rho.diagonal() - np.diag(want))) < 1e-12
            assert abs(rho[3, 5] - want[3, 5]) < 1e-12
            assert abs(np.sum(qvm._density.diagonal()).real - 1.0) < 1e-12
            np.random.seed(17)
            shots = qvm.run_and_measure(prog, trials=500)
            ref = orc.inverse_cdf_sample(np.diag(want).real, np.random.RandomState(17).random_sample(500))
            assert [int("".join(map(str, s)), 2) for s in shots] == [int(i) for i in ref]
            plain = QVMConnection(type_trans="density")
            mprog = spec_program(spec)
            mprog.measure(2, [0]).measure(6, [1])
            np.random.seed(5)
            plain.density(mprog)
            rs = np.random.RandomState(5)
            vec = oracle_vec(spec_program(spec), n)
            outcomes = []
            for q in (2, 6):
                p0 = orc.fast_density_prob_zero(vec, q, n)
                o = 0 if rs.random_sample() < p0 else 1
                outcomes.append(o)
                vec = orc.fast_density_collapse(vec, q, o, p0 if o == 0 else 1 - p0, n)
            assert [int(b) for b in plain.classical_memory[[0, 1]]] == outcomes
            assert np.max(np.abs(np.asarray(plain._density.todense()) - orc.vec_to_rho(vec, n))) < 1e-12
            before = np.asarray(plain._density.todense())
            out = plain.apply_kraus(G.bit_flip(0.2), 1)
            assert np.max(np.abs(np.asarray(out.todense()) - orc.vec_to_rho(
                orc.fast_apply_kraus(orc.rho_to_vec(before), G.bit_flip(0.2), 1, n), n))) < 1e-12
            assert np.max(np.abs(np.asarray(plain._density.todense()) - before)) == 0
            qvm._engine.close()
            plain._engine.close()
        finally:
            _sharded.disable()

    run_ranks(world, rank_main)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_expectation_on_one_gpu(world):
    """Pauli expectation values on sharded states (CUDA kernels): X / Y factors on global qubits trigger an exchange,
    Z factors on global qubits are signs of the shard; density: gather through the layout."""
    from _programs import oracle_ops as prog_ops
    from pyquil.paulis import PauliSum, PauliTerm

    def rank_main(comm):
        _sharded.enable(device=0, comm=comm)
        try:
            n = 16
            spec = orc.random_circuit_spec(n, 5, 3)
            prog = spec_program(spec)
            psi, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
            ops = [PauliTerm("Z", n - 1) * PauliTerm("Z", 0), PauliTerm("X", n - 1) * PauliTerm("Y", n - 2, 0.5),
                   PauliSum([PauliTerm("Y", 3, 2.0), PauliTerm("X", n - 1) * PauliTerm("Z", n - 2) * PauliTerm("X", 1)]),
                   spec_program([("X", [], [n - 1]), ("Z", [], [2])]), spec_program([])]
            terms = [[(1.0, {n - 1: "Z", 0: "Z"})], [(0.5, {n - 1: "X", n - 2: "Y"})],
                     [(2.0, {3: "Y"}), (1.0, {n - 1: "X", n - 2: "Z", 1: "X"})], [(1.0, {n - 1: "X", 2: "Z"})], [(1.0, {})]]
            qvm = QVMConnection()
            got = qvm.expectation(prog, ops)
            for g, t in zip(got, terms):
                assert abs(g - orc.fast_expectation_state(psi, t, n)) < 1e-12
            qvm._engine.close()
            nd = 7
            dspec = orc.random_circuit_spec(nd, 3, 6)
            vec = np.zeros(4 ** nd, dtype=complex)
            vec[0] = 1.0
            for _, m, q in prog_ops(spec_program(dspec)):
                vec = orc.fast_density_gate(vec, m, q, nd)
            rho = orc.vec_to_rho(vec, nd)
            dops = [PauliTerm("Z", nd - 1) * PauliTerm("X", 0), PauliTerm("Y", nd - 1) * PauliTerm("Y", nd - 2, 0.25),
                    spec_program([("X", [], [nd - 1])])]
            dterms = [[(1.0, {nd - 1: "Z", 0: "X"})], [(0.25, {nd - 1: "Y", nd - 2: "Y"})], [(1.0, {nd - 1: "X"})]]
            dq = QVMConnection(type_trans="density")
            got = dq.expectation(spec_program(dspec), dops)
            for g, t in zip(got, dterms):
                assert abs(g - orc.expectation_density(rho, t, nd)) < 1e-12
            dq._engine.close()
        finally:
            _sharded.disable()

    run_ranks(world, rank_main)
