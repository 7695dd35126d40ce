"""Worker of tests/test_gpu_sharded.py: one rank (= one GPU) of an NCCL group running the product
sharded path (CudaShard: libqvm_b200 kernels + NCCL send/recv), diffed against the oracle."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, os.path.join(REPO, "reference-qvm_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import referenceqvm  # noqa: E402,F401
from _programs import spec_program  # noqa: E402
from oracle import qvm_oracle as orc  # noqa: E402
from referenceqvm import _sharded, gates as G  # noqa: E402
from referenceqvm.api import QVMConnection  # noqa: E402


def oracle_ops(spec):
    out = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        out.append(("gate", np.asarray(m(*params) if params else m, dtype=complex), qubits))
    return out


def main():
    rank, world = int(os.environ["QVM_TEST_RANK"]), int(os.environ["QVM_TEST_WORLD"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ["QVM_TEST_PORT"], rank=rank,
                            world_size=world, device_id=torch.device("cuda", rank))
    ctx = _sharded.enable(device=rank)
    for n, depth, seed, chunk in ((12, 6, 1, 64), (16, 8, 2, 1 << 20), (20, 5, 1234, 1 << 12), (14, 10, 5, 1 << 9)):
        spec = orc.random_circuit_spec(n, depth, seed)
        eng = _sharded.ShardedEngine(n, ctx, chunk_amps=chunk)
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
        eng.flush()
        assert eng.swaps > 0
        assert abs(eng.norm2() - 1.0) < 1e-12
        got = eng.amplitudes()
        err = float(np.max(np.abs(got - want)))
        assert err < 1e-12, (n, depth, seed, err)
        psi = want
        for q, r in ((n - 1, 0.3), (0, 0.8), (n // 2, 0.4375)):      # (off 0.5: an exact |+> has p0 = 0.5 +- 1 ulp)
            outcome, p0 = eng.measure(q, r if rank == 0 else -1.0)
            o_want, p_want, psi = orc.fast_measure(psi, q, n, r)
            assert outcome == o_want and abs(p0 - p_want) < 1e-12
        err = float(np.max(np.abs(eng.amplitudes() - psi)))
        assert err < 1e-12, ("after measure", err)
        eng.close()

    # a run of MEASUREs on the sharded state (measure_hist per shard, all-gathered): all qubits, shuffled
    for n, depth, seed in ((18, 5, 7), (16, 6, 9)):
        spec = orc.random_circuit_spec(n, depth, seed)
        eng = _sharded.ShardedEngine(n, ctx)
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
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

    # the public API, sharded: QVMConnection on every rank (SPMD)
    n = 15
    spec = orc.random_circuit_spec(n, 6, 7)
    qvm = QVMConnection()
    qvm.device = rank
    wf, _ = qvm.wavefunction(spec_program(spec))
    want, _ = orc.fast_run_wavefunction(oracle_ops(spec), n)
    assert isinstance(qvm._engine, _sharded.ShardedEngine)
    assert float(np.max(np.abs(wf.amplitudes - want))) < 1e-12
    np.random.seed(11)
    shots = np.asarray(qvm.run_and_measure(spec_program(spec), list(range(n)), trials=20000))
    idx = (shots << np.arange(n)).sum(axis=1)
    u = np.random.RandomState(11).random_sample(20000)
    # same uniforms (rank 0's); the CDF is walked in HBM order, indices come back through the layout
    logical = np.arange(1 << n, dtype=np.int64)
    phys = np.zeros_like(logical)
    for q, p in enumerate(qvm._engine.pos_of):
        phys |= ((logical >> q) & 1) << p
    probs_hbm = np.empty(1 << n)
    probs_hbm[phys] = np.abs(want) ** 2
    logical_of_phys = np.empty_like(logical)
    logical_of_phys[phys] = logical
    ref = logical_of_phys[np.asarray(orc.inverse_cdf_sample(probs_hbm, u), dtype=np.int64)]
    assert np.mean(idx == ref) > 0.999
    # density matrices sharded by rows ("Density" row): small case against the oracle ...
    from referenceqvm.qvm_density import INFINITY, NoiseModel
    from _programs import oracle_ops as prog_ops

    def dephased_plus_state(nq, f):
        """rho after H(0) .. H(nq-1) with dephasing on every qubit after every instruction, closed form: qubit q is
        in superposition from instruction q on, so its coherences are damped (nq - q) times."""
        def entry(i, j):
            d = i ^ j
            return 2.0 ** -nq * float(np.prod([f ** (nq - q) for q in range(nq) if (d >> q) & 1] or [1.0]))
        return entry

    p1 = 1.0 - np.exp(-50e-9 / 30e-6)
    sup = sum(np.kron(k, np.conj(k)) for k in G.dephasing(p1))
    f = float(sup[1, 1].real)
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    for nq in ((6, 16) if os.environ.get("QVM_TEST_BIG") == "1" and world == 2 else (6,)):
        dq = QVMConnection(type_trans="density", noise_model=nm)
        dq.device = rank
        hprog = spec_program([("H", [], [q]) for q in range(nq)])
        rho = dq.density(hprog)
        assert isinstance(dq._engine, _sharded.ShardedEngine)
        entry = dephased_plus_state(nq, f)
        if nq <= 8:                        # the closed form itself, against the oracle's Kraus arithmetic
            vec = np.zeros(4 ** nq, dtype=complex)
            vec[0] = 1.0
            for _, m, q in prog_ops(hprog):
                vec = orc.fast_density_gate(vec, m, q, nq)
                for i in range(nq):
                    vec = orc.fast_apply_kraus(vec, G.dephasing(p1), i, nq)
            want = orc.vec_to_rho(vec, nq)
            assert np.max(np.abs(np.asarray(rho.todense()) - want)) < 1e-12
            assert max(abs(want[i, j] - entry(i, j)) for i in range(0, 64, 7) for j in range(0, 64, 5)) < 1e-14
        rs = np.random.RandomState(nq)
        pairs = [(int(rs.randint(1 << nq)), int(rs.randint(1 << nq))) for _ in range(48)] + [(0, 0), ((1 << nq) - 1, 0)]
        worst = max(abs(rho[i, j] - entry(i, j)) for i, j in pairs)
        assert worst < 1e-13, (nq, worst)
        assert abs(np.sum(rho.diagonal()).real - 1.0) < 1e-12
        print("rank %d: %d-qubit rho (%.0f GiB over %d GPUs) matches the closed form, worst %.1e"
              % (rank, nq, 16.0 * 4 ** nq / 2 ** 30, world, worst))
        del rho
        dq._density = None
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d/%d ok" % (rank, world))


if __name__ == "__main__":
    main()
