"""GPU parity, kernel level: every C-ABI compute entry point against the CPU oracle on the same
seeded inputs.  Tolerance: 1e-12 max-abs on amplitudes (float64), as BASELINE.json's north_star
states; integers (outcomes, sampled indices) must be identical."""
import itertools

import numpy as np
import pytest

from _model import random_state, random_unitary
from oracle import qvm_oracle as orc
from referenceqvm import _engine, _native as nat, _planner
from referenceqvm import gates as G

pytestmark = pytest.mark.gpu
TOL = 1e-12


def fresh(n, psi):
    st = nat.DeviceState(n)
    st.set_state(psi)
    return st


def check_apply(n, matrix, qubits, seed=0, tile=None):
    psi = random_state(n, seed)
    st = fresh(n, psi)
    if tile is not None:
        st.set_tile_bits(*tile)
    st.apply(matrix, qubits)
    got = st.get_state()
    st.close()
    want = orc.fast_apply_gate(psi, matrix, qubits, n)
    err = np.max(np.abs(got - want))
    assert err < TOL, (n, qubits, err)


def test_state_roundtrip_and_reset():
    for n in (0, 1, 5, 13, 20):
        psi = random_state(n, n)
        st = fresh(n, psi)
        assert np.array_equal(st.get_state(), psi)
        if n >= 5:
            assert np.array_equal(st.get_state(7, 9), psi[7:16])
            assert np.array_equal(st.get_strided(3, 5, 4), psi[3:3 + 20:5])
        assert abs(st.norm2() - 1.0) < 1e-12
        st.reset_zero(True)
        z = st.get_state()
        assert z[0] == 1.0 and not z[1:].any()
        st.close()


def test_one_qubit_gate_every_position():
    rs = np.random.RandomState(1)
    g = rs.standard_normal((2, 2)) + 1j * rs.standard_normal((2, 2))       # non-unitary on purpose
    for n in (1, 3, 9, 14):
        for q in range(n):
            check_apply(n, g, [q], seed=q)
    check_apply(16, G.H, [15])
    check_apply(16, G.H, [0])


def test_two_qubit_gate_all_orders():
    rs = np.random.RandomState(2)
    g = rs.standard_normal((4, 4)) + 1j * rs.standard_normal((4, 4))
    n = 14
    pairs = [(0, 1), (1, 0), (0, 13), (13, 0), (12, 13), (13, 12), (5, 9), (9, 5), (3, 12), (2, 6)]
    for a, b in pairs:
        check_apply(n, g, [a, b], seed=a + b)
    for a, b in itertools.permutations(range(4), 2):
        check_apply(4, g, [a, b])


def test_standard_gate_set_far_and_near():
    n = 14
    cases = [("CNOT", [], [0, 13]), ("CNOT", [], [13, 0]), ("CNOT", [], [6, 7]), ("CCNOT", [], [13, 0, 6]),
             ("CCNOT", [], [1, 2, 12]), ("CSWAP", [], [12, 0, 13]), ("CSWAP", [], [3, 13, 1]),
             ("SWAP", [], [0, 13]), ("ISWAP", [], [13, 2]), ("PSWAP", [0.4], [5, 11]), ("CZ", [], [13, 0]),
             ("CPHASE", [1.1], [12, 13]), ("CPHASE00", [1.1], [0, 13]), ("CPHASE01", [1.1], [13, 0]),
             ("CPHASE10", [1.1], [7, 1]), ("RZ", [0.3], [13]), ("RZ", [0.3], [0]), ("PHASE", [0.3], [12]),
             ("S", [], [13]), ("T", [], [0]), ("Z", [], [6]), ("X", [], [13]), ("Y", [], [0]), ("H", [], [12]),
             ("RX", [0.7], [13]), ("RY", [0.7], [1]), ("I", [], [4]), ("BARENCO", [0.3, 0.7, -1.1], [13, 0])]
    for name, params, qubits in cases:
        m = G.gate_matrix[name]
        check_apply(n, m(*params) if params else m, qubits, seed=len(name))


def test_dense_k3_to_k6():
    rs = np.random.RandomState(3)
    n = 14
    for k, qubits in [(3, [13, 0, 6]), (3, [2, 1, 0]), (4, [0, 13, 5, 9]), (5, [12, 3, 13, 0, 7]),
                      (6, [1, 13, 4, 10, 0, 8])]:
        g = rs.standard_normal((2 ** k,) * 2) + 1j * rs.standard_normal((2 ** k,) * 2)
        check_apply(n, g, qubits, seed=k)
    check_apply(3, random_unitary(8, rs), [1, 2, 0])
    check_apply(6, random_unitary(64, rs), [5, 0, 3, 1, 4, 2])


def test_tile_configurations():
    rs = np.random.RandomState(4)
    g2 = rs.standard_normal((4, 4)) + 1j * rs.standard_normal((4, 4))
    for tile in [(4, 0), (6, 3), (8, 4), (10, 5), (12, 5), (13, 5), (13, 4)]:
        check_apply(15, g2, [14, 9], tile=tile)
        check_apply(15, G.CNOT, [3, 14], tile=tile)
        check_apply(15, G.RZ(0.2), [14], tile=tile)


def test_errors_map_to_python_exceptions():
    st = nat.DeviceState(5)
    with pytest.raises(ValueError):
        st.apply(G.H, [5])
    with pytest.raises(ValueError):
        st.apply(G.CNOT, [2, 2])
    with pytest.raises(ValueError):
        st.apply(G.H, [-1])
    with pytest.raises(ValueError):
        st.measure(9, 0.5)
    st.close()


def compile_spec(spec):
    ops = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        ops.append(_engine.compile_gate(m(*params) if params else m, qubits))
    return ops


@pytest.mark.parametrize("n,depth,seed,tile,low", [(14, 8, 1, 12, 5), (16, 6, 2, 12, 5), (15, 8, 3, 10, 4),
                                                    (13, 10, 4, 8, 3), (18, 4, 5, 13, 5), (17, 6, 6, 12, 3),
                                                    (20, 4, 1234, 12, 5)])
def test_fused_groups_random_circuits(n, depth, seed, tile, low):
    for path in (1, 0):
        _fused_groups_random_circuits(n, depth, seed, tile, low, path)


def _fused_groups_random_circuits(n, depth, seed, tile, low, path):
    spec = orc.random_circuit_spec(n, depth, seed)
    psi = random_state(n, seed)
    want = psi
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, m(*params) if params else m, qubits, n)
    st = fresh(n, psi)
    st.set_kernel_path(path)
    groups = _planner.plan(compile_spec(spec), n, tile, low)
    for g in groups:
        st.apply_group(g.tile, g.ops)
    got = st.get_state()
    stats = st.stats()
    st.close()
    assert np.max(np.abs(got - want)) < TOL
    assert stats["hbm_passes"] == len(groups) and stats["gate_ops"] == len(spec)


def test_group_with_outer_controls_and_diagonals_on_non_tile_bits():
    n = 16
    rs = np.random.RandomState(7)
    mats = [(G.H, [3]), (G.CNOT, [15, 3]), (G.RZ(0.4), [14]), (G.CPHASE(0.9), [15, 2]), (G.CCNOT, [14, 13, 0]),
            (G.CPHASE01(0.3), [13, 1]), (G.T, [15]), (random_unitary(4, rs), [4, 2]), (G.CSWAP, [12, 0, 4]),
            (G.CZ, [15, 14])]
    ops = [_engine.compile_gate(m, q) for m, q in mats]
    psi = random_state(n, 3)
    st = fresh(n, psi)
    st.apply_group(list(range(8)), ops)            # one launch; bits 8..15 are outer
    got = st.get_state()
    st.close()
    want = psi
    for m, q in mats:
        want = orc.fast_apply_gate(want, m, q, n)
    assert np.max(np.abs(got - want)) < TOL


def test_sharded_positions_diagonals_and_controls():
    """Global positions (>= n_local) feed controls / diagonal factors from the shard index."""
    n_local, n_global = 10, 2
    n = n_local + n_global
    mats = [(G.H, [4]), (G.CNOT, [11, 4]), (G.RZ(0.7), [10]), (G.CPHASE(0.5), [11, 10]), (G.CZ, [10, 0]),
            (G.CCNOT, [11, 10, 9]), (G.CPHASE10(0.2), [11, 3]), (G.PHASE(1.3), [11])]
    ops = [_engine.compile_gate(m, q) for m, q in mats]
    psi = random_state(n, 9)
    want = psi
    for m, q in mats:
        want = orc.fast_apply_gate(want, m, q, n)
    for shard in range(4):
        st = nat.DeviceState(n_local)
        st.set_shard(n_global, shard)
        st.set_state(psi[shard << n_local:(shard + 1) << n_local])
        for g in _planner.plan(ops, n_local, 8, 4):
            st.apply_group(g.tile, g.ops)
        got = st.get_state()
        st.close()
        assert np.max(np.abs(got - want[shard << n_local:(shard + 1) << n_local])) < TOL, shard


def test_measure_kernel():
    for n, q in [(1, 0), (6, 0), (6, 5), (14, 2), (14, 3), (14, 13), (20, 11)]:
        psi = random_state(n, n + q)
        for r in (0.0, 0.3, 0.77, 0.999999):
            st = fresh(n, psi)
            outcome, p0 = st.measure(q, r)
            got = st.get_state()
            st.close()
            w_out, w_p0, want = orc.fast_measure(psi, q, n, r)
            assert outcome == w_out and abs(p0 - w_p0) < 1e-13
            assert np.max(np.abs(got - want)) < TOL
    # split form used by sharded states
    psi = random_state(12, 5)
    st = fresh(12, psi)
    p0 = st.prob_zero(7)
    assert abs(p0 - orc.fast_prob_zero(psi, 7, 12)) < 1e-13
    st.collapse(7, 1, 1.0 / np.sqrt(1 - p0))
    assert np.max(np.abs(st.get_state() - orc.fast_collapse(psi, 7, 1, 1 - p0, 12))) < TOL
    st.close()


def test_sampling_matches_inverse_cdf():
    rs = np.random.RandomState(11)
    for n in (1, 4, 11, 12, 17):
        psi = random_state(n, n)
        if n >= 11:
            psi[rs.randint(0, 2 ** n, size=2 ** n // 2)] = 0.0      # holes, whole zero blocks
            psi[: 2 ** (n - 2)] = 0.0
            psi /= np.linalg.norm(psi)
        st = fresh(n, psi)
        total = st.probs_prepare(0)
        u = rs.random_sample(5000)
        got = st.sample(u)
        st.close()
        probs = psi.real ** 2 + psi.imag ** 2
        want = orc.inverse_cdf_sample(probs, u)
        assert abs(total - 1.0) < 1e-12
        mism = np.nonzero(got != want)[0]
        # ties at cdf boundaries can move a draw to a neighbour when the summation order differs
        assert len(mism) <= 2, (n, len(mism))
        assert np.all(probs[got] > 0)


def test_sampling_chi_squared():
    n = 5
    psi = random_state(n, 2)
    st = fresh(n, psi)
    st.probs_prepare(0)
    shots = 200000
    got = st.sample(np.random.RandomState(0).random_sample(shots))
    st.close()
    probs = psi.real ** 2 + psi.imag ** 2
    counts = np.bincount(got.astype(np.int64), minlength=2 ** n)
    chi2 = np.sum((counts - shots * probs) ** 2 / (shots * probs))
    assert chi2 < 70.0        # 31 dof: P(chi2 > 70) ~ 7e-5


def test_density_diag_sampling_and_prob():
    n = 5
    rs = np.random.RandomState(4)
    a = rs.standard_normal((2 ** n, 2 ** n)) + 1j * rs.standard_normal((2 ** n, 2 ** n))
    rho = a @ a.conj().T
    rho /= np.trace(rho).real
    st = fresh(2 * n, rho.reshape(-1))
    for q in range(n):
        want = orc.fast_density_prob_zero(rho.reshape(-1), q, n)
        assert abs(st.density_prob_zero(n, q) - want) < 1e-13
    total = st.probs_prepare(1, n)
    u = rs.random_sample(3000)
    got = st.sample(u)
    st.close()
    want = orc.inverse_cdf_sample(np.real(np.diag(rho)), u)
    assert abs(total - 1.0) < 1e-12 and np.sum(got != want) <= 1


def test_dephase_all_matches_kraus_pairs():
    n = 4
    rs = np.random.RandomState(6)
    rho = rs.standard_normal((16, 16)) + 1j * rs.standard_normal((16, 16))
    p = 0.2
    vec = rho.reshape(-1)
    want = vec
    for q in range(n):
        want = orc.fast_apply_kraus(want, G.dephasing(p), q, n)
    st = fresh(2 * n, vec)
    st.dephase_all(n, (1 << n) - 1, 1 - p)
    got = st.get_state()
    st.close()
    assert np.max(np.abs(got - want)) < TOL


def test_pack_unpack_half_roundtrip():
    import torch
    n = 12
    psi = random_state(n, 1)
    st = fresh(n, psi)
    half = 2 ** (n - 1)
    buf = torch.empty(half * 2, dtype=torch.float64, device="cuda")
    for bit in (0, 5, 11):
        for val in (0, 1):
            st.pack_half(bit, val, 0, half, buf.data_ptr())
            st.sync()
            got = buf.cpu().numpy().view(np.complex128)
            idx = np.arange(2 ** n)
            assert np.array_equal(got, psi[((idx >> bit) & 1) == val])
    # swap the two halves of bit 5 through the buffer
    buf2 = torch.empty_like(buf)
    st.pack_half(5, 0, 0, half, buf.data_ptr())
    st.pack_half(5, 1, 0, half, buf2.data_ptr())
    st.unpack_half(5, 1, 0, half, buf.data_ptr())
    st.unpack_half(5, 0, 0, half, buf2.data_ptr())
    got = st.get_state()
    st.close()
    assert np.array_equal(got, orc.fast_apply_gate(psi, G.X, [5], n))


def test_regtile_mixed_ops_one_group():
    """Register-tiled kernel: dense1/X/H/dense2/diag with controls on register, thread and outer bits,
    targets on the lowest tile bits (forces middle rounds) and reversed dense2 argument orders."""
    n = 17
    rs = np.random.RandomState(21)
    u2a, u2b, u1 = random_unitary(4, rs), random_unitary(4, rs), random_unitary(2, rs)
    mats = [(G.H, [0]), (G.H, [1]), (G.CNOT, [0, 1]), (G.CNOT, [16, 2]), (G.CNOT, [11, 3]), (u1, [4]),
            (G.RZ(0.3), [0]), (G.RZ(0.4), [9]), (G.RZ(0.5), [15]), (G.CPHASE(0.6), [0, 9]),
            (G.CPHASE(0.7), [16, 1]), (G.CPHASE01(0.2), [3, 10]), (u2a, [11, 5]), (u2a, [5, 11]), (u2b, [0, 10]),
            (G.CCNOT, [0, 16, 6]), (G.CSWAP, [15, 7, 2]), (G.ISWAP, [8, 1]), (G.H, [11]), (G.X, [10]),
            (G.BARENCO(0.1, 0.2, 0.3), [13, 8]), (G.CZ, [2, 3]), (G.T, [4]), (G.PSWAP(0.9), [6, 7]), (G.H, [7]),
            (G.Y, [9]), (G.CPHASE10(0.8), [14, 0]), (G.PHASE(1.2), [16])]
    ops = [_engine.compile_gate(m, q) for m, q in mats]
    psi = random_state(n, 8)
    want = psi
    for m, q in mats:
        want = orc.fast_apply_gate(want, m, q, n)
    for tile in (list(range(12)), list(range(13)), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13], list(range(12)) + [15]):
        st = fresh(n, psi)
        st.apply_group(tile, ops)
        got = st.get_state()
        st.close()
        assert np.max(np.abs(got - want)) < TOL, tile


def test_exchange_half_between_two_shards_on_one_device():
    """qvm_exchange_half (the peer-memory swap kernel) with the "peer" being a second shard on the
    same GPU: after both sides ran their half of the range, the moving halves have traded places."""
    n = 14
    for bit in (0, 3, 7, 13):
        a, b = random_state(n, 1), random_state(n, 2)
        sa, sb = fresh(n, a), fresh(n, b)
        half = 1 << (n - 1)
        # rank "0" owns global bit 0: its local-bit-1 half moves; rank "1": its local-bit-0 half moves
        sa.exchange_half(bit, 1, sb.device_ptr(), 0, half // 2)
        sa.sync()
        sb.exchange_half(bit, 0, sa.device_ptr(), half // 2, half - half // 2)
        sb.sync()
        ga, gb = sa.get_state(), sb.get_state()
        sa.close()
        sb.close()
        idx = np.arange(1 << n)
        one = ((idx >> bit) & 1) == 1
        want_a, want_b = a.copy(), b.copy()
        want_a[one] = b[~one]
        want_b[~one] = a[one]
        assert np.array_equal(ga, want_a) and np.array_equal(gb, want_b), bit


def test_exchange_block_two_global_bits_on_one_device():
    """qvm_exchange_block with m = 2: four shards on one GPU play the four ranks of a subgroup.  After
    every rank traded sub-blocks with its three peers, global bits (g0, g1) and the victim local bits
    have swapped places in the full state."""
    n_local, m = 12, 2
    for vpos in ([3, 9], [0, 11], [5, 6]):
        full = random_state(n_local + m, 5)
        shards = [fresh(n_local, full[r << n_local:(r + 1) << n_local]) for r in range(4)]
        block = 1 << (n_local - m)
        for r in range(4):
            for p in range(4):
                if p == r:
                    continue
                lo, cnt = (0, block // 2) if r < p else (block // 2, block - block // 2)
                shards[r].exchange_block(vpos, p, r, shards[p].device_ptr(), lo, cnt)
            shards[r].sync()
        got = np.concatenate([s.get_state() for s in shards])
        for s in shards:
            s.close()
        idx = np.arange(1 << (n_local + m))
        # new index: global bit i <-> local bit vpos[i]
        src = idx.copy()
        for i, lp in enumerate(vpos):
            gp = n_local + i
            lb, gb = (idx >> lp) & 1, (idx >> gp) & 1
            src = src & ~((1 << lp) | (1 << gp)) | (gb << lp) | (lb << gp)
        assert np.array_equal(got, full[src]), vpos


def _run_circuit_state(n, spec, jit, repeats=1):
    eng = _engine.Engine(n)
    eng.state.set_jit(jit)
    out = None
    for _ in range(repeats):
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        out = eng.amplitudes()
    stats = eng.stats()
    eng.close()
    return out, stats


@pytest.mark.parametrize("n,depth,seed", [(14, 10, 1), (17, 8, 2), (20, 6, 1234)])
def test_specialised_kernels_match_oracle_and_interpreter(n, depth, seed):
    """JIT path (policy 2: compile every group at first sight): same amplitudes as the interpreted kernel
    bit for bit (same device functions, same order) and as the oracle within 1e-12."""
    spec = orc.random_circuit_spec(n, depth, seed)
    want = random_state(n, 0) * 0
    want[0] = 1.0
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    interp, s0 = _run_circuit_state(n, spec, jit=0)
    jitted, s2 = _run_circuit_state(n, spec, jit=2)
    assert s0["jit_launches"] == 0 and s2["jit_launches"] > 0
    assert np.max(np.abs(jitted - want)) < TOL
    assert np.array_equal(jitted, interp)


def test_specialised_kernels_background_policy_and_new_angles():
    """Default policy: first run interpreted, second run submits the structures, after qvm_jit_wait the
    third run uses the compiled kernels; a circuit with the same structure but other angles reuses them."""
    n = 16
    spec = orc.random_circuit_spec(n, 6, 9)
    eng = _engine.Engine(n)
    eng.state.set_jit(1)

    def run(sp):
        eng.reset()
        for name, params, qubits in sp:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        return eng.amplitudes()

    hits = nat.jit_counters()["disk_hits"]
    a1 = run(spec)
    # (cubins an earlier run of this suite left in the disk cache are loaded at first sight)
    assert eng.stats()["jit_launches"] == 0 or nat.jit_counters()["disk_hits"] > hits
    a2 = run(spec)
    nat.jit_wait()
    before = eng.stats()["jit_launches"]
    a3 = run(spec)
    assert eng.stats()["jit_launches"] > before
    assert np.array_equal(a1, a2) and np.array_equal(a1, a3)
    # same gates, other rotation angles: the structure (and therefore the kernels) is the same
    other = [(name, [p + 0.37 for p in params], qubits) for name, params, qubits in spec]
    compiles = eng.stats()["jit_compiles"]
    mid = eng.stats()["jit_launches"]
    got = run(other)
    assert eng.stats()["jit_launches"] > mid and eng.stats()["jit_compiles"] == compiles
    want = np.zeros(1 << n, dtype=complex)
    want[0] = 1.0
    for name, params, qubits in other:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    assert np.max(np.abs(got - want)) < TOL
    eng.close()


# ---- in-tile qubit relabelling --------------------------------------------------------------------------
def _logical_view(psi_phys, pos_of):
    n = len(pos_of)
    logical = np.arange(1 << n, dtype=np.uint64)
    phys = np.zeros_like(logical)
    for q, p in enumerate(pos_of):
        phys |= ((logical >> np.uint64(q)) & np.uint64(1)) << np.uint64(p)
    return psi_phys[phys]


@pytest.mark.parametrize("jit", [0, 2])
def test_relabelling_pass_on_device(jit):
    """qvm_apply_group_relabel: ops are applied and the requested qubits leave the pass on the low
    positions; the reported layout explains the buffer exactly (interpreted and specialised kernels)."""
    n, tile = 15, [0, 1, 2, 3, 5, 6, 8, 9, 10, 12, 14]
    rs = np.random.RandomState(3)
    psi0 = random_state(n, 17)
    spec = []
    for _ in range(30):
        q = int(rs.choice(tile))
        c = rs.randint(4)
        if c == 0:
            spec.append((G.H, [q]))
        elif c == 1:
            spec.append((G.RZ(rs.uniform(0, 6)), [int(rs.randint(n))]))
        elif c == 2:
            ctl = int(rs.randint(n))
            if ctl != q:
                spec.append((G.CNOT, [ctl, q]))
        else:
            a, b = [int(x) for x in rs.permutation(n)[:2]]
            spec.append((G.CPHASE(rs.uniform(0, 6)), [a, b]))
    ops = []
    for m, q in spec:
        ops += list(_engine.expand_op(_engine.compile_gate(m, q)))
    want = psi0
    for m, q in spec:
        want = orc.fast_apply_gate(want, np.asarray(m, dtype=complex), q, n)
    for bring, group in (([9, 14, 5], ops), ([12, 1, 6, 10], ops), ([8], []), ([], ops)):
        st = nat.DeviceState(n)
        st.set_jit(jit)
        st.set_state(psi0)
        new_pos = st.apply_group(tile, group, bring_low=bring, n_low=4)
        pos_of = list(range(n))
        for p, np_ in zip(tile, new_pos):
            pos_of[p] = np_
        assert sorted(pos_of) == list(range(n)) and all(pos_of[b] < 4 for b in bring)
        got = st.get_state()
        expect = want if group else psi0
        assert np.max(np.abs(_logical_view(got, pos_of) - expect)) < TOL
        # the layout-aware readback returns logical order directly, whole and strided
        assert np.array_equal(st.get_layout(0, 1, 1 << n, pos_of), _logical_view(got, pos_of))
        assert np.array_equal(st.get_layout(5, 7, 1000, pos_of), _logical_view(got, pos_of)[5:5 + 7000:7])
        if jit == 2 and (group or bring):
            assert st.stats()["jit_launches"] == 1
        st.close()


@pytest.mark.parametrize("n,depth,seed", [(14, 10, 4), (18, 8, 5), (21, 5, 1234)])
def test_engine_with_and_without_relabelling(n, depth, seed):
    """The relabelling planner saves passes and changes nothing else: amplitudes (read through the
    layout), MEASURE outcomes and the restored buffer all agree with the fixed-layout engine and the oracle."""
    spec = orc.random_circuit_spec(n, depth, seed)
    want = np.zeros(1 << n, dtype=complex)
    want[0] = 1.0
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    out = {}
    for relabel in (False, True):
        eng = _engine.Engine(n)
        eng.relabel = relabel
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        amps = eng.amplitudes()
        assert np.max(np.abs(amps - want)) < TOL
        assert eng.layout_is_identity() != relabel or n <= eng.tile_bits
        p0 = [eng.prob_zero(q) for q in range(n)]
        assert np.max(np.abs(np.array(p0) - [orc.fast_prob_zero(want, q, n) for q in range(n)])) < 1e-12
        idx, _ = eng.sample(np.random.RandomState(1).random_sample(2000), physical=True)
        bits = np.stack([(idx >> np.uint64(eng.position(q))) & np.uint64(1) for q in range(n)], axis=1)
        logical_idx, _ = eng.sample(np.random.RandomState(1).random_sample(2000))
        assert np.array_equal((bits << np.arange(n, dtype=np.uint64)).sum(axis=1).astype(np.uint64), logical_idx)
        assert np.all(np.abs(want[logical_idx.astype(np.int64)]) > 0)
        passes = eng.groups_launched
        eng.restore_layout()
        assert eng.layout_is_identity()
        assert np.max(np.abs(eng.state.get_state() - want)) < TOL
        outcome, p = eng.measure(n - 1, 0.3)
        out[relabel] = (passes, outcome, p)
        eng.close()
    assert out[True][1:] == pytest.approx(out[False][1:], abs=1e-12)
    assert out[True][0] <= out[False][0]


# ---- runs of MEASUREs ---------------------------------------------------------------------------------------
def test_measure_hist_kernel_against_numpy():
    """qvm_measure_hist: joint distribution over up to 5 high positions x the 5 lowest index bits, with and
    without the fused collapse of the previous chunk; deterministic."""
    n = 14
    psi = random_state(n, 31)
    st = nat.DeviceState(n)
    st.set_state(psi)
    idx = np.arange(1 << n, dtype=np.int64)

    def model(vec, high):
        bins = np.zeros_like(idx)
        for i, p in enumerate(high):
            bins |= ((idx >> p) & 1) << i
        hist = np.zeros((1 << len(high), 32))
        np.add.at(hist, (bins, idx & 31), np.abs(vec) ** 2)
        return hist

    for high in ([], [5], [13, 6, 9], [7, 8, 9, 10, 11]):
        got = st.measure_hist(high)
        assert got.shape == (1 << len(high), 32)
        assert np.max(np.abs(got - model(psi, high))) < 1e-14
        assert np.array_equal(got, st.measure_hist(high))                 # fixed summation order
    mask, value, scale = (1 << 3) | (1 << 9) | (1 << 12), (1 << 9), 1.7
    got = st.measure_hist([6, 12], (mask, value, scale))
    want_vec = np.where((idx & mask) == value, psi * scale, 0.0)
    assert np.max(np.abs(got - model(want_vec, [6, 12]))) < 1e-13
    assert np.max(np.abs(st.get_state() - want_vec)) < 1e-15
    st.measure_hist([], (1 << 6, 0, 0.5), want_hist=False)
    want_vec = np.where((idx & (1 << 6)) == 0, want_vec * 0.5, 0.0)
    assert np.max(np.abs(st.get_state() - want_vec)) < 1e-15
    with pytest.raises(ValueError):
        st.measure_hist([3])
    st.close()


def test_measure_hist_filter_mode_and_single_survivor():
    """do_collapse = 2: the histogram of the amplitudes consistent with (mask, value), times scale^2, state left
    alone -- for masks that fix none, some or all of the five lowest bits; then the write-only collapse of a run
    that measured every position."""
    n = 13
    psi = random_state(n, 77)
    st = nat.DeviceState(n)
    st.set_state(psi)
    idx = np.arange(1 << n, dtype=np.int64)

    def model(vec, high):
        bins = np.zeros_like(idx)
        for i, p in enumerate(high):
            bins |= ((idx >> p) & 1) << i
        hist = np.zeros((1 << len(high), 32))
        np.add.at(hist, (bins, idx & 31), np.abs(vec) ** 2)
        return hist

    rs = np.random.RandomState(5)
    cases = [(0, 0), ((1 << 0) | (1 << 3), 1 << 3), (31, 21), ((1 << 9) | (1 << 12), 1 << 12), (0b1010101010101, 0b1000100000001),
             ((1 << 10) - 1, 0b1100110011), ((1 << n) - 1 - (1 << 6) - (1 << 11), 0b1001100110101 & ~((1 << 6) | (1 << 11)))]
    for mask, value in cases:
        free = [p for p in range(5, n) if not (mask >> p) & 1]
        high = [int(p) for p in rs.permutation(free)[:min(5, len(free))]]
        scale = float(rs.uniform(0.5, 3.0))
        got = st.measure_hist(high, (mask, value, scale), apply=False)
        want = model(np.where((idx & mask) == value, psi * scale, 0.0), high)
        assert np.max(np.abs(got - want)) < 1e-13, (mask, value, high)
        assert np.array_equal(got, st.measure_hist(high, (mask, value, scale), apply=False))     # fixed summation order
        assert np.array_equal(st.get_state(), psi)                                                # nothing written
    assert np.count_nonzero(st.measure_hist([7], (3, 1, 0.0), apply=False)) == 0
    x0 = 0b1011001110101
    st.measure_hist([], ((1 << n) - 1, x0, 2.5), want_hist=False)
    want = np.zeros(1 << n, dtype=np.complex128)
    want[x0] = psi[x0] * 2.5
    assert np.array_equal(st.get_state(), want)
    with pytest.raises(ValueError):
        st.measure_hist([7], (1 << n, 0, 1.0), apply=False)           # mask outside the shard
    st.close()


@pytest.mark.parametrize("n,seed", [(12, 3), (16, 4), (21, 5)])
def test_measure_many_on_device_matches_sequential_measure(n, seed):
    """Engine.measure_many == the same MEASUREs one by one (same uniforms): outcomes, p0 and the collapsed
    state, on a relabelled layout."""
    spec = orc.random_circuit_spec(n, 5, seed)
    rs = np.random.RandomState(seed)
    order = [int(q) for q in rs.permutation(n)]
    uniforms = rs.random_sample(n)
    out = {}
    for fused in (True, False):
        eng = _engine.Engine(n)
        eng.reset()
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            eng.queue_matrix(m(*params) if params else m, qubits)
        if fused:
            res = eng.measure_many(order, uniforms)
        else:
            res = [eng.measure(q, r) for q, r in zip(order, uniforms)]
        passes = eng.stats()["hbm_passes"]
        out[fused] = (res, eng.amplitudes(), passes)
        eng.close()
    assert [o for o, _ in out[True][0]] == [o for o, _ in out[False][0]]
    assert np.max(np.abs(np.array([p for _, p in out[True][0]]) - np.array([p for _, p in out[False][0]]))) < 1e-12
    assert np.max(np.abs(out[True][1] - out[False][1])) < 1e-12
    index = sum(o << q for (o, _), q in zip(out[True][0], order))
    assert abs(abs(out[True][1][index]) - 1.0) < 1e-12


def test_relabel_and_layout_argument_checks():
    """Bad requests are refused with QVM_ERR_INVALID (ValueError), and the state is left alone."""
    n = 12
    psi = random_state(n, 2)
    st = nat.DeviceState(n)
    st.set_state(psi)
    tile = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    with pytest.raises(ValueError):
        st.apply_group(tile, [], bring_low=[11], n_low=4)             # position outside the tile
    with pytest.raises(ValueError):
        st.get_layout(0, 1, 1 << n, [0] * n)                           # not a permutation
    with pytest.raises(ValueError):
        st.get_layout(0, 1, 1 << n, list(range(n - 1)))                # wrong length
    with pytest.raises(ValueError):
        st.get_layout(5, 3, 1 << n, list(range(n)))                    # range runs off the shard
    assert np.array_equal(st.get_state(), psi)
    # a request for more qubits than there are low positions: the first ones win, the report is truthful
    new_pos = st.apply_group(tile, [], bring_low=[5, 6, 7, 8, 9, 10], n_low=4)
    pos_of = list(range(n))
    for p, q in zip(tile, new_pos):
        pos_of[p] = q
    assert sorted(pos_of) == list(range(n)) and sum(1 for p in (5, 6, 7, 8, 9, 10) if pos_of[p] < 4) == 4
    assert np.array_equal(st.get_layout(0, 1, 1 << n, pos_of), psi)
    st.close()


# ---- handle parking (qvm_destroy / qvm_create / qvm_trim) ------------------------------------------------------
def test_parked_handle_resources_are_reused_cleanly():
    """A destroyed handle's resources are taken over by the next handle of the same shape: the new handle starts
    with zeroed counters, no sticky state, and computes the same as a fresh one; qvm_trim empties the lot."""
    n = 13
    spec = orc.random_circuit_spec(n, 6, 3)
    want, fresh_stats = _run_circuit_state(n, spec, jit=0)
    nat.trim()
    first = nat.DeviceState(n)
    first.set_state(random_state(n, 9))
    first.measure(3, 0.4)                      # leaves sampling / scratch state behind
    ptr = first.device_ptr()
    first.close()
    again = nat.DeviceState(n)                 # same shape, same device: the parked set
    assert again.device_ptr() == ptr
    st = again.stats()
    assert st["kernel_launches"] == 0 and st["hbm_passes"] == 0 and st["h2d_bytes"] == 0
    again.close()
    got, stats = _run_circuit_state(n, spec, jit=0)
    assert np.array_equal(got, want) and stats["gate_ops"] == fresh_stats["gate_ops"]
    other = nat.DeviceState(n + 1)             # another shape never takes it
    assert other.device_ptr() != ptr
    other.close()
    nat.trim()
    got, _ = _run_circuit_state(n, spec, jit=0)
    assert np.array_equal(got, want)
