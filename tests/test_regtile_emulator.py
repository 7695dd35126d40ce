"""The register-tiled kernel's encoder and round body, executed on the CPU (tests/native), against
the oracle: same source as the CUDA kernel (qvm_regtile.cuh is __host__ __device__), so opcode
selection, round planning, relabelling and the op handlers are all covered without a GPU."""
import numpy as np
import pytest

import _emu
from _model import apply_op, random_state, random_unitary
from _numpy_shard import localise, apply_local
from oracle import qvm_oracle as orc
from referenceqvm import _engine, _native as nat, _planner
from referenceqvm import gates as G

TOL = 1e-12


def compile_spec(spec):
    ops = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        ops.append(_engine.compile_gate(m(*params) if params else m, qubits))
    return ops


REG_BITS = 4      # overridden by the parametrised fixture below


@pytest.fixture(autouse=True, params=[4, 5])
def reg_bits(request):
    """Every test of this file runs for 16 and for 32 amplitudes per thread."""
    global REG_BITS
    REG_BITS = request.param
    yield request.param
    REG_BITS = 4


def run_emulated(psi, ops, n, tile_bits, low_bits, force_full=False):
    stats = {"groups": 0, "rounds": 0, "fallback": 0}
    for g in _planner.plan(ops, n, tile_bits, low_bits):
        new, info = _emu.apply_group(psi, n, g.tile, g.ops, force_full=force_full, reg_bits=REG_BITS)
        if info["status"] == 1:                 # outside the register-tile encoder: model the fallback
            stats["fallback"] += 1
            for op in g.ops:
                psi = apply_op(psi, op, n)
            continue
        assert info["status"] == 0
        psi = new
        stats["groups"] += 1
        stats["rounds"] += info["rounds"]
    return psi, stats


@pytest.mark.parametrize("n,depth,seed,tile", [(12, 12, 1, 12), (13, 10, 2, 12), (14, 8, 3, 12), (14, 6, 1234, 13),
                                               (11, 20, 5, 9), (15, 5, 7, 12), (13, 12, 11, 10)])
def test_random_circuit_through_emulated_kernel(n, depth, seed, tile):
    spec = orc.random_circuit_spec(n, depth, seed)
    ops = compile_spec(spec)
    psi0 = random_state(n, seed)
    want = psi0
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    for force_full in (False, True):
        got, stats = run_emulated(psi0, ops, n, tile, 5, force_full)
        assert stats["fallback"] == 0 and stats["groups"] >= 1
        err = np.max(np.abs(got - want))
        assert err < TOL, (n, depth, seed, force_full, err)


def test_qft_and_inverse():
    n = 13
    spec = orc.qft_spec(n)
    ops = compile_spec(spec)
    psi0 = random_state(n, 4)
    want = psi0
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    got, stats = run_emulated(psi0, ops, n, 12, 5)
    assert np.max(np.abs(got - want)) < TOL


def test_every_gate_kind_and_position():
    """Dense 1q/2q (incl. non-unitary), controlled, diagonal k = 1..3, scalars, projectors."""
    rs = np.random.RandomState(0)
    n = 12
    g1 = rs.standard_normal((2, 2)) + 1j * rs.standard_normal((2, 2))
    g2 = rs.standard_normal((4, 4)) + 1j * rs.standard_normal((4, 4))
    d3 = np.diag(np.exp(1j * rs.uniform(0, 6, 8)))
    d2 = np.diag(rs.standard_normal(4) + 1j * rs.standard_normal(4))
    cg1 = np.eye(4, dtype=complex)
    cg1[2:, 2:] = g1
    crz = np.diag([1, 1, np.exp(0.3j), 0.5 * np.exp(-1.1j)])                   # control + diag(a, b)
    ccrz = np.diag([1, 1, 1, 1, 1, 1, np.exp(0.7j), np.exp(-0.2j)])            # two controls + diag(a, b)
    mats = []
    for q in range(n):
        mats += [(g1, [q]), (G.H, [q]), (G.RZ(0.3 + q), [q]), (G.X, [q]), (G.P0 + 0.5 * G.P1, [q]), (G.P1, [(q + 3) % n])]
        mats[-1:] = []          # keep the state away from zero: drop the projector
        a, b, c = q, (q + 5) % n, (q + 7) % n
        mats += [(g2, [a, b]), (G.CNOT, [a, b]), (G.CNOT, [b, a]), (G.CPHASE(0.7), [a, c]), (cg1, [c, a]),
                 (d2, [b, c]), (G.CPHASE00(1.1), [a, b]), (G.CPHASE01(0.4), [c, b]), (G.SWAP, [a, c]),
                 (G.CCNOT, [a, b, c]), (d3, [c, a, b]), (G.ISWAP, [b, a]), (G.CZ, [a, b]), (np.diag([1, 1j]), [c]),
                 (2.0 * np.eye(2), [a]), (crz, [a, b]), (crz, [c, a]), (ccrz, [a, b, c]), (ccrz, [c, b, a])]
    ops = [_engine.compile_gate(m, q) for m, q in mats]
    psi0 = random_state(n, 9)
    want = psi0
    for m, q in mats:
        want = orc.fast_apply_gate(want, np.asarray(m, dtype=complex), q, n)
    scale = np.max(np.abs(want))
    for tile, low in ((12, 5), (10, 5), (9, 3)):
        got, stats = run_emulated(psi0, ops, n, tile, low)
        assert np.max(np.abs(got - want)) / scale < TOL, (tile, low)
    # the same zoo under the relabelling planner (rare handlers, wide gates and relabelling together)
    expanded = []
    for op in ops:
        expanded += list(_engine.expand_op(op))
    for tile, low in ((10, 4), (9, 3), (11, 5)):
        got, pos_of, passes, moved = run_emulated_relabelled(psi0, expanded, n, tile, low)
        assert np.max(np.abs(logical_view(got, pos_of) - want)) / scale < TOL, (tile, low)
        assert moved > 0


def test_projector_diagonal_with_zero_entry():
    """diag(0, 1) / diag(1, 0) on register slots must not be normalised by a zero d0."""
    n = 12
    mats = [(G.H, [q]) for q in range(n)] + [(G.P1, [11]), (G.H, [11]), (G.P0, [10]), (G.H, [10]), (G.P1, [0])]
    ops = [_engine.compile_gate(m, q) for m, q in mats]
    psi0 = random_state(n, 2)
    want = psi0
    for m, q in mats:
        want = orc.fast_apply_gate(want, np.asarray(m, dtype=complex), q, n)
    got, _ = run_emulated(psi0, ops, n, 12, 5)
    assert np.max(np.abs(got - want)) < TOL


def test_sharded_group_with_global_controls_and_diagonals():
    """Ops whose controls / diagonal targets sit on global positions, per shard index."""
    n_local, n_global = 11, 2
    n = n_local + n_global
    spec = orc.random_circuit_spec(n, 8, 21)
    ops = []
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        op = _engine.compile_gate(m(*params) if params else m, qubits)
        if op.kind == nat.OP_DENSE and any(t >= n_local for t in op.targets):
            continue                # dense targets on global positions need the exchange step
        ops.append(op)
    for shard in range(1 << n_global):
        psi0 = random_state(n_local, shard)
        want = psi0
        for op in ops:
            lop = localise(op, n_local, shard)
            if lop is not None:
                want = apply_local(want, lop, n_local)
        psi = psi0
        for g in _planner.plan(ops, n_local, 10, 5):
            psi, info = _emu.apply_group(psi, n_local, g.tile, g.ops, n_global=n_global, shard_index=shard,
                                         reg_bits=REG_BITS)
            assert info["status"] == 0
        assert np.max(np.abs(psi - want)) < TOL, shard


def test_expanded_two_target_diagonals():
    """_engine.expand_op: diag(d00..d11) -> two 1-qubit diagonals + a doubly controlled scalar."""
    rs = np.random.RandomState(3)
    n = 12
    mats = [(G.H, [q]) for q in range(n)]
    for a, b in ((0, 11), (11, 0), (5, 6), (3, 9), (10, 2)):
        d = np.diag(rs.standard_normal(4) + 1j * rs.standard_normal(4))
        mats += [(d, [a, b]), (G.H, [a]), (np.diag([1, 0.9, 0.9, 1]), [b, a]), (G.CPHASE00(0.4), [a, b])]
    ops = []
    for m, q in mats:
        ops.extend(_engine.expand_op(_engine.compile_gate(m, q)))
    assert len(ops) > len(mats)
    psi0 = random_state(n, 8)
    want = psi0
    for m, q in mats:
        want = orc.fast_apply_gate(want, np.asarray(m, dtype=complex), q, n)
    got, _ = run_emulated(psi0, ops, n, 12, 5)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < TOL


def test_density_style_superoperators():
    """Non-unitary 4x4 Kraus superoperators and conj(U) pairs as the density QVM queues them."""
    from referenceqvm.qvm_density import kraus_superoperator
    from referenceqvm.gates import dephasing, relaxation, depolarizing
    nq = 6
    n = 2 * nq
    mats = []
    rs = np.random.RandomState(5)
    for q in range(nq):
        u = random_unitary(2, rs)
        mats += [(u, [q + nq]), (np.conj(u), [q])]
        for kraus in (relaxation(0.1), dephasing(0.2), depolarizing(0.05)):
            mats.append((kraus_superoperator(kraus), [q + nq, q]))
    mats += [(G.CNOT, [0 + nq, 3 + nq]), (np.conj(G.CNOT), [0, 3])]
    ops = [_engine.compile_gate(m, q) for m, q in mats]
    psi0 = random_state(n, 3)
    want = psi0
    for m, q in mats:
        want = orc.fast_apply_gate(want, np.asarray(m, dtype=complex), q, n)
    got, stats = run_emulated(psi0, ops, n, 12, 5)
    assert np.max(np.abs(got - want)) < TOL


def test_specialised_source_compiles_with_nvrtc():
    """The JIT path's generated source (qvm_jit.h) for real planner groups compiles to an sm_100a cubin
    with NVRTC -- no device needed.  Execution parity of those kernels is a GPU test."""
    if REG_BITS != 4:
        pytest.skip("specialised kernels exist for 16 amplitudes per thread")
    n = 16
    ops = compile_spec(orc.random_circuit_spec(n, 10, 3))
    groups = _planner.plan(ops, n, 11, 4)
    codes = [nat.spec_compile_check(n, g.tile, g.ops) for g in groups[:2]]
    assert codes == [0, 0]
    # sharded group: controls / diagonal targets on global positions
    n_local, n_global = 11, 2
    ops = [o for o in compile_spec(orc.random_circuit_spec(n_local + n_global, 6, 21))
           if not (o.kind == nat.OP_DENSE and any(t >= n_local for t in o.targets))]
    g = _planner.plan(ops, n_local, 10, 4)[0]
    assert nat.spec_compile_check(n_local, g.tile, g.ops, n_global=n_global, shard_index=3) == 0
    # relabelling passes (with and without ops): the permuted shared-memory write is generated too
    g = groups[1]
    assert nat.spec_compile_check(n, g.tile, g.ops, bring_low=list(g.tile[-3:]), n_low=4) == 0
    assert nat.spec_compile_check(n, g.tile, [], bring_low=[g.tile[5]], n_low=4) == 0
    # a group with generic dense matrices is left to the interpreter
    rs = np.random.RandomState(0)
    dense = [_engine.compile_gate(rs.standard_normal((4, 4)) + 0j, [3, 9])]
    assert nat.spec_compile_check(12, tuple(range(12)), dense) == 1


# ---- in-tile qubit relabelling (qvm_apply_group_relabel / _planner.run_relabelled) --------------------
def logical_view(psi_phys, pos_of):
    """Amplitudes in logical index order of a vector stored under the layout pos_of."""
    n = len(pos_of)
    logical = np.arange(1 << n, dtype=np.uint64)
    phys = np.zeros_like(logical)
    for q, p in enumerate(pos_of):
        phys |= ((logical >> np.uint64(q)) & np.uint64(1)) << np.uint64(p)
    return psi_phys[phys]


def run_emulated_relabelled(psi, ops, n, tile_bits, low_bits):
    pos_of = list(range(n))
    box = {"psi": psi, "moved": 0}

    def launch(tile, phys_ops, bring):
        new, info = _emu.apply_group(box["psi"], n, tile, phys_ops, reg_bits=REG_BITS, bring_low=bring or (),
                                     n_low=low_bits)
        assert info["status"] == 0
        for p in (bring or ()):
            assert info["new_pos"][list(tile).index(p)] < low_bits          # the library did what was asked
        box["moved"] += sum(1 for a, b in zip(tile, info["new_pos"]) if a != b)
        box["psi"] = new
        return info["new_pos"]

    passes = _planner.run_relabelled(ops, n, pos_of, launch, tile_bits, low_bits)
    return box["psi"], pos_of, passes, box["moved"]


@pytest.mark.parametrize("n,depth,seed,tile,low", [(14, 10, 1, 9, 4), (15, 8, 2, 10, 4), (14, 12, 3, 11, 4),
                                                   (15, 6, 1234, 12, 5), (13, 14, 5, 9, 3)])
def test_relabelled_random_circuit(n, depth, seed, tile, low):
    spec = orc.random_circuit_spec(n, depth, seed)
    ops = []
    for op in compile_spec(spec):
        ops += list(_engine.expand_op(op))
    psi0 = random_state(n, seed)
    want = psi0
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    got, pos_of, passes, moved = run_emulated_relabelled(psi0, ops, n, tile, low)
    assert sorted(pos_of) == list(range(n))
    assert np.max(np.abs(logical_view(got, pos_of) - want)) < TOL
    fixed = len(_planner.plan(ops, n, tile, low))
    assert passes <= fixed and moved > 0, (passes, fixed, moved)


@pytest.mark.parametrize("n,depth,seed,tile,low", [(14, 30, 11, 11, 4), (13, 50, 15, 11, 4), (15, 24, 17, 11, 5)])
def test_deep_circuits_with_full_groups(n, depth, seed, tile, low):
    """Deep circuits on few qubits: every group is full (60-90 ops), so the planner's group search and the encoder's
    round search (bits kept out of a round so that the group finishes in fewer rounds) both have choices to make."""
    spec = orc.random_circuit_spec(n, depth, seed)
    ops = []
    for op in compile_spec(spec):
        ops += list(_engine.expand_op(op))
    psi0 = random_state(n, seed)
    want = psi0
    for name, params, qubits in spec:
        m = G.gate_matrix[name]
        want = orc.fast_apply_gate(want, np.asarray(m(*params) if params else m, dtype=complex), qubits, n)
    got, pos_of, passes, _ = run_emulated_relabelled(psi0, ops, n, tile, low)
    assert sorted(pos_of) == list(range(n))
    assert np.max(np.abs(logical_view(got, pos_of) - want)) < TOL
    assert passes >= 3


def test_relabel_only_pass_and_partial_requests():
    """A pass without ops still relabels; requests for qubits already low, or more requests than low
    positions, are honoured as far as possible and reported truthfully."""
    n, tile = 13, [0, 1, 2, 3, 5, 6, 8, 9, 10, 11, 12]
    psi0 = random_state(n, 21)
    for bring in ([5], [12, 9, 6, 8], [1, 10], [0, 1, 2, 3], [5, 6, 8, 9, 10], []):
        got, info = _emu.apply_group(psi0, n, tile, [], reg_bits=REG_BITS, bring_low=bring, n_low=4)
        assert info["status"] == 0
        pos_of = list(range(n))
        for j, p in enumerate(tile):
            pos_of[p] = info["new_pos"][j]
        assert sorted(pos_of) == list(range(n))
        assert sum(1 for b in bring if pos_of[b] < 4) == min(len(bring), 4)
        assert np.array_equal(logical_view(got, pos_of), psi0)
