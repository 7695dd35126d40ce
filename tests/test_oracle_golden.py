"""Pin the oracle (oracle/qvm_oracle.py) to the reference: every fixture under tests/golden/ was
produced by the unmodified reference code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sps

from _programs import load_golden, oracle_ops, program_from_json, straight_line
from oracle import qvm_oracle as orc
from referenceqvm import gates

TOL = 1e-13


def test_ref_apply_gate_operators_exact():
    meta, arr = load_golden("operators")
    n_ops = 0
    for case in meta:
        if "args" not in case or "vec" in case:
            continue
        g = arr["G%d" % case["k"]]
        ours = orc.ref_apply_gate(g, case["args"], case["n"]).toarray()
        assert np.array_equal(ours, arr[case["key"]]), case["key"]     # exact: same algorithm
        n_ops += 1
    assert n_ops == 40


def test_closed_form_matches_reference_operators():
    meta, arr = load_golden("operators")
    rs = np.random.RandomState(0)
    for case in meta:
        if "args" not in case:
            continue
        g = arr["G%d" % case["k"]]
        if "vec" in case:
            vec, want = arr[case["vec"]], arr[case["key"]]
        else:
            vec = rs.standard_normal(2 ** case["n"]) + 1j * rs.standard_normal(2 ** case["n"])
            want = arr[case["key"]].dot(vec)
        got = orc.fast_apply_gate(vec, g, case["args"], case["n"])
        assert np.max(np.abs(got - want)) < 1e-12 * np.max(np.abs(want)), case["key"]


def test_lifted_gate_fixture_and_errors():
    meta, arr = load_golden("operators")
    for case in meta:
        if "lift" in case:
            got = orc.ref_lifted_gate(case["lift"], gates.SWAP, case["n"]).toarray()
            assert np.array_equal(got, arr[case["key"]])
    for i, n in [(2, 3), (3, 3), (-1, 3), (3, 4)]:       # reference tests/test_unitary_generator.py:34-41
        with pytest.raises(ValueError):
            orc.ref_lifted_gate(i, gates.SWAP, n)


@pytest.mark.parametrize("which", ["ref", "fast"])
def test_wavefunctions(which):
    meta, arr = load_golden("wavefunctions")
    done = 0
    for case in meta:
        if not straight_line(case["program"]):
            continue
        n = case["n"]
        if which == "ref" and n > 10:
            continue
        ops = oracle_ops(case["program"])
        run = orc.ref_run_wavefunction if which == "ref" else orc.fast_run_wavefunction
        psi, _ = run(ops, n)
        assert np.max(np.abs(psi - arr[case["name"]])) < TOL, case["name"]
        done += 1
    assert done >= 24


def test_seeded_measurement_matches_draw_for_draw():
    meta, arr = load_golden("measure")
    done = 0
    for case in meta:
        if "seed" not in case or not straight_line(case["program"]) or ":wf" not in "".join(arr.files) \
                or case["key"] + ":wf" not in arr.files:
            continue
        for run in (orc.ref_run_wavefunction, orc.fast_run_wavefunction):
            rng = np.random.RandomState(case["seed"])
            psi, cmem = run(oracle_ops(case["program"]), case["n"], rng)
            assert np.max(np.abs(psi - arr[case["key"] + ":wf"])) < 1e-12, case["key"]
            want_mem = arr[case["key"] + ":mem"]
            for c, v in cmem.items():
                assert want_mem[c] == v
            assert rng.random_sample() == arr[case["key"] + ":next"][0]     # same number of draws consumed
        done += 1
    assert done >= 20


def _noise_post(vec, n, instr_qubits, noise):
    """Restatement of reference qvm_density.py:267-348 (_post) on the vectorised rho."""
    if noise is None:
        return vec
    get = lambda k, d: (float("inf") if noise.get(k, d) == "inf" else noise.get(k, d))
    t1, t2 = get("T1", 30e-6), get("T2", 30e-6)
    g1, g2 = noise.get("gate_time_1q", 50e-9), noise.get("gate_time_2q", 150e-9)
    span = len(instr_qubits)
    gt = g1 if span == 1 else g2
    if t1 != float("inf"):
        p = 1.0 - np.exp(-gt / t1)
        for q in range(n):
            vec = orc.fast_apply_kraus(vec, gates.relaxation(p), q, n)
    if t2 != float("inf"):
        p = 1.0 - np.exp(-gt / t2)
        for q in range(n):
            vec = orc.fast_apply_kraus(vec, gates.dephasing(p), q, n)
    for key, chan in (("depolarizing", gates.depolarizing), ("bitflip", gates.bit_flip),
                      ("phaseflip", gates.phase_flip), ("bitphaseflip", gates.bitphase_flip)):
        if noise.get(key) is not None:
            for q in range(n):
                vec = orc.fast_apply_kraus(vec, chan(noise[key]), q, n)
    return vec


def test_density_vectorised_matches_reference():
    meta, arr = load_golden("density")
    done = 0
    for case in meta:
        if "program" not in case or "seed" in case:
            continue
        n = case["n"]
        vec = np.zeros(4 ** n, dtype=complex)
        vec[0] = 1.0
        for op in oracle_ops(case["program"]):
            vec = orc.fast_density_gate(vec, op[1], op[2], n)
            vec = _noise_post(vec, n, op[2], case["noise"])
        want = arr[case["name"]]
        assert np.max(np.abs(orc.vec_to_rho(vec, n) - want)) < TOL, case["name"]
        done += 1
    assert done >= 10


def test_kraus_sparse_and_superoperator_forms():
    meta, arr = load_golden("density")
    rho0 = arr["kraus_rho0"]
    done = 0
    for case in meta:
        if "channel" not in case:
            continue
        ks = gates.noise_gates[case["channel"]](case["p"])
        want = arr[case["name"]]
        got_ref = orc.ref_apply_kraus(sps.csc_matrix(rho0), ks, case["qubit"], case["n"]).toarray()
        got_fast = orc.vec_to_rho(orc.fast_apply_kraus(orc.rho_to_vec(rho0), ks, case["qubit"], case["n"]), case["n"])
        assert np.max(np.abs(got_ref - want)) < TOL
        assert np.max(np.abs(got_fast - want)) < TOL
        done += 1
    assert done == 18


def test_density_seeded_measure():
    meta, arr = load_golden("density")
    for case in meta:
        if "seed" not in case or case["noise"] is not None:
            continue
        n = case["n"]
        rng = np.random.RandomState(case["seed"])
        vec = np.zeros(4 ** n, dtype=complex)
        vec[0] = 1.0
        mem = np.zeros(8, dtype=np.int8)
        for op in oracle_ops(case["program"]):
            if op[0] == "gate":
                vec = orc.fast_density_gate(vec, op[1], op[2], n)
            else:
                p0 = orc.fast_density_prob_zero(vec, op[1], n)
                b = 0 if rng.random_sample() < p0 else 1
                vec = orc.fast_density_collapse(vec, op[1], b, p0 if b == 0 else 1 - p0, n)
                mem[op[2]] = b
        assert np.array_equal(mem, arr[case["name"] + ":mem"])
        assert np.max(np.abs(orc.vec_to_rho(vec, n) - arr[case["name"]])) < 1e-12


def test_inverse_cdf_matches_numpy_choice_stream():
    meta, arr = load_golden("sampling")
    for case in meta:
        if case["name"] == "density_run":
            continue
        n, trials = case["n"], case["trials"]
        rng = np.random.RandomState(case["seed"])
        idx = orc.inverse_cdf_sample(arr[case["name"] + ":probs"], rng.random_sample(trials))
        bits = np.array([orc.bits_big_endian(i, n) for i in idx], dtype=np.int8)
        assert np.array_equal(bits, arr[case["name"] + ":draws"]), case["name"]


def test_unitary_fixture_via_closed_form():
    meta, arr = load_golden("unitary")
    for case in meta:
        n = case["n"]
        u = np.eye(2 ** n, dtype=complex).reshape(-1)
        for op in oracle_ops(case["program"]):
            u = orc.fast_apply_gate(u, op[1], [a + n for a in op[2]], 2 * n)     # gate acts on the row index
        assert np.max(np.abs(u.reshape(2 ** n, 2 ** n) - arr[case["name"]])) < TOL, case["name"]


def test_generators_are_the_documented_ones():
    spec = orc.random_circuit_spec(10, 20, 1234)
    assert len(spec) == 200
    assert {s[0] for s in spec} == {"H", "RZ", "CNOT", "CPHASE"}
    q = orc.qft_spec(8)
    assert len(q) == 8 + 28 + 4 and q[0] == ("H", [], [7]) and q[1][0] == "CPHASE" and q[1][2] == [6, 7]
    assert len(orc.qft_spec(28)) == 420


def test_expectation_oracle_against_reference_pieces():
    """oracle.ref_tensor_up / expectation_state / expectation_density vs values assembled by the reference's own
    wavefunction, density and tensor_up (tests/golden/make_golden.py::golden_expectation)."""
    from _programs import load_golden
    from referenceqvm import gates as G
    meta, arrays = load_golden("expectation")
    for case in meta:
        n = case["n"]
        terms = [(complex(t["re"], t["im"]), {int(q): p for q, p in t["ops"].items()}) for t in case["terms"]]
        ops = []
        for name, params, qubits in case["spec"]:
            m = G.gate_matrix[name]
            ops.append(("gate", np.asarray(m(*params) if params else m, dtype=complex), qubits))
        psi, _ = orc.fast_run_wavefunction(ops, n)
        want_psi, want_rho = arrays[case["key"]]
        assert abs(orc.expectation_state(psi, terms, n) - want_psi) < 1e-13
        assert abs(orc.fast_expectation_state(psi, terms, n) - want_psi) < 1e-13
        # density with T2 dephasing on every qubit after every gate (the fixture's noise model)
        vec = np.zeros(4 ** n, dtype=complex)
        vec[0] = 1.0
        for _, m, q in ops:
            vec = orc.fast_density_gate(vec, m, q, n)
            p = 1.0 - np.exp(-(50e-9 if len(q) == 1 else 150e-9) / 30e-6)
            for i in range(n):
                vec = orc.fast_apply_kraus(vec, G.dephasing(p), i, n)
        assert abs(orc.expectation_density(orc.vec_to_rho(vec, n), terms, n) - want_rho) < 1e-13
