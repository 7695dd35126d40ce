"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) in this container.

Run once, here (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference imports through (1) ``collections.Sequence = collections.abc.Sequence`` (removed in
Python 3.10; reference ``unitary_generator.py:24``) and (2) the in-repo ``pyquil`` stand-in, which
supplies instruction containers only -- every number stored below is produced by the reference's own
arithmetic (``gates.py``, ``unitary_generator.py``, ``qvm_wavefunction.py``, ``qvm_density.py``,
``qvm_unitary.py``).

Outputs (all under tests/golden/):
  gates.npz            every standard gate matrix / Kraus set at fixed parameters
  operators.npz        apply_gate / lifted_gate full operators (all argument permutations, n=4)
  wavefunctions.json   programs (serialised) ; wavefunctions.npz  their reference amplitudes
  density.json/.npz    density-matrix programs incl. noise models, Kraus applications
  measure.json/.npz    seeded MEASURE programs: classical memory + collapsed state, RNG stream position
  sampling.json/.npz   seeded density run_and_measure draws (np.random.choice stream)
  unitary.json/.npz    QVM_Unitary outputs
  expectation.json/.npz  psi^dagger tensor_up(PauliSum) psi and tr(rho tensor_up(PauliSum)) from the reference's pieces
  stabilizer.json/.npz QVM_Stabilizer: tableaus, projected wavefunctions, densities, seeded run() outcomes
"""
import collections
import collections.abc
import itertools
import json
import os
import sys

collections.Sequence = collections.abc.Sequence
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "reference-qvm_b200", "_compat"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import scipy.sparse as sps  # noqa: E402

import pyquil.gates as pg  # noqa: E402
from pyquil.quil import Program  # noqa: E402
from pyquil.quilbase import (Gate, Measurement, Halt, Jump, JumpTarget, JumpWhen, JumpUnless,  # noqa: E402
                             UnaryClassicalInstruction, BinaryClassicalInstruction, Nop, Wait)
import referenceqvm  # noqa: E402
from referenceqvm.api import QVMConnection  # noqa: E402
from referenceqvm.gates import gate_matrix, noise_gates, utility_gates  # noqa: E402
from referenceqvm.qvm_density import NoiseModel, INFINITY  # noqa: E402
from referenceqvm.unitary_generator import apply_gate, lifted_gate  # noqa: E402
from referenceqvm.tests.test_data import data_containers as dc  # noqa: E402

from oracle.qvm_oracle import random_circuit_spec, qft_spec  # noqa: E402  (generator only, no arithmetic)

assert referenceqvm.__file__.startswith("/root/reference/"), referenceqvm.__file__


# ---------------------------------------------------------------------------------------------
def ser_program(prog):
    """pyquil Program -> JSON-able instruction list (the inverse lives in tests/_programs.py)."""
    out = []
    for ins in prog:
        if isinstance(ins, Gate):
            out.append({"op": "gate", "name": ins.name, "params": [complex(p).real if np.isreal(p) else p
                                                                   for p in ins.params],
                        "qubits": [q.index for q in ins.qubits]})
        elif isinstance(ins, Measurement):
            out.append({"op": "measure", "q": ins.qubit.index, "c": ins.classical_reg.address})
        elif isinstance(ins, UnaryClassicalInstruction):
            out.append({"op": ins.op, "c": ins.target.address})
        elif isinstance(ins, BinaryClassicalInstruction):
            out.append({"op": ins.op, "l": ins.left.address, "r": ins.right.address})
        elif isinstance(ins, JumpTarget):
            out.append({"op": "LABEL", "label": ins.label.name})
        elif isinstance(ins, (JumpWhen, JumpUnless)):
            out.append({"op": ins.op, "label": ins.target.name, "c": ins.condition.address})
        elif isinstance(ins, Jump):
            out.append({"op": "JUMP", "label": ins.target.name})
        elif isinstance(ins, Halt):
            out.append({"op": "HALT"})
        elif isinstance(ins, (Nop, Wait)):
            out.append({"op": ins.op})
        else:
            raise TypeError(ins)
    defs = [{"name": d.name, "re": np.real(d.matrix).tolist(), "im": np.imag(d.matrix).tolist()}
            for d in prog.defined_gates]
    return {"instructions": out, "defgates": defs}


def spec_program(spec):
    p = Program()
    for name, params, qubits in spec:
        p.inst(getattr(pg, name)(*(list(params) + list(qubits))))
    return p


def dump(name, meta, arrays):
    with open(os.path.join(HERE, name + ".json"), "w") as fh:
        json.dump(meta, fh, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("%-14s %3d cases, %d arrays" % (name, len(meta), len(arrays)))


# ---------------------------------------------------------------------------------------------
def golden_gates():
    arrays = {}
    theta = [0.0, 0.3928244130249029, -1.3778211380875056, np.pi / 3, 2.0 * np.pi, 4.71572463191]
    for name, g in gate_matrix.items():
        if callable(g):
            if name == "BARENCO":
                arrays["BARENCO"] = np.asarray(g(0.3, 0.7, -1.1), dtype=np.complex128)
                continue
            for i, t in enumerate(theta):
                arrays["%s@%d" % (name, i)] = np.asarray(g(t), dtype=np.complex128)
        else:
            arrays[name] = np.asarray(g, dtype=np.complex128)
    arrays["theta"] = np.array(theta)
    for name, g in utility_gates.items():
        arrays[name] = np.asarray(g, dtype=np.complex128)
    for name, f in noise_gates.items():
        for i, p in enumerate([0.0, 0.372, 0.75, 1.0 - np.exp(-1.0 / 600.0)]):
            arrays["noise:%s@%d" % (name, i)] = np.stack([np.asarray(k, dtype=np.complex128) for k in f(p)])
    arrays["noise_p"] = np.array([0.0, 0.372, 0.75, 1.0 - np.exp(-1.0 / 600.0)])
    np.savez_compressed(os.path.join(HERE, "gates.npz"), **arrays)
    print("gates          %d arrays" % len(arrays))


def golden_operators():
    rs = np.random.RandomState(20261019)
    arrays, meta = {}, []
    n = 4
    for k in (1, 2, 3):
        g = rs.standard_normal((2 ** k, 2 ** k)) + 1j * rs.standard_normal((2 ** k, 2 ** k))  # non-unitary on purpose
        arrays["G%d" % k] = g
        for args in itertools.permutations(range(n), k):
            key = "op_n%d_%s" % (n, "".join(map(str, args)))
            arrays[key] = np.asarray(apply_gate(g, list(args), n).toarray())
            meta.append({"key": key, "n": n, "k": k, "args": list(args)})
    # far-apart qubits on a larger register: store operator . fixed vector
    n = 9
    vec = rs.standard_normal(2 ** n) + 1j * rs.standard_normal(2 ** n)
    arrays["vec_n9"] = vec
    for args in [(0,), (8,), (4,), (0, 8), (8, 0), (3, 7), (7, 2), (0, 4, 8), (8, 4, 0), (5, 0, 7), (2, 8, 1)]:
        key = "apply_n%d_%s" % (n, "".join(map(str, args)))
        arrays[key] = apply_gate(arrays["G%d" % len(args)], list(args), n).dot(vec)
        meta.append({"key": key, "n": n, "k": len(args), "args": list(args), "vec": "vec_n9"})
    for i, n_ in [(0, 3), (1, 3), (2, 4), (0, 2)]:
        key = "lift_swap_%d_%d" % (i, n_)
        arrays[key] = lifted_gate(i, gate_matrix["SWAP"], n_).toarray()
        meta.append({"key": key, "n": n_, "lift": i})
    dump("operators", meta, arrays)


def golden_wavefunctions():
    qvm = QVMConnection(type_trans="wavefunction")
    meta, arrays = [], {}

    def add(name, prog):
        wf, _ = qvm.wavefunction(prog)
        arrays[name] = np.asarray(wf.amplitudes)
        meta.append({"name": name, "n": int(qvm.num_qubits), "program": ser_program(prog)})

    add("bell", Program().inst([pg.H(0), pg.CNOT(0, 1)]))
    add("occupation", Program().inst([pg.X(0), pg.X(1), pg.I(2), pg.I(3)]))
    add("hadamard", Program(pg.H(0)))
    add("qft8_fixture", Program(dc.QFT_8_INSTRUCTIONS))
    for k in sorted(dc.ARBITRARY_STATE_GEN_INSTRUCTIONS):
        add("stateprep_%d" % k, Program(dc.ARBITRARY_STATE_GEN_INSTRUCTIONS[k]))
    # random circuits from the benchmark generator (config 1 = n10 d20, three seeds)
    for n, depth, seeds in [(2, 6, (1,)), (3, 8, (1, 2)), (5, 10, (1234,)), (8, 10, (1234,)),
                            (10, 20, (1234, 1, 2)), (12, 4, (1234,))]:
        for seed in seeds:
            add("random_n%d_d%d_s%d" % (n, depth, seed), spec_program(random_circuit_spec(n, depth, seed)))
    # QFT on a non-trivial product state
    for n in (5, 9):
        rs = np.random.RandomState(n)
        prep = Program()
        for q in range(n):
            prep.inst(pg.RY(float(rs.uniform(0, np.pi)), q), pg.RZ(float(rs.uniform(0, 2 * np.pi)), q))
        add("qft_prod_n%d" % n, prep + spec_program(qft_spec(n)))
    # every standard gate at least once, scattered qubits, incl. 3-qubit gates and far pairs
    zoo = Program()
    zoo.inst(pg.H(0), pg.H(3), pg.H(6), pg.RX(0.3, 1), pg.RY(-1.2, 2), pg.RZ(0.7, 4), pg.X(5), pg.Y(6),
             pg.Z(0), pg.S(3), pg.T(6), pg.PHASE(1.1, 1), pg.CZ(0, 6), pg.CNOT(6, 0), pg.CNOT(2, 5),
             pg.CCNOT(0, 3, 6), pg.CCNOT(6, 1, 2), pg.CPHASE00(0.4, 1, 4), pg.CPHASE01(0.5, 4, 1),
             pg.CPHASE10(0.6, 2, 6), pg.CPHASE(0.9, 6, 3), pg.SWAP(0, 6), pg.SWAP(2, 3),
             pg.CSWAP(1, 5, 2), pg.CSWAP(6, 0, 3), pg.ISWAP(1, 4), pg.ISWAP(5, 0), pg.PSWAP(0.8, 6, 2),
             pg.I(4))
    zoo.inst(("BARENCO", [0.3, 0.7, -1.1], 2, 5))
    add("gate_zoo_n7", zoo)
    # DefGates: 1-, 2- and 3-qubit, dense non-controlled, reversed / scattered arguments
    rs = np.random.RandomState(7)

    def haar(dim):
        z = rs.standard_normal((dim, dim)) + 1j * rs.standard_normal((dim, dim))
        q, r = np.linalg.qr(z)
        return q * (np.diag(r) / np.abs(np.diag(r)))

    dg = Program()
    dg.defgate("U1", haar(2)).defgate("U2", haar(4)).defgate("U3", haar(8))
    dg.inst(pg.H(0), pg.H(1), pg.H(2), pg.H(3), pg.H(4))
    dg.inst(("U1", 3), ("U2", 0, 4), ("U2", 4, 1), ("U3", 0, 2, 4), ("U3", 3, 1, 0), ("U2", 2, 3), ("U1", 0))
    add("defgates_n5", dg)
    # control flow (host side, but it drives the path): HALT, if_then both ways, classical ops
    halt = Program(pg.X(0)); halt.inst(pg.HALT); halt.inst(pg.X(0))
    add("halt", halt)
    for name, prep in (("ifthen_true", pg.TRUE(0)), ("ifthen_false", pg.FALSE(0))):
        main = Program().inst(pg.X(0))
        main.if_then(0, Program().inst(pg.X(0)), Program().inst(pg.H(1)))
        add(name, Program().inst(prep) + main)
    dump("wavefunctions", meta, arrays)


def golden_measure():
    qvm = QVMConnection(type_trans="wavefunction")
    meta, arrays = [], {}
    progs = {
        "coin2": Program().inst(pg.H(0)).measure(0, [0]).measure(0, [1]),
        "ghz_measure": Program().inst(pg.H(0), pg.CNOT(0, 1), pg.CNOT(1, 2), pg.RX(0.7, 3), pg.CNOT(3, 4))
        .measure(1, [5]).inst(pg.H(1), pg.RY(0.4, 2)).measure(2, [0]).measure(4, [3]).measure(1, [1]),
        "biased": Program().inst(pg.RX(np.pi / 3, 0), pg.RX(np.pi / 5, 1), pg.CNOT(0, 2)).measure(0, [2])
        .measure(1, [3]).measure(2, [4]),
        "while": Program(pg.TRUE([2])).while_do(2, Program(pg.X(0), pg.H(0)).measure(0, 2)),
    }
    rc = spec_program(random_circuit_spec(6, 6, 5))
    for q in range(6):
        rc.measure(q, [q])
    progs["random6_measure_all"] = rc
    for name, prog in progs.items():
        for seed in (0, 1, 2, 3, 11):
            np.random.seed(seed)
            wf, mem = qvm.wavefunction(prog)
            nxt = np.random.random()           # pins how many draws the reference consumed
            key = "%s@%d" % (name, seed)
            arrays[key + ":wf"] = np.asarray(wf.amplitudes)
            arrays[key + ":mem"] = np.asarray(mem[:16], dtype=np.int8)
            arrays[key + ":next"] = np.array([nxt])
            meta.append({"name": name, "seed": seed, "key": key, "n": int(qvm.num_qubits),
                         "program": ser_program(prog)})
    # run(): per-trial classical results under a seed
    np.random.seed(42)
    res = qvm.run(progs["biased"], [2, 3, 4], trials=50)
    arrays["run_biased@42"] = np.asarray(res, dtype=np.int8)
    np.random.seed(43)
    res = qvm.run_and_measure(spec_program(random_circuit_spec(4, 4, 9)), [0, 1, 2, 3, 6], trials=40)
    arrays["ram_random4@43"] = np.asarray(res, dtype=np.int8)
    meta.append({"name": "ram_random4", "seed": 43, "key": "ram_random4@43", "qubits": [0, 1, 2, 3, 6],
                 "trials": 40, "program": ser_program(spec_program(random_circuit_spec(4, 4, 9)))})
    dump("measure", meta, arrays)


def golden_density():
    meta, arrays = [], {}

    def add(name, prog, noise=None, noise_kw=None):
        qvm = QVMConnection(type_trans="density", noise_model=noise)
        rho = qvm.density(prog)
        arrays[name] = np.asarray(rho.todense())
        meta.append({"name": name, "n": int(qvm.num_qubits), "program": ser_program(prog),
                     "noise": noise_kw})

    qaoa2 = Program().inst([pg.RY(np.pi / 2, 0), pg.RX(np.pi, 0), pg.RY(np.pi / 2, 1), pg.RX(np.pi, 1),
                            pg.CNOT(0, 1), pg.RX(-np.pi / 2, 1), pg.RY(4.71572463191, 1),
                            pg.RX(np.pi / 2, 1), pg.CNOT(0, 1), pg.RX(-2 * 2.74973750579, 0),
                            pg.RX(-2 * 2.74973750579, 1)])
    add("qaoa2", qaoa2)
    add("random_n4_pure", spec_program(random_circuit_spec(4, 6, 3)))
    noise_cases = {
        "t2": dict(T1="inf", T2=30e-6, ro_fidelity=1.0),
        "t1": dict(T1=30e-6, T2="inf", ro_fidelity=1.0),
        "t1t2": dict(ro_fidelity=1.0),
        "depol": dict(T1="inf", T2="inf", depolarizing=0.05),
        "flips": dict(T1="inf", T2="inf", bitflip=0.02, phaseflip=0.03),
        "bitphase": dict(T1="inf", T2="inf", phaseflip=0.01, bitphaseflip=0.04),
        "config2": dict(T1="inf", T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0),
    }

    def mk(kw):
        return NoiseModel(**{k: (INFINITY if v == "inf" else v) for k, v in kw.items()})

    circ3 = Program().inst(pg.H(0), pg.H(1), pg.H(2)) + spec_program(random_circuit_spec(3, 4, 21))
    circ4 = Program().inst(pg.H(0), pg.H(1), pg.H(2), pg.H(3)) + spec_program(random_circuit_spec(4, 4, 1234))
    ccn = Program().inst(pg.H(0), pg.H(1), pg.CCNOT(0, 1, 2), pg.RX(0.4, 2))
    for tag, kw in noise_cases.items():
        add("noisy3_" + tag, circ3, mk(kw), kw)
    add("noisy4_config2", circ4, mk(noise_cases["config2"]), noise_cases["config2"])
    add("noisy3q_gate_t1t2", ccn, mk(noise_cases["t1t2"]), noise_cases["t1t2"])
    circ6 = Program().inst([pg.H(q) for q in range(6)]) + spec_program(random_circuit_spec(6, 2, 1234))
    add("noisy6_config2", circ6, mk(noise_cases["config2"]), noise_cases["config2"])

    # apply_kraus on a random (non-physical is fine, it is linear) rho, n=3, every channel x qubit
    rs = np.random.RandomState(5)
    n = 3
    rho0 = rs.standard_normal((8, 8)) + 1j * rs.standard_normal((8, 8))
    arrays["kraus_rho0"] = rho0
    for cname, f in noise_gates.items():
        for q in range(n):
            qvm = QVMConnection(type_trans="density")
            qvm._density = sps.csc_matrix(rho0)
            qvm.num_qubits = n
            out = qvm.apply_kraus(f(0.372), q)
            key = "kraus_%s_q%d" % (cname, q)
            arrays[key] = np.asarray(out.todense())
            meta.append({"name": key, "n": n, "channel": cname, "p": 0.372, "qubit": q})

    # seeded MEASURE in the density QVM (no noise model)
    mprog = Program().inst(pg.H(0), pg.CNOT(0, 1), pg.RX(0.9, 2)).measure(0, [0]).measure(2, [1]).measure(1, [2])
    for seed in (0, 1, 5):
        np.random.seed(seed)
        qvm = QVMConnection(type_trans="density")
        rho = qvm.density(mprog)
        key = "dmeasure@%d" % seed
        arrays[key] = np.asarray(rho.todense())
        arrays[key + ":mem"] = np.asarray(qvm.classical_memory[:8], dtype=np.int8)
        meta.append({"name": key, "n": 3, "seed": seed, "program": ser_program(mprog), "noise": None})
    # the reference's (non-physical) readout-noise pre-hook on MEASURE, reproduced bug-for-bug
    for seed in (0, 3):
        np.random.seed(seed)
        kw = dict(T1="inf", T2="inf", ro_fidelity=0.95)
        qvm = QVMConnection(type_trans="density", noise_model=mk(kw))
        p = Program().inst(pg.RX(0.9, 0), pg.H(1)).measure(0, [0])
        rho = qvm.density(p)
        key = "dmeasure_ro@%d" % seed
        arrays[key] = np.asarray(rho.todense())
        arrays[key + ":mem"] = np.asarray(qvm.classical_memory[:8], dtype=np.int8)
        meta.append({"name": key, "n": 2, "seed": seed, "program": ser_program(p), "noise": kw})
    dump("density", meta, arrays)


def golden_sampling():
    meta, arrays = [], {}
    progs = {"coin": Program().inst(pg.H(0)),
             "bell": Program().inst(pg.H(0), pg.CNOT(0, 1)),
             "biased": Program().inst(pg.RX(np.pi / 3, 0)),
             "random5": spec_program(random_circuit_spec(5, 6, 77))}
    for name, prog in progs.items():
        qvm = QVMConnection(type_trans="density")
        np.random.seed(99)
        res = qvm.run_and_measure(prog, trials=2000)
        arrays[name + ":draws"] = np.asarray(res, dtype=np.int8)
        arrays[name + ":probs"] = np.real(np.asarray(qvm._density.todense()).diagonal())
        meta.append({"name": name, "n": int(qvm.num_qubits), "seed": 99, "trials": 2000,
                     "program": ser_program(prog)})
    np.random.seed(7)
    qvm = QVMConnection(type_trans="density")
    mp = Program().inst(pg.H(0)).measure(0, [0]).measure(0, [1])
    res = qvm.run(mp, [0, 1], trials=30)
    arrays["density_run:draws"] = np.asarray(res, dtype=np.int8)
    meta.append({"name": "density_run", "n": 1, "seed": 7, "trials": 30, "addresses": [0, 1],
                 "program": ser_program(mp)})
    dump("sampling", meta, arrays)


def golden_unitary():
    qvm = QVMConnection(type_trans="unitary")
    meta, arrays = [], {}
    progs = {"hhh": Program().inst([pg.H(0), pg.H(1), pg.H(0)]),
             "hxyz": Program().inst([pg.H(0), pg.X(1), pg.Y(2), pg.Z(3)]),
             "xcc": Program().inst([pg.X(2), pg.CNOT(2, 1), pg.CNOT(1, 0)]),
             "empty": Program(),
             "random5": spec_program(random_circuit_spec(5, 4, 8)),
             "zoo": Program().inst(pg.H(0), pg.CCNOT(3, 0, 2), pg.CSWAP(1, 3, 0), pg.ISWAP(0, 3),
                                   pg.PSWAP(0.3, 2, 0), pg.CPHASE01(0.2, 1, 3), pg.RY(0.5, 2))}
    for name, prog in progs.items():
        u = qvm.unitary(prog)
        arrays[name] = np.asarray(u, dtype=np.complex128)
        meta.append({"name": name, "n": int(qvm.num_qubits), "program": ser_program(prog)})
    dump("unitary", meta, arrays)


def golden_expectation():
    """Expectation values.  The reference's QVMs declare expectation() and raise NotImplementedError
    (qvm_unitary.py:99-112, qvm_density.py:396-408); what is stored here is assembled from the reference's own
    pieces only: its wavefunction / density of a program and its tensor_up(PauliSum) operator
    (unitary_generator.py:366-411): value = psi^dagger . tensor_up . psi, resp. tr(rho . tensor_up)."""
    from pyquil.paulis import PauliSum, PauliTerm
    from referenceqvm.unitary_generator import tensor_up
    meta, arrays = [], {}
    wq = QVMConnection()
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    dq = QVMConnection(type_trans="density", noise_model=nm)
    rs = np.random.RandomState(77)
    for case, (n, depth, seed) in enumerate(((3, 4, 2), (5, 5, 9), (6, 3, 1234))):
        spec = random_circuit_spec(n, depth, seed)
        prog = spec_program(spec)
        psi = np.asarray(wq.wavefunction(prog)[0].amplitudes)
        rho = np.asarray(dq.density(prog).todense())
        sums = []
        for k in range(6):
            terms = []
            for _ in range(1 + k % 3):
                term = PauliTerm("I", 0, complex(rs.uniform(-1, 1), 0.0))
                for q in rs.permutation(n)[:int(rs.randint(0, n + 1))]:
                    term = term * PauliTerm("XYZ"[int(rs.randint(3))], int(q))
                terms.append(term)
            sums.append(PauliSum(terms))
        for j, ps in enumerate(sums):
            op = tensor_up(ps, n)
            key = "c%d_s%d" % (case, j)
            arrays[key] = np.array([np.vdot(psi, op.dot(psi)), np.trace(rho.dot(op))])
            meta.append({"key": key, "n": n, "spec": [[a, list(b), list(c)] for a, b, c in spec],
                         "terms": [{"re": t.coefficient.real, "im": t.coefficient.imag,
                                    "ops": {str(q): p for q, p in t._ops.items()}} for t in ps.terms]})
    dump("expectation", meta, arrays)


def golden_stabilizer():
    """The reference's tableau simulator (qvm_stabilizer.py) on Clifford programs: final tableau, projected
    wavefunction, density matrix, and seeded run() outcomes (one np.random.random() per random MEASURE)."""
    rs = np.random.RandomState(42)
    meta, arrays = [], {}
    sq = QVMConnection(type_trans="stabilizer")
    progs = {"bell": Program().inst(pg.H(0), pg.CNOT(0, 1)),
             "ghz4_s": Program().inst(pg.H(0), pg.CNOT(0, 1), pg.CNOT(1, 2), pg.CNOT(2, 3), pg.S(3), pg.S(0)),
             "hs": Program().inst(pg.H(0), pg.S(0), pg.S(0), pg.H(1), pg.S(1), pg.I(2))}
    for k in range(4):                       # random Clifford circuits (the generators the reference executes)
        n = 2 + k
        prog = Program()
        for _ in range(12 * n):
            g = rs.randint(3)
            if g == 0:
                prog.inst(pg.H(int(rs.randint(n))))
            elif g == 1:
                prog.inst(pg.S(int(rs.randint(n))))
            else:
                a = int(rs.randint(n))
                b = int((a + 1 + rs.randint(n - 1)) % n)
                prog.inst(pg.CNOT(a, b))
        progs["random%d" % n] = prog
    for name, prog in progs.items():
        wf = sq.wavefunction(prog)
        arrays[name + ":tableau"] = np.array(sq.tableau)
        arrays[name + ":wf"] = np.asarray(wf.amplitudes, dtype=np.complex128)
        arrays[name + ":rho"] = np.asarray(sq.density(prog), dtype=np.complex128)
        mprog = Program().inst(prog)
        nq = int(sq.num_qubits)
        for q in range(nq):
            mprog.measure(q, [q])
        np.random.seed(1000 + len(meta))
        try:
            arrays[name + ":runs"] = np.array(sq.run(mprog, list(range(nq)), trials=25))
            arrays[name + ":rng_after"] = np.array([np.random.random()])
            has_runs = True
        except ValueError:
            # the reference's rowsum also runs over DESTABILIZER rows, whose products may carry a phase of +-i
            # (Aaronson-Gottesman ignore those signs); it raises "impossible value for the phase_accumulator"
            has_runs = False
        meta.append({"name": name, "n": nq, "program": ser_program(prog), "seed": 1000 + len(meta), "runs": has_runs})
    dump("stabilizer", meta, arrays)


if __name__ == "__main__":
    only = [a[2:-5] for a in sys.argv[1:] if a.startswith("--") and a.endswith("-only")]
    if only:
        for name in only:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_gates()
    golden_operators()
    golden_wavefunctions()
    golden_measure()
    golden_density()
    golden_sampling()
    golden_unitary()
    golden_expectation()
    golden_stabilizer()
    os.system("du -sh %s" % HERE)
