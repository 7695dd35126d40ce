"""Run the staged, UNMODIFIED reference (oracle/_ref/referenceqvm) on a benchmark circuit and time it.
BENCH INFRASTRUCTURE: executed only by bench.py's CPU legs, as a subprocess.

    python oracle/run_reference.py --qubits 10 --depth 20 --seed 1234 [--copies C] [--density] [--out amps.npy]

The reference is pure Python on numpy / scipy.sparse / pyquil.  Two things it needs that this image lacks are
supplied from outside its files: ``collections.Sequence`` (removed in Python 3.10; the reference imports it
at unitary_generator.py:24) and ``pyquil`` (un-vendored, absent here: the repository's container-only
stand-in ``reference-qvm_b200/_compat/pyquil`` provides the instruction containers; it does no arithmetic).
The program goes through the reference's own public API: ``QVMConnection().wavefunction(program)`` or
``QVMConnection(type_trans='density', noise_model=...).density(program)``.

The algorithm is serial (scipy.sparse products), so "all host cores" = C independent copies of the same
circuit side by side; prints one JSON line {gates, copies, seconds (max over copies), gate_apps_per_s}.
"""
import argparse
import collections
import collections.abc
import json
import multiprocessing as mp
import os
import sys
import time

collections.Sequence = collections.abc.Sequence
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
# the staged reference first, then the pyquil stand-in; this repository's own `referenceqvm` package is
# deliberately NOT on the path
sys.path[:0] = [os.path.join(HERE, "_ref"), os.path.join(REPO, "reference-qvm_b200", "_compat"), REPO]


def build(args):
    from pyquil.quil import Program
    from pyquil.quilbase import Gate
    from workloads import density_circuit_spec, random_circuit_spec
    spec = density_circuit_spec(args.qubits, args.depth, args.seed) if args.density else \
        random_circuit_spec(args.qubits, args.depth, args.seed)
    prog = Program()
    for name, params, qubits in spec:
        prog.inst(Gate(name, params, qubits))
    return prog, len(spec)


def one_copy(task):
    args, want_out = task
    import numpy as np
    import referenceqvm
    assert os.path.realpath(referenceqvm.__file__).startswith(os.path.realpath(os.path.join(HERE, "_ref"))), \
        referenceqvm.__file__
    from referenceqvm.api import QVMConnection
    prog, gates = build(args)
    if args.density:
        from referenceqvm.qvm_density import INFINITY, NoiseModel
        nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
        qvm = QVMConnection(type_trans="density", noise_model=nm)
        t0 = time.perf_counter()
        rho = qvm.density(prog)
        dt = time.perf_counter() - t0
        result = np.asarray(rho.todense())
    else:
        qvm = QVMConnection()
        t0 = time.perf_counter()
        wf, _ = qvm.wavefunction(prog)
        dt = time.perf_counter() - t0
        result = np.asarray(wf.amplitudes)
    if want_out:
        np.save(want_out, result)
    return gates, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=10)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--copies", type=int, default=1)
    ap.add_argument("--density", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(HERE, "_ref", "referenceqvm")):
        print(json.dumps({"unavailable": "oracle/_ref/referenceqvm is not staged (python oracle/stage_reference.py)"}))
        return
    tasks = [(args, args.out if c == 0 else None) for c in range(args.copies)]
    if args.copies == 1:
        results = [one_copy(tasks[0])]
    else:
        with mp.get_context("fork").Pool(args.copies) as pool:
            results = pool.map(one_copy, tasks)
    wall = max(dt for _, dt in results)
    gates = results[0][0]
    print(json.dumps({"gates": gates, "copies": args.copies, "seconds": wall,
                      "gate_apps_per_s": gates * args.copies / wall}))


if __name__ == "__main__":
    main()
