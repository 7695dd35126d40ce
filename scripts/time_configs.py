"""Wall-clock timings of BASELINE.json configs 1, 2 and 3 through the public API (not bench lines)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "reference-qvm_b200"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import referenceqvm  # noqa: E402,F401
from _programs import spec_program  # noqa: E402
import workloads as orc  # noqa: E402
from referenceqvm.api import QVMConnection  # noqa: E402
from referenceqvm.qvm_density import INFINITY, NoiseModel  # noqa: E402


def config1():
    # config 1: 10-qubit random circuit, depth 20, QVMConnection.wavefunction (16 KiB state: latency-bound)
    n, depth = 10, 20
    spec = orc.random_circuit_spec(n, depth, 1234)
    prog = spec_program(spec)
    qvm = QVMConnection()
    best = None
    for rep in range(20):
        t0 = time.perf_counter()
        wf, _ = qvm.wavefunction(prog)
        amps = np.asarray(wf.amplitudes)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    st = qvm._engine.stats()
    print("config1 wavefunction n=10 depth=20: %d gates, best of 20 %.3f ms (%.0f gate-apps/s), norm %.15f, launches %d"
          % (len(spec), 1e3 * best, len(spec) / best, float(np.vdot(amps, amps).real), st["kernel_launches"]))


def main():
    config1()
    # config 2: 14-qubit density matrix, T2 dephasing on every qubit after every instruction
    n = 14
    spec = [("H", [], [q]) for q in range(n)] + orc.random_circuit_spec(n, 4, 1234)
    nm = NoiseModel(T1=INFINITY, T2=30e-6, gate_time_1q=50e-9, gate_time_2q=150e-9, ro_fidelity=1.0)
    qvm = QVMConnection(type_trans="density", noise_model=nm)
    prog = spec_program(spec)
    for rep in range(2):
        t0 = time.perf_counter()
        rho = qvm.density(prog)
        tr = float(np.sum(rho.diagonal()).real)
        dt = time.perf_counter() - t0
        st = qvm._engine.stats()
    print("config2 density n=14: %d instructions in %.3f s (%.2f ms/instruction), trace %.12f, hbm passes %d"
          % (len(spec), dt, 1e3 * dt / len(spec), tr, st["hbm_passes"]))

    # config 3: 28-qubit QFT (420 gates) + MEASURE all
    n = 28
    qspec = orc.qft_spec(n)
    qprog = spec_program(qspec)
    for q in range(n):
        qprog.measure(q, [q])
    wq = QVMConnection()
    np.random.seed(1)
    for rep in range(2):
        t0 = time.perf_counter()
        _, mem = wq.wavefunction(qprog, list(range(n))) if False else (None, wq.run(qprog, list(range(n)), trials=1))
        dt = time.perf_counter() - t0
    print("config3 QFT-28 + MEASURE all via run(trials=1): %.3f s, bits %s" % (dt, mem[0][:8]))
    t0 = time.perf_counter()
    res = wq.run(qprog, list(range(n)), trials=100000)
    dt = time.perf_counter() - t0
    print("config3 QFT-28 run(trials=100000) sampling fast path: %.3f s, passes %d" % (dt, wq._engine.stats()["hbm_passes"]))


if __name__ == "__main__":
    main()
