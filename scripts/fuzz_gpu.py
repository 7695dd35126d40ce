"""Randomised parity run on a GPU (not part of the test suite): random circuits with a few extra gate kinds,
random tile configurations, relabelling on, interpreted and (for some) specialised kernels, amplitudes and
MEASURE runs against the closed-form model used by the tests.  `python scripts/fuzz_gpu.py [seconds]`."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "reference-qvm_b200"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import referenceqvm  # noqa: E402,F401
from _model import apply_op  # noqa: E402
import workloads  # noqa: E402
from referenceqvm import _engine, gates as G  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    t0, cases, worst, jitted = time.time(), 0, 0.0, 0
    while time.time() - t0 < budget:
        n = int(rs.randint(12, 21))
        depth = int(rs.randint(3, 12))
        seed = int(rs.randint(1 << 30))
        tile = int(rs.choice([9, 10, 11, 12]))
        low = int(rs.choice([3, 4, 5]))
        jit = 2 if rs.randint(6) == 0 else 0
        spec = []
        for name, params, qubits in workloads.random_circuit_spec(n, depth, seed):
            spec.append((name, params, qubits))
            r = rs.randint(12)
            if r == 0:
                spec.append(("X", [], [qubits[0]]))
            elif r == 1 and len(qubits) == 2:
                spec.append(("CZ", [], list(qubits)))
            elif r == 2:
                spec.append(("T", [], [qubits[-1]]))
            elif r == 3 and len(qubits) == 2:
                spec.append(("SWAP", [], list(qubits)))
        eng = _engine.Engine(n, tile_bits=min(tile, n - 1), low_bits=low)
        eng.state.set_jit(jit)
        eng.reset()
        want = np.zeros(1 << n, dtype=np.complex128)
        want[0] = 1.0
        for name, params, qubits in spec:
            m = G.gate_matrix[name]
            m = np.asarray(m(*params) if params else m, dtype=np.complex128)
            eng.queue_matrix(m, qubits)
            want = apply_op(want, _engine.compile_gate(m, qubits), n)
        got = eng.amplitudes()
        err = float(np.max(np.abs(got - want)))
        order = [int(q) for q in rs.permutation(n)[:int(rs.randint(2, n + 1))]]
        uniforms = rs.random_sample(len(order))
        res = eng.measure_many(order, uniforms)
        probs_ok = True
        psi = want
        for (outcome, p0), q, r in zip(res, order, uniforms):
            mask = ((np.arange(1 << n) >> q) & 1) == 0
            p_want = float(np.sum(np.abs(psi[mask]) ** 2))
            o_want = 0 if r < p_want else 1
            probs_ok &= (outcome == o_want) and abs(p0 - p_want) < 1e-11
            keep = mask if o_want == 0 else ~mask
            psi = np.where(keep, psi, 0.0) / np.sqrt(p_want if o_want == 0 else 1.0 - p_want)
        err2 = float(np.max(np.abs(eng.amplitudes() - psi)))
        jitted += eng.stats()["jit_launches"]
        eng.close()
        worst = max(worst, err, err2)
        cases += 1
        if err > 1e-12 or err2 > 1e-11 or not probs_ok:
            print("FAIL n=%d depth=%d seed=%d tile=%d low=%d jit=%d err=%g err2=%g probs_ok=%s"
                  % (n, depth, seed, tile, low, jit, err, err2, probs_ok))
            sys.exit(1)
    print("fuzz ok: %d cases, worst error %.3g, %d specialised launches" % (cases, worst, jitted))


if __name__ == "__main__":
    main()
