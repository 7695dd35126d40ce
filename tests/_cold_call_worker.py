"""Run by tests/test_gpu_cold_call.py in a fresh process with an EMPTY cubin cache directory: the first call of
a program on a large state walks it ahead (Engine._compile_ahead), NVRTC compiles in the background, and launches
wait for a kernel that is on its way while the GPU has earlier passes queued (qvm_capi.cu launch_rec).  Whatever
mix of interpreted and specialised passes results, U then U^dagger must return |0...0>."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "reference-qvm_b200"), os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    n, depth = int(sys.argv[1]), int(sys.argv[2])
    import referenceqvm  # noqa: F401  (puts the pyquil stand-in on sys.path when pyquil is absent)
    from pyquil.quil import Program
    from pyquil.quilbase import Gate
    from referenceqvm import _native as nat
    from referenceqvm.api import QVMConnection
    from workloads import inverse_spec, random_circuit_spec

    spec = random_circuit_spec(n, depth, 97)
    prog = Program()
    for name, params, qubits in list(spec) + list(inverse_spec(spec)):
        prog.inst(Gate(name, params, qubits))
    qvm = QVMConnection()
    c0 = nat.jit_counters()
    qvm.load_program(prog)
    qvm._fresh_engine()
    # the public call: first sight of this program in this process, nothing on disk
    res = qvm.run_and_measure(prog, list(range(n)), trials=16)
    st = qvm._engine.state.stats()
    back = qvm._engine.probe(list(range(256)))
    c1 = nat.jit_counters()
    out = {"all_zero_bits": bool(all(sum(r) == 0 for r in res)), "amp0_minus_1": float(abs(back[0] - 1.0)),
           "max_rest": float(np.max(np.abs(back[1:]))), "norm2_minus_1": float(qvm._engine.norm2() - 1.0),
           "passes": int(st["hbm_passes"]), "specialised": int(st["jit_launches"]), "paced_waits": int(st["paced_waits"]),
           "submitted": int(c1["nvrtc_compiles"] - c0["nvrtc_compiles"]), "disk_hits": int(c1["disk_hits"] - c0["disk_hits"])}
    nat.jit_wait()
    # second call: every structure now has its cubin (memory or disk)
    qvm.run_and_measure(prog, list(range(n)), trials=16)
    st2 = qvm._engine.state.stats()
    out["second_call_specialised_share"] = float(st2["jit_launches"] - st["jit_launches"]) / max(1, st2["hbm_passes"] - st["hbm_passes"])
    back = qvm._engine.probe(list(range(256)))
    out["second_amp0_minus_1"] = float(abs(back[0] - 1.0))
    print("COLD " + json.dumps(out))


if __name__ == "__main__":
    main()
