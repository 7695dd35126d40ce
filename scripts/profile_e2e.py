"""Host-side profile of the end-to-end call (QVMConnection.run_and_measure, 10^6 shots) at a size where the
device time is small, to see what the fixed e2e overhead consists of."""
import cProfile
import os
import pstats
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
prog = spec_program(orc.random_circuit_spec(n, 40, 1234))
qvm = QVMConnection()
for rep in range(3):
    t0 = time.perf_counter()
    res = qvm.run_and_measure(prog, list(range(n)), trials=10 ** 6)
    print("run_and_measure n=%d: %.3f s" % (n, time.perf_counter() - t0))
cProfile.run("qvm.run_and_measure(prog, list(range(n)), trials=10 ** 6)", "/tmp/e2e.prof")
pstats.Stats("/tmp/e2e.prof").sort_stats("cumulative").print_stats(22)
