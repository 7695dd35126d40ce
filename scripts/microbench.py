"""Bandwidth of the simple (non-fused) kernels of libqvm_b200 on one GPU: achieved GB/s against the
algorithmic bytes each moves.  Dev tool; prints one line per kernel."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "reference-qvm_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import referenceqvm  # noqa: E402,F401
from referenceqvm import _native as nat  # noqa: E402
from referenceqvm import gates as G  # noqa: E402


def timed(st, fn, reps=5):
    fn()
    st.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    st.sync()
    return (time.perf_counter() - t0) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    st = nat.DeviceState(n)
    N = float(1 << n)
    st.reset_zero(True)
    # a non-trivial state: H on every qubit through the fused path
    for q in range(n):
        st.apply(G.H, [q])
    st.sync()
    rows = []
    rows.append(("fill_zero (16 B/amp)", timed(st, lambda: st.reset_zero(True)), 16 * N))
    for q in range(n):
        st.apply(G.H, [q])
    rows.append(("norm2 (16 B/amp read)", timed(st, st.norm2), 16 * N))
    rows.append(("prob_zero q=n-1 (8 B/amp read)", timed(st, lambda: st.prob_zero(n - 1)), 8 * N))
    rows.append(("prob_zero q=0 (16 B/amp touched)", timed(st, lambda: st.prob_zero(0)), 16 * N))
    rows.append(("collapse q=5, scale 1 (24 B/amp)", timed(st, lambda: st.collapse(5, 0, 1.0)), 24 * N))
    st.reset_zero(True)
    for q in range(n):
        st.apply(G.H, [q])
    rows.append(("probs_prepare (16 B/amp read)", timed(st, lambda: st.probs_prepare(0, 0)), 16 * N))
    u = np.random.RandomState(0).random_sample(1000000)
    st.probs_prepare(0, 0)
    t = timed(st, lambda: st.sample(u), reps=3)
    rows.append(("sample 1e6 shots (incl. H2D/D2H)", t, 1e6 * (2048 * 16 + 16)))
    rows.append(("single H gate, q=n-1 (32 B/amp)", timed(st, lambda: st.apply(G.H, [n - 1])), 32 * N))
    rows.append(("single RZ gate, q=3 (32 B/amp)", timed(st, lambda: st.apply(G.RZ(0.3), [3])), 32 * N))
    rows.append(("single CNOT 7->n-2 (32 B/amp)", timed(st, lambda: st.apply(G.CNOT, [7, n - 2])), 32 * N))
    if n % 2 == 0:
        rows.append(("dephase_all (32 B/amp)", timed(st, lambda: st.dephase_all(n // 2, (1 << (n // 2)) - 1, 0.99)), 32 * N))
    np.random.seed(0)
    rows.append(("measure q=9 (16 read + 24 r/w B/amp)", timed(st, lambda: st.measure(9, 0.3), reps=3), 40 * N))
    other = nat.DeviceState(n - 1)
    half = 1 << (n - 2)
    small = nat.DeviceState(n - 1)
    small.reset_zero(True)
    rows.append(("exchange_half same device (64 B/moved amp)",
                 timed(small, lambda: small.exchange_half(7, 1, other.device_ptr(), 0, half)), 64.0 * half))
    for name, t, b in rows:
        print("%-46s %9.3f ms  %8.1f GB/s" % (name, t * 1e3, b / t / 1e9))


if __name__ == "__main__":
    main()
