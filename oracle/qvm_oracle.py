"""CPU oracle for the gate-application hot path of rigetti/reference-qvm.

TEST INFRASTRUCTURE ONLY.  Nothing under ``reference-qvm_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and there only as the checker or the timed CPU baseline -- never as the product path.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *unmodified* reference (imported
from /root/reference through the pyquil stand-in) and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against those
fixtures (operators exactly, amplitudes to 1e-13).

Two restatements live here:

* ``ref_*`` -- the reference's own algorithm: build the full 2^n x 2^n scipy-sparse operator
  (identity (x) gate (x) identity, conjugated by a chain of adjacent-SWAP operators) and multiply
  it into the state.  This is what the reference spends its time on, so it is what
  ``cpu_baseline`` times.
* ``fast_*`` -- the closed form the operator reduces to: amplitude index is
  little-endian, gate-local index bit (k-1-j) <-> qubit a_j.  Same results (the sparse operator's
  entries are exact products of 1.0 and gate entries), used for parity at sizes where building
  2^n x 2^n operators is hopeless.

All arithmetic is complex128.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
_P0 = np.array([[1, 0], [0, 0]])
_P1 = np.array([[0, 0], [0, 1]])


# --------------------------------------------------------------------------------------------
# ref_*: the reference's sparse-operator algorithm
# --------------------------------------------------------------------------------------------
def ref_lifted_gate(i, matrix, num_qubits):
    """I_{2^(n-i-k)} (x) G (x) I_{2^i} as a sparse matrix.

    Follows reference ``unitary_generator.py:45-88`` (size check :73-77, range check :79-80,
    kron :84-88).
    """
    matrix = np.asarray(matrix)
    dim = matrix.shape[0]
    if dim & (dim - 1):
        raise TypeError("gate dimension %r is not a power of two" % (matrix.shape,))
    k = np.log2(dim)
    if not (0 <= i < num_qubits + 1 - k):
        raise ValueError("Gate index out of range!")
    below = sps.eye(2 ** i).astype(np.complex128)
    above = sps.eye(2 ** int(num_qubits - i - k)).astype(np.complex128)
    return sps.kron(above, sps.kron(matrix, below))


def ref_move_qubit(src, dst, num_qubits, positions):
    """Product of adjacent SWAPs carrying the Hilbert-space factor at ``src`` to ``dst``.

    Follows reference ``unitary_generator.py:104-150`` (two_swap_helper): returns the permutation
    operator and the updated position->qubit table.
    """
    if not (0 <= src < num_qubits and 0 <= dst < num_qubits):
        raise ValueError("Permutation SWAP index not valid")
    op = sps.eye(2 ** num_qubits).astype(np.complex128)
    table = np.array(positions, copy=True)
    step = -1 if src > dst else 1
    for pos in range(src, dst, step):
        low = pos - 1 if step < 0 else pos          # SWAP acts on (low+1, low)
        op = ref_lifted_gate(low, _SWAP, num_qubits).dot(op)
        table[low], table[low + 1] = table[low + 1], table[low]
    return op, table


def ref_permutation(args, num_qubits):
    """Permutation that makes ``args`` contiguous (in order) around their median.

    Follows reference ``unitary_generator.py:153-256``: contiguous window placed at
    ``median - len//2`` (:214-225), ping-pong sweeps of qubit moves until the window reads
    ``args`` (:236-252).  Returns (perm, start_index).
    """
    inds = np.array([int(a) for a in args])
    if inds.size == 0:
        raise ValueError("need at least one qubit")
    if np.any(inds < 0) or np.any(inds >= num_qubits):
        raise ValueError("Permutation SWAP index not valid")
    k = len(inds)
    ordered = np.sort(inds)
    start = int(ordered[k // 2]) - k // 2
    window = np.arange(start, start + k)[::-1]      # window[j] = target position of args[j]
    table = np.arange(num_qubits)                   # table[pos] = qubit currently at pos
    perm = sps.eye(2 ** num_qubits).astype(np.complex128)

    def done():
        return np.array_equal(table[window[-1]:window[0] + 1][::-1], inds)

    forward = True
    while True:
        order = range(k) if forward else range(k - 1, -1, -1)
        finished = False
        for j in order:
            here = int(np.where(table == inds[j])[0][0])
            step_op, table = ref_move_qubit(here, int(window[j]), num_qubits, table)
            perm = step_op.dot(perm)
            if done():
                finished = True
                break
        if finished:
            break
        forward = not forward
    return perm, int(window[-1])


def ref_apply_gate(matrix, args, num_qubits):
    """Full-space operator P^dagger . lifted(G) . P.  Reference ``unitary_generator.py:266-323``."""
    matrix = np.asarray(matrix)
    if int(num_qubits) != num_qubits or num_qubits < 1:
        raise ValueError("Improper number of qubits passed.")
    if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
        raise TypeError("Gate array must be two-dimensional and square matrix.")
    dim = matrix.shape[0]
    if dim & (dim - 1):
        raise TypeError("Invalid gate size. Must be power of 2!")
    k = int(np.log2(dim))
    if not (1 <= k <= num_qubits):
        raise TypeError("Invalid gate size.")
    if not isinstance(args, (list, tuple)):
        args = [args]
    perm, start = ref_permutation(args, num_qubits)
    lifted = ref_lifted_gate(start, matrix, num_qubits)
    return np.conj(perm.T).dot(lifted.dot(perm))


def ref_run_wavefunction(ops, num_qubits, rng=None):
    """Run an op list with the reference's algorithm.  Reference ``qvm_wavefunction.py:147-162,
    :354-355``.

    ``ops`` items: ("gate", matrix, qubits) | ("measure", qubit, cbit).  Returns (psi, cmem dict).
    """
    psi = np.zeros(2 ** num_qubits, dtype=np.complex128)
    psi[0] = 1.0
    cmem = {}
    for op in ops:
        if op[0] == "gate":
            psi = ref_apply_gate(op[1], list(op[2]), num_qubits).dot(psi)
        elif op[0] == "measure":
            outcome, psi = ref_measure(psi, op[1], num_qubits, rng)
            cmem[op[2]] = outcome
        else:
            raise ValueError(op[0])
    return psi, cmem


def ref_measure(psi, qubit, num_qubits, rng=None):
    """Projective measurement with collapse.  Reference ``qvm_wavefunction.py:96-129, :147-156``:
    p0 = <P0 psi|P0 psi>; ONE uniform draw; outcome 0 iff r < p0; psi <- P_b psi / sqrt(p_b)."""
    rng = np.random if rng is None else rng
    proj0 = ref_lifted_gate(qubit, _P0, num_qubits)
    kept = proj0.dot(psi)
    p0 = np.dot(np.conj(kept).T, kept)
    if rng.random_sample() < p0.real:
        op = proj0.dot(sps.eye(2 ** num_qubits) / np.sqrt(p0))
        return 0, op.dot(psi)
    proj1 = ref_lifted_gate(qubit, _P1, num_qubits)
    op = proj1.dot(sps.eye(2 ** num_qubits) / np.sqrt(1 - p0))
    return 1, op.dot(psi)


def ref_density_gate(rho, matrix, args, num_qubits):
    """rho <- U rho U^dagger with the sparse operator.  Reference ``qvm_density.py:162-166``."""
    u = sps.csc_matrix(ref_apply_gate(matrix, list(args), num_qubits))
    return u.dot(rho).dot(np.conj(u).T)


def ref_apply_kraus(rho, operators, index, num_qubits):
    """rho <- sum_k L_k rho L_k^dagger, L_k = lifted(K_k at ``index``).  Reference
    ``qvm_density.py:350-361``."""
    dim = 2 ** num_qubits
    out = sps.csc_matrix((dim, dim), dtype=complex)
    for k_op in operators:
        big = sps.csc_matrix(ref_lifted_gate(index, k_op, num_qubits))
        out = out + big.dot(rho).dot(np.conj(big.T))
    return out


# --------------------------------------------------------------------------------------------
# fast_*: the closed form (what the CUDA kernels implement)
# --------------------------------------------------------------------------------------------
def fast_apply_gate(psi, matrix, args, num_qubits):
    """psi'[x] = sum_c G[r(x), c] psi[x with bits a_j := c_{k-1-j}], r(x)_{k-1-j} = x_{a_j}.

    Closed form of reference ``unitary_generator.py:266-323``.  ``psi`` may
    carry trailing batch axes (used for density matrices / unitaries held as 2n-qubit vectors).
    """
    matrix = np.asarray(matrix, dtype=np.complex128)
    args = [int(a) for a in args]
    k = len(args)
    if matrix.shape != (2 ** k, 2 ** k):
        raise TypeError("gate shape does not match qubit count")
    if len(set(args)) != k or min(args) < 0 or max(args) >= num_qubits:
        raise ValueError("bad qubit indices %r" % (args,))
    t = np.asarray(psi, dtype=np.complex128).reshape((2,) * num_qubits)
    # tensor axis of qubit q is (n-1-q) because index bit q is little-endian
    axes = [num_qubits - 1 - a for a in args]
    g = matrix.reshape((2,) * (2 * k))               # out bits (a_0..a_{k-1}), in bits (a_0..a_{k-1})
    moved = np.tensordot(g, t, axes=(list(range(k, 2 * k)), axes))
    out = np.moveaxis(moved, list(range(k)), axes)
    return np.ascontiguousarray(out).reshape(-1)


def fast_prob_zero(psi, qubit, num_qubits):
    """p0 = sum_{x: bit q = 0} |psi[x]|^2.  Reference ``qvm_wavefunction.py:112-117``."""
    t = np.asarray(psi).reshape(2 ** (num_qubits - 1 - qubit), 2, 2 ** qubit)
    kept = t[:, 0, :]
    return float(np.sum(kept.real ** 2 + kept.imag ** 2))


def fast_collapse(psi, qubit, outcome, p_outcome, num_qubits):
    """psi <- P_b psi / sqrt(p_b).  Reference ``qvm_wavefunction.py:122, :125-126, :152``."""
    t = np.array(psi, dtype=np.complex128).reshape(2 ** (num_qubits - 1 - qubit), 2, 2 ** qubit)
    t[:, 1 - outcome, :] = 0.0
    t[:, outcome, :] *= 1.0 / np.sqrt(p_outcome)
    return t.reshape(-1)


def fast_measure(psi, qubit, num_qubits, r):
    """Measurement with a host-supplied uniform ``r``: outcome 0 iff r < p0."""
    p0 = fast_prob_zero(psi, qubit, num_qubits)
    if r < p0:
        return 0, p0, fast_collapse(psi, qubit, 0, p0, num_qubits)
    return 1, p0, fast_collapse(psi, qubit, 1, 1.0 - p0, num_qubits)


def fast_run_wavefunction(ops, num_qubits, rng=None, psi0=None):
    """Closed-form counterpart of :func:`ref_run_wavefunction` (same op list, same RNG use)."""
    rng = np.random if rng is None else rng
    if psi0 is None:
        psi = np.zeros(2 ** num_qubits, dtype=np.complex128)
        psi[0] = 1.0
    else:
        psi = np.array(psi0, dtype=np.complex128)
    cmem = {}
    for op in ops:
        if op[0] == "gate":
            psi = fast_apply_gate(psi, op[1], op[2], num_qubits)
        elif op[0] == "measure":
            outcome, _, psi = fast_measure(psi, op[1], num_qubits, rng.random_sample())
            cmem[op[2]] = outcome
        else:
            raise ValueError(op[0])
    return psi, cmem


# -- density matrices as 2n-qubit vectors --------------------------------------
def rho_to_vec(rho):
    """Row-major flatten: vector index (i << n) | j  <->  rho[i, j]."""
    rho = rho.toarray() if sps.issparse(rho) else np.asarray(rho)
    return np.ascontiguousarray(rho, dtype=np.complex128).reshape(-1)


def vec_to_rho(vec, num_qubits):
    return np.asarray(vec).reshape(2 ** num_qubits, 2 ** num_qubits)


def fast_density_gate(vec, matrix, args, num_qubits):
    """rho <- U rho U^dagger on the vectorised rho: U on vector-qubits a+n, conj(U) on a.
    Reference ``qvm_density.py:162-166``."""
    matrix = np.asarray(matrix, dtype=np.complex128)
    n2 = 2 * num_qubits
    vec = fast_apply_gate(vec, matrix, [a + num_qubits for a in args], n2)
    return fast_apply_gate(vec, np.conj(matrix), list(args), n2)


def kraus_superop(operators):
    """S = sum_k K_k (x) conj(K_k): the one (non-unitary) matrix a channel on qubits ``a`` becomes
    when applied to vector-qubits (a+n ..., a ...).  Reference ``qvm_density.py:350-361``."""
    ops = [np.asarray(k, dtype=np.complex128) for k in operators]
    return sum(np.kron(k, np.conj(k)) for k in ops)


def fast_apply_kraus(vec, operators, index, num_qubits):
    """Channel on qubit ``index`` as a 4x4 matrix on vector-qubits (index+n, index)."""
    return fast_apply_gate(vec, kraus_superop(operators), [index + num_qubits, index],
                           2 * num_qubits)


def fast_density_prob_zero(vec, qubit, num_qubits):
    """p0 = tr(P0 rho).  Reference ``qvm_density.py:127-129`` (``sparse_trace`` :51-54)."""
    diag = vec_to_rho(vec, num_qubits).diagonal()
    idx = np.arange(2 ** num_qubits)
    return float(np.sum(diag[(idx >> qubit) & 1 == 0]).real)


def fast_density_collapse(vec, qubit, outcome, p_outcome, num_qubits):
    """rho <- M rho M^dagger, M = P_b / sqrt(p_b).  Reference ``qvm_density.py:131-143, :153``."""
    rho = np.array(vec_to_rho(vec, num_qubits))
    idx = np.arange(2 ** num_qubits)
    keep = ((idx >> qubit) & 1) == outcome
    rho[~keep, :] = 0.0
    rho[:, ~keep] = 0.0
    return (rho * (1.0 / np.sqrt(p_outcome)) * (1.0 / np.sqrt(p_outcome))).reshape(-1)


# -- sampling -----------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------
# expectation values: the operator is the reference's tensor_up(PauliSum)
# --------------------------------------------------------------------------------------------
_PAULI = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
          "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.array([[1, 0], [0, -1]], dtype=complex)}


def ref_tensor_up(terms, num_qubits):
    """Dense matrix of a Pauli sum given as [(coefficient, {qubit: 'X'|'Y'|'Z'})].

    Follows reference ``unitary_generator.py:366-411`` (tensor_up): per term the left Kronecker product
    kron(P_q, acc) for q = 0 .. n-1 (qubit 0 is the least significant index bit), times the coefficient,
    summed; IndexError when a term reaches beyond the register (:393-396)."""
    dim = 2 ** num_qubits
    total = np.zeros((dim, dim), dtype=complex)
    for coefficient, ops in terms:
        if ops and max(ops) >= num_qubits:
            raise IndexError("pauli_terms has higher index than qubits")
        acc = np.array([[1.0 + 0j]])
        for q in range(num_qubits):
            acc = np.kron(_PAULI[ops.get(q, "I")], acc)
        total += coefficient * acc
    return total


def expectation_state(psi, terms, num_qubits):
    """<psi| tensor_up(terms) |psi> (the meaning of the QVM's expectation(); the reference's own methods raise
    NotImplementedError, qvm_unitary.py:99-112)."""
    return complex(np.vdot(psi, ref_tensor_up(terms, num_qubits).dot(psi)))


def fast_expectation_state(psi, terms, num_qubits):
    """The same number without the 4^n-entry operator: every Pauli factor is applied to a copy of psi with the
    closed-form gate rule (``fast_apply_gate``), then <psi|.> -- for registers where tensor_up's dense matrix
    (16 * 4^n bytes) is out of reach.  Checked against ``expectation_state`` in tests/test_oracle_golden.py."""
    total = 0.0 + 0.0j
    for coefficient, ops in terms:
        if ops and max(ops) >= num_qubits:
            raise IndexError("pauli_terms has higher index than qubits")
        phi = np.asarray(psi, dtype=np.complex128)
        for q, letter in ops.items():
            phi = fast_apply_gate(phi, _PAULI[letter], [q], num_qubits)
        total += coefficient * np.vdot(psi, phi)
    return complex(total)


def expectation_density(rho, terms, num_qubits):
    """tr(rho tensor_up(terms))."""
    return complex(np.trace(np.asarray(rho).dot(ref_tensor_up(terms, num_qubits))))


def inverse_cdf_sample(probabilities, uniforms):
    """Index draws the way numpy's legacy ``RandomState.choice(p=)`` makes them (reference
    ``qvm_density.py:445-447``): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')."""
    cdf = np.cumsum(np.asarray(probabilities, dtype=np.float64))
    cdf /= cdf[-1]
    return np.searchsorted(cdf, np.asarray(uniforms, dtype=np.float64), side="right")


def bits_big_endian(index, width):
    """``np.binary_repr`` bit list (qubit n-1 first).  Reference ``qvm_density.py:449-451``."""
    return [(int(index) >> (width - 1 - b)) & 1 for b in range(width)]


# -- circuit generators: kept importable from here for the tests; they live in /workloads.py -------
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from workloads import qft_spec, random_circuit_spec  # noqa: E402,F401
