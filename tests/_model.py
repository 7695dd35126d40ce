"""Numpy model of the op semantics the tile kernel implements (tests only).

A ClassifiedOp (kind, targets, ctrl_mask, mat) means: on the amplitudes whose control bits are all
1, apply ``mat`` (dense 2^k x 2^k, or a 2^k diagonal) to the target bits, first target = MSB."""
import numpy as np

from oracle import qvm_oracle as orc
from referenceqvm import _native as nat


def op_matrix(op):
    """(full matrix incl. controls, qubit list) equivalent to the classified op."""
    ctrls = op.ctrl_positions
    k = len(op.targets)
    if op.kind == nat.OP_DENSE:
        block = op.mat.reshape(2 ** k, 2 ** k)
    elif op.kind == nat.OP_DIAG:
        block = np.diag(op.mat)
    else:
        raise ValueError("NOP has no matrix")
    dim = 2 ** (len(ctrls) + k)
    full = np.eye(dim, dtype=np.complex128)
    full[dim - 2 ** k:, dim - 2 ** k:] = block
    return full, list(ctrls) + list(op.targets)


def apply_op(psi, op, n):
    if op.kind == nat.OP_NOP:
        return psi
    full, qubits = op_matrix(op)
    if not qubits:                       # bare scalar
        return psi * full[0, 0]
    return orc.fast_apply_gate(psi, full, qubits, n)


def run_groups(psi, groups, n):
    for g in groups:
        for op in g.ops:
            assert all(t in g.tile for t in op.targets) or op.kind != nat.OP_DENSE
            psi = apply_op(psi, op, n)
    return psi


def random_state(n, seed=0):
    rs = np.random.RandomState(seed)
    v = rs.standard_normal(2 ** n) + 1j * rs.standard_normal(2 ** n)
    return v / np.linalg.norm(v)


def random_unitary(dim, rs):
    z = rs.standard_normal((dim, dim)) + 1j * rs.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))
