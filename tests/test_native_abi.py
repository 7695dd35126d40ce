"""The C-ABI library loads on a CPU-only box and exports exactly what include/qvm_b200.h declares.
No compute call is made here (there is no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO
from referenceqvm import _native as nat
from referenceqvm import gates as G


def _header_functions():
    text = open(os.path.join(REPO, "include", "qvm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qvm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _header_functions()
    assert len(declared) >= 30
    handle = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), "libqvm_b200.so does not export %s" % name
    assert sorted(nat.EXPORTED) == declared, "ctypes binding and header disagree"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(nat.Op) == 4 + 4 + 4 * 6 + 8 + 4 + 4
    assert ctypes.sizeof(nat.Stats) == 14 * 8


def test_no_cpu_fallback_without_device():
    if nat.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(nat.NativeError) as err:
        nat.DeviceState(3)
    assert "no CPU fallback" in str(err.value)


def test_classification_of_the_standard_gate_set():
    c = nat.classify
    op = c(G.H, [3])
    assert (op.kind, op.targets, op.ctrl_mask) == (nat.OP_DENSE, (3,), 0)
    op = c(G.CNOT, [5, 2])                       # control + X
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DENSE, (2,), [5])
    assert np.array_equal(op.mat, [0, 1, 1, 0])
    op = c(G.CCNOT, [0, 3, 6])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DENSE, (6,), [0, 3])
    op = c(G.CSWAP, [4, 0, 1])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DENSE, (0, 1), [4])
    op = c(G.CZ, [1, 4])                         # two controls + scalar -1
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DIAG, (), [1, 4])
    assert op.mat.tolist() == [-1]
    op = c(G.CPHASE(0.3), [7, 1])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DIAG, (), [1, 7])
    assert op.mat[0] == np.exp(0.3j)
    op = c(G.CPHASE01(0.3), [7, 1])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DIAG, (7,), [1])
    op = c(G.CPHASE00(0.3), [7, 1])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DIAG, (7, 1), [])
    op = c(G.RZ(0.4), [2])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DIAG, (2,), [])
    assert c(G.I, [0]).kind == nat.OP_NOP
    op = c(G.BARENCO(0.1, 0.2, 0.3), [2, 9])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DENSE, (9,), [2])
    op = c(G.ISWAP, [2, 9])
    assert (op.kind, op.targets, op.ctrl_positions) == (nat.OP_DENSE, (2, 9), [])


def test_classified_ops_reproduce_the_original_matrix():
    from _model import op_matrix
    from oracle import qvm_oracle as orc
    rs = np.random.RandomState(3)
    n = 5
    vec = rs.standard_normal(32) + 1j * rs.standard_normal(32)
    cases = [(G.gate_matrix[name], q) for name, q in
             [("H", [1]), ("X", [0]), ("Y", [4]), ("Z", [2]), ("S", [3]), ("T", [1]), ("CNOT", [4, 0]),
              ("CNOT", [0, 4]), ("CCNOT", [1, 3, 0]), ("CCNOT", [4, 2, 3]), ("SWAP", [3, 1]), ("CSWAP", [2, 0, 4]),
              ("ISWAP", [0, 3]), ("CZ", [4, 1])]]
    cases += [(G.gate_matrix[name](0.37), q) for name, q in
              [("RX", [2]), ("RY", [0]), ("RZ", [4]), ("PHASE", [1]), ("CPHASE00", [3, 0]), ("CPHASE01", [0, 3]),
               ("CPHASE10", [2, 4]), ("CPHASE", [4, 2]), ("PSWAP", [1, 3])]]
    cases += [(G.BARENCO(0.3, 0.7, -1.1), [3, 0])]
    for k in (1, 2, 3):
        cases.append((rs.standard_normal((2 ** k,) * 2) + 1j * rs.standard_normal((2 ** k,) * 2), [4, 1, 2][:k]))
    for matrix, qubits in cases:
        op = nat.classify(matrix, qubits)
        want = orc.fast_apply_gate(vec, matrix, qubits, n)
        full, qs = op_matrix(op)
        got = orc.fast_apply_gate(vec, full, qs, n) if qs else vec * full[0, 0]
        assert np.array_equal(got, want) or np.max(np.abs(got - want)) < 1e-15, (qubits, op)


def test_classify_rejects_bad_shapes():
    with pytest.raises(TypeError):
        nat.classify(np.eye(3), [0])
    with pytest.raises(TypeError):
        nat.classify(np.zeros((2, 4)), [0])
    with pytest.raises(TypeError):
        nat.classify(np.eye(4), [0])
