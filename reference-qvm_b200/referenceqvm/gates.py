"""Gate tables handed (by value) to the device kernels.

Mirror of the data in reference ``referenceqvm/gates.py`` (standard gates :87-181, projectors
:185-189, BARENCO :193-197, ``gate_matrix`` :203-227, ``stabilizer_gate_matrix`` :229-238, Kraus
generators :241-301).  Same names, same numerical values (checked bit-for-bit against a dump of the
reference's tables in ``tests/test_gates_golden.py``); the construction is this repo's own.

Bit convention (reference ``docs/source/walkthrough.rst:8-44``): for a k-qubit gate applied as
``G a_0 ... a_{k-1}``, row/column index bit (k-1-j) belongs to qubit a_j (first argument is the
most significant bit).
"""
import cmath

import numpy as np

_SQRT1_2 = 1.0 / np.sqrt(2.0)


def _real(rows):
    return np.array(rows, dtype=float)


def _perm(images):
    """Integer permutation matrix with column c mapped to row images[c]."""
    m = np.zeros((len(images), len(images)), dtype=int)
    for col, row in enumerate(images):
        m[row, col] = 1
    return m


# ---- fixed one-qubit gates -------------------------------------------------------------------
I = _real([[1.0, 0.0], [0.0, 1.0]])
X = _real([[0.0, 1.0], [1.0, 0.0]])
Y = np.array([[0.0, 0.0 - 1.0j], [0.0 + 1.0j, 0.0]])
Z = _real([[1.0, 0.0], [0.0, -1.0]])
H = _SQRT1_2 * _real([[1.0, 1.0], [1.0, -1.0]])
S = np.array([[1.0, 0.0], [0.0, 1.0j]])
T = np.array([[1.0, 0.0], [0.0, cmath.exp(1.0j * np.pi / 4.0)]])


# ---- parametric one-qubit gates ----------------------------------------------------------------
def PHASE(phi):
    return np.array([[1.0, 0.0], [0.0, np.exp(1j * phi)]])


def RX(phi):
    c, s = np.cos(phi / 2.0), np.sin(phi / 2.0)
    return np.array([[c, -1j * s], [-1j * s, c]])


def RY(phi):
    c, s = np.cos(phi / 2.0), np.sin(phi / 2.0)
    return np.array([[c, -s], [s, c]])


def RZ(phi):
    c, s = np.cos(phi / 2.0), np.sin(phi / 2.0)
    return np.array([[c - 1j * s, 0], [0, c + 1j * s]])


# ---- two- and three-qubit gates ------------------------------------------------------------------
CZ = np.diag([1, 1, 1, -1])
CNOT = _perm([0, 1, 3, 2])
CCNOT = _perm([0, 1, 2, 3, 4, 5, 7, 6])
SWAP = _perm([0, 2, 1, 3])
CSWAP = _perm([0, 1, 2, 3, 4, 6, 5, 7])
ISWAP = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])


def _cphase_at(slot):
    def gate(phi):
        d = [1.0, 1.0, 1.0, 1.0]
        d[slot] = np.exp(1j * phi)
        return np.diag(d)
    return gate


CPHASE00 = _cphase_at(0)
CPHASE01 = _cphase_at(1)
CPHASE10 = _cphase_at(2)
CPHASE = _cphase_at(3)
for _f, _n in ((CPHASE00, "CPHASE00"), (CPHASE01, "CPHASE01"), (CPHASE10, "CPHASE10"), (CPHASE, "CPHASE")):
    _f.__name__ = _n


def PSWAP(phi):
    e = np.exp(1j * phi)
    return np.array([[1, 0, 0, 0], [0, 0, e, 0], [0, e, 0, 0], [0, 0, 0, 1]])


# ---- projectors used by MEASURE ------------------------------------------------------------------
P0 = np.array([[1, 0], [0, 0]])
P1 = np.array([[0, 0], [0, 1]])
utility_gates = {"P0": P0, "P1": P1}


def BARENCO(alpha, phi, theta):
    c, s = np.cos(theta), np.sin(theta)
    block = np.array([[np.exp(1j * phi) * c, -1j * np.exp(1j * (alpha - phi)) * s],
                      [-1j * np.exp(1j * (alpha + phi)) * s, np.exp(1j * alpha) * c]])
    return np.kron(P0, np.eye(2)) + np.kron(P1, block)


gate_matrix = {
    "I": I, "X": X, "Y": Y, "Z": Z, "H": H, "S": S, "T": T,
    "PHASE": PHASE, "RX": RX, "RY": RY, "RZ": RZ,
    "CNOT": CNOT, "CCNOT": CCNOT,
    "CPHASE00": CPHASE00, "CPHASE01": CPHASE01, "CPHASE10": CPHASE10, "CPHASE": CPHASE,
    "SWAP": SWAP, "CSWAP": CSWAP, "ISWAP": ISWAP, "PSWAP": PSWAP,
    "BARENCO": BARENCO, "CZ": CZ,
}

stabilizer_gate_matrix = {name: gate_matrix[name] for name in ("I", "X", "Y", "Z", "H", "CNOT", "S", "CZ")}


# ---- single-qubit Kraus sets (density QVM noise hooks) ------------------------------------------------
def relaxation(p):
    """Amplitude damping."""
    return [np.array([[1.0, 0.0], [0.0, np.sqrt(1 - p)]]), np.array([[0.0, np.sqrt(p)], [0.0, 0.0]])]


def dephasing(p):
    """Phase damping."""
    return [np.eye(2) * np.sqrt(1 - p / 2), np.sqrt(p / 2) * Z]


def depolarizing(p):
    w = np.sqrt(p / 3.0)
    return [np.sqrt(1.0 - p) * I, w * X, w * Y, w * Z]


def _flip(pauli):
    def kraus(p):
        return [np.sqrt(1 - p) * I, np.sqrt(p) * pauli]
    return kraus


phase_flip = _flip(Z)
bit_flip = _flip(X)
bitphase_flip = _flip(Y)

noise_gates = {"relaxation": relaxation, "dephasing": dephasing, "depolarizing": depolarizing,
               "phase_flip": phase_flip, "bit_flip": bit_flip, "bitphase_flip": bitphase_flip}
