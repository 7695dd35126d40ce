"""``QVMConnection``: the one factory users call -- unchanged signature and return types.

Reference: ``referenceqvm/api.py:11-44``.
"""
from referenceqvm.gates import gate_matrix, stabilizer_gate_matrix
from referenceqvm.qvm_wavefunction import QVM_Wavefunction
from referenceqvm.qvm_unitary import QVM_Unitary
from referenceqvm.qvm_density import QVM_Density
from referenceqvm.qvm_stabilizer import QVM_Stabilizer


def _wavefunction(gate_set, noise_model):
    return QVM_Wavefunction(gate_set=gate_set)


def _unitary(gate_set, noise_model):
    return QVM_Unitary(gate_set=gate_set)


def _density(gate_set, noise_model):
    return QVM_Density(gate_set=gate_set, noise_model=noise_model)


def _stabilizer(gate_set, noise_model):
    return QVM_Stabilizer(gate_set=stabilizer_gate_matrix)


_FACTORIES = {"wavefunction": _wavefunction, "unitary": _unitary, "density": _density,
              "stabilizer": _stabilizer}


def QVMConnection(type_trans='wavefunction', gate_set=gate_matrix, noise_model=None):
    """Build a QVM of the requested transition type.

    :param type_trans: 'wavefunction' (default), 'unitary', 'density' or 'stabilizer'
    :param gate_set: dict name -> matrix (or callable(params) -> matrix) of allowed gates
    :param noise_model: optional ``qvm_density.NoiseModel`` (density QVM only)
    """
    try:
        make = _FACTORIES[type_trans]
    except (KeyError, TypeError):
        raise TypeError("{} is not a valid QVM type.".format(type_trans))
    return make(gate_set, noise_model)
