"""Placeholder so ``referenceqvm.api`` imports like the reference (api.py:8 imports all four QVMs).

The Aaronson-Gottesman tableau simulator (reference ``qvm_stabilizer.py``, ``stabilizer_utils.py``)
is a different, polynomial-time integer algorithm that is not part of the gate-application hot
path this package rebuilds (out of scope)."""


class QVM_Stabilizer(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the stabilizer QVM is outside the scope of the B200 gate-application path; "
                                  "use type_trans='wavefunction', 'density' or 'unitary'")
