"""B200-native gate-application path behind the ``referenceqvm`` import surface.

Same import names as rigetti/reference-qvm (``referenceqvm.api.QVMConnection`` etc.); the state
lives in HBM and every Gate / MEASURE / sampling / Kraus step runs in hand-written sm_100a kernels
reached through the C-ABI library ``libqvm_b200.so`` (see ``include/qvm_b200.h``).
"""
import os as _os
import sys as _sys

__version__ = "0.3+b200.1"

try:  # the reference's only third-party dependency besides numpy/scipy; containers only
    import pyquil as _pyquil  # noqa: F401
except ImportError:  # pragma: no cover - depends on the image
    _compat = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "_compat")
    if _compat not in _sys.path:
        _sys.path.insert(0, _compat)
    import pyquil as _pyquil  # noqa: F401
