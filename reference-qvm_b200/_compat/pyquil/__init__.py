"""Minimal stand-in for the un-vendored third-party ``pyquil`` 1.x object model.

reference-qvm depends on ``pyquil >= 1.4.2`` (reference ``setup.py:19``) purely for
*instruction containers*: no arithmetic on the gate-application path lives in pyquil
(SURVEY.md section 8c).  pyquil is not installable in this image (no network), so this package
provides just the surface the reference, its tests and this repo's host code touch.
It is only put on ``sys.path`` when the real ``pyquil`` cannot be imported.

Semantics of ``if_then`` / ``while_do`` / ``defgate`` / ``exponentiate`` are reconstructions of
the pyquil 1.x behaviour from memory; tests relying on them are marked shim-dependent.
"""
__version__ = "1.9.0+b200shim"
IS_B200_SHIM = True
