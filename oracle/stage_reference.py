"""Stage the UNMODIFIED reference package for the CPU arm of bench.py.  TEST / BENCH INFRASTRUCTURE.

    python oracle/stage_reference.py

Copies ``/root/reference/referenceqvm/*.py`` (the package only: no tests, no docs) byte for byte into
``oracle/_ref/referenceqvm/``.  ``oracle/_ref/`` is git-ignored (reference sources never enter this
repository's history) but not gpurun-ignored, so the staged copy travels to the GPU box next to the built
``.so`` files, where ``/root/reference`` does not exist.  ``oracle/run_reference.py`` runs it in a
subprocess of its own (its package name, ``referenceqvm``, is also this repository's).
``__graft_entry__.build()`` calls this whenever ``/root/reference`` is present.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/referenceqvm"
DST = os.path.join(HERE, "_ref", "referenceqvm")


def stage(verbose=False):
    """Returns the staged directory, or None when the reference is not on this machine."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    os.makedirs(DST, exist_ok=True)
    for name in sorted(os.listdir(SRC)):
        if not name.endswith(".py"):
            continue
        src, dst = os.path.join(SRC, name), os.path.join(DST, name)
        if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", name)
    return DST


if __name__ == "__main__":
    out = stage(verbose=True)
    print(out or "reference not available")
    sys.exit(0 if out else 1)
