"""Build libqvm_b200.so (sm_100a only) in-tree so the shared object travels with the repo snapshot.

    python reference-qvm_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The library links the static CUDA runtime, so it has no
dependency on a particular libcudart at load time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "referenceqvm", "libqvm_b200.so")
SOURCES = ["qvm_capi.cu"]
DEPS = ["qvm_capi.cu", "qvm_kernels.cuh", "qvm_regtile.cuh", "qvm_encode.h", os.path.join("..", "..", "include", "qvm_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC,-Wall,-Wno-unused-function", "-shared", "-cudart", "static"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(OUT):
        return True
    built = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > built for d in DEPS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%d):\n%s\n%s" % (proc.returncode, " ".join(cmd), proc.stdout))
    if verbose:
        print(proc.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
