"""ctypes binding of the CPU emulator of regtile_kernel (tests/native/emu_regtile.cu): the product's
encoder and round body compiled for the host.  Built on first use with nvcc (host code only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from referenceqvm import _native as nat

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "emu_regtile.cu")
OUT = os.path.join(HERE, "native", "libqvm_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "reference-qvm_b200", "csrc")
DEPS = [SRC] + [os.path.join(CSRC, f) for f in ("qvm_regtile.cuh", "qvm_encode.h")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)
        if stale:
            nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
            cmd = [nvcc, "-std=c++17", "-O2", "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                   "-Wno-deprecated-gpu-targets", "-o", OUT, SRC]
            proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if proc.returncode != 0:
                raise RuntimeError("emulator build failed:\n" + proc.stdout)
        h = C.CDLL(OUT)
        h.qvm_emu_apply_group.restype = C.c_int
        h.qvm_emu_apply_group.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                                          C.c_uint, C.c_int, C.c_void_p]
        _lib = h
    return _lib


def apply_group(psi, n_local, tile_bits, ops, n_global=0, shard_index=0, force_full=False, reg_bits=4,
                bring_low=(), n_low=0):
    """Run one fused group on a host vector through the emulated kernel.  Returns (psi', info)
    with info = dict(rounds, need_full, status, n_ops); status 1 = encoder asked for the fallback.
    ``bring_low``: positions (members of the tile) whose qubits should leave the pass on positions
    < n_low; info["new_pos"][j] = position after the pass of the qubit that was at sorted(tile)[j]."""
    ops = [o for o in ops if o.kind != nat.OP_NOP]
    out = np.array(psi, dtype=np.complex128)
    arr = (nat.Op * max(len(ops), 1))()
    total = sum(o.mat.size for o in ops)
    pool = np.empty(max(total, 1), dtype=np.complex128)
    off = 0
    for i, o in enumerate(ops):
        a = arr[i]
        a.kind, a.k, a.ctrl_mask = o.kind, len(o.targets), o.ctrl_mask
        for j in range(nat.MAX_TARGETS):
            a.targets[j] = o.targets[j] if j < len(o.targets) else -1
        a.mat_offset, a.mat_entries = off, o.mat.size
        pool[off:off + o.mat.size] = o.mat
        off += o.mat.size
    tb = np.ascontiguousarray(sorted(tile_bits), dtype=np.int32)
    info = np.zeros(38, dtype=np.int32)
    sigma = np.zeros(16, dtype=np.int32)
    low_in = 0
    for p in bring_low:
        low_in |= 1 << list(tb).index(p)
    rc = lib().qvm_emu_apply_group(out.ctypes.data, n_local, n_global, shard_index, tb.size, tb.ctypes.data,
                                   len(ops), C.cast(arr, C.c_void_p), pool.ctypes.data, total,
                                   1 if force_full else 0, reg_bits, info.ctypes.data, low_in, n_low,
                                   sigma.ctypes.data)
    return out, {"new_pos": [int(tb[sigma[j]]) for j in range(tb.size)],
                 "rounds": int(info[0]), "need_full": bool(info[1]), "status": rc, "n_ops": int(info[3]),
                 "opc_hist": info[4:36].copy(), "layers": int(info[36]), "dirty_rounds": int(info[37])}
