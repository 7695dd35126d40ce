"""ctypes binding of libqvm_b200.so -- the only door from the Python host to the CUDA kernels.

Signatures follow ``include/qvm_b200.h`` one to one.  There is deliberately no fallback: if the
shared object is missing or no CUDA device is present, the first call raises ``NativeError``.
"""
import atexit
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QVM_B200_LIB") or os.path.join(_HERE, "libqvm_b200.so")   # env override: A/B builds

QVM_OK, QVM_ERR_INVALID, QVM_ERR_CUDA, QVM_ERR_NOMEM, QVM_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
OP_NOP, OP_DENSE, OP_DIAG = 0, 1, 2
MAX_TARGETS = 6
MAX_TILE_BITS = 13


class NativeError(RuntimeError):
    """The CUDA library reported a failure (or is not loadable)."""

    def __init__(self, code, message):
        super(NativeError, self).__init__("libqvm_b200 [%d]: %s" % (code, message))
        self.code = code


class Op(C.Structure):
    """``qvm_op_t``"""
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("targets", C.c_int32 * MAX_TARGETS),
                ("ctrl_mask", C.c_uint64), ("mat_offset", C.c_uint32), ("mat_entries", C.c_uint32)]


class Stats(C.Structure):
    """``qvm_stats_t``"""
    _fields_ = [(name, C.c_uint64) for name in
                ("kernel_launches", "gate_ops", "hbm_passes", "hbm_bytes", "h2d_bytes", "d2h_bytes", "regtile_rounds",
                 "jit_launches", "jit_compiles", "push_launches", "fused_push_launches", "graph_launches", "zero_fused_launches", "paced_waits")]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


_P = C.c_void_p
_SIGNATURES = {
    "qvm_version": (C.c_char_p, []),
    "qvm_last_error": (C.c_char_p, []),
    "qvm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qvm_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "qvm_create_external": (C.c_int, [C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "qvm_destroy": (C.c_int, [_P]),
    "qvm_set_shard": (C.c_int, [_P, C.c_int, C.c_uint64]),
    "qvm_set_stream": (C.c_int, [_P, _P]),
    "qvm_set_tile_bits": (C.c_int, [_P, C.c_int, C.c_int]),
    "qvm_set_kernel_path": (C.c_int, [_P, C.c_int]),
    "qvm_set_jit": (C.c_int, [_P, C.c_int]),
    "qvm_jit_wait": (C.c_int, []),
    "qvm_jit_shutdown": (C.c_int, []),
    "qvm_spec_compile_check": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int, _P, C.c_int, _P, _P, C.c_uint64, _P,
                                         C.c_uint64, C.c_int, C.c_int, _P]),
    "qvm_spec_precompile": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int, _P, C.c_int, _P, _P, C.c_uint64, C.c_int, C.c_int,
                                      _P, _P]),
    "qvm_jit_counters": (C.c_int, [_P]),
    "qvm_group_probe": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int, _P, C.c_int, _P, _P, C.c_uint64, C.c_int, C.c_int,
                                  _P, _P, _P]),
    "qvm_jit_cache_dir": (C.c_int, [C.c_char_p, C.c_uint64]),
    "qvm_jit_defer": (C.c_int, [C.c_int]),
    "qvm_trim": (C.c_int, []),
    "qvm_plan_choose_group": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                        C.c_uint64, C.c_int, C.c_int, _P, _P, _P]),
    "qvm_jit_share": (C.c_int, [C.c_int, C.c_int]),
    "qvm_debug_opcode_hist": (C.c_int, [_P, _P]),
    "qvm_n_local": (C.c_int, [_P]),
    "qvm_device_ptr": (_P, [_P]),
    "qvm_stream": (_P, [_P]),
    "qvm_sync": (C.c_int, [_P]),
    "qvm_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "qvm_stats_reset": (C.c_int, [_P]),
    "qvm_reset_zero": (C.c_int, [_P, C.c_int]),
    "qvm_set_state": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    "qvm_get_state": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    "qvm_measure_hist": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_uint64, C.c_uint64, C.c_double, _P]),
    "qvm_reset_batched": (C.c_int, [_P, C.c_int]),
    "qvm_measure_batched": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "qvm_gather_segments": (C.c_int, [_P, _P, C.c_int, C.c_uint64, _P]),
    "qvm_get_strided": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "qvm_get_state_layout": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, _P, _P]),
    "qvm_apply_group_relabel": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_uint64, C.c_int, C.c_int, _P, _P]),
    "qvm_copy_state": (C.c_int, [_P, _P]),
    "qvm_norm2": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "qvm_classify": (C.c_int, [C.c_int, _P, _P, C.POINTER(Op), _P]),
    "qvm_apply": (C.c_int, [_P, C.c_int, _P, _P]),
    "qvm_apply_group": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_uint64]),
    "qvm_dephase_all": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_double]),
    "qvm_measure": (C.c_int, [_P, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "qvm_prob_zero": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "qvm_density_prob_zero": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "qvm_collapse": (C.c_int, [_P, C.c_int, C.c_int, C.c_double]),
    "qvm_density_diag": (C.c_int, [_P, C.c_int, _P, _P]),
    "qvm_density_offdiag": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P]),
    "qvm_pauli_expectation": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    "qvm_inner_product": (C.c_int, [_P, _P, _P]),
    "qvm_probs_prepare": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "qvm_sample": (C.c_int, [_P, C.c_uint64, _P, _P]),
    "qvm_pack_half": (C.c_int, [_P, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _P]),
    "qvm_unpack_half": (C.c_int, [_P, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _P]),
    "qvm_ipc_export": (C.c_int, [_P, _P]),
    "qvm_ipc_open": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "qvm_ipc_close": (C.c_int, [_P, _P]),
    "qvm_exchange_half": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_uint64, C.c_uint64]),
    "qvm_record_begin": (C.c_int, [_P]),
    "qvm_record_end": (C.c_int, [_P, C.POINTER(_P)]),
    "qvm_plan_run": (C.c_int, [_P, _P, C.c_int]),
    "qvm_plan_launches": (C.c_int, [_P]),
    "qvm_plan_destroy": (C.c_int, [_P]),
    "qvm_alt_alloc": (C.c_int, [_P]),
    "qvm_alt_free": (C.c_int, [_P]),
    "qvm_alt_ptr": (_P, [_P]),
    "qvm_ipc_export_alt": (C.c_int, [_P, _P]),
    "qvm_exchange_push": (C.c_int, [_P, C.c_int, _P, C.c_uint32, _P]),
    "qvm_apply_group_push": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_uint64, C.c_int, C.c_int, _P, _P, C.c_int, _P,
                                       C.c_uint32, _P, C.POINTER(C.c_int)]),
    "qvm_exchange_block": (C.c_int, [_P, C.c_int, _P, C.c_uint32, C.c_uint32, _P, C.c_uint64, C.c_uint64]),
}
EXPORTED = tuple(sorted(_SIGNATURES))

_lib = None


def lib():
    """Load (once) and return the shared library with argtypes/restypes set."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(QVM_ERR_CUDA, "%s is missing: build it with `python reference-qvm_b200/build.py` "
                                            "(there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
        # background NVRTC compilations must not be in flight when the interpreter tears libraries down
        atexit.register(handle.qvm_jit_shutdown)
    return _lib


def last_error():
    return lib().qvm_last_error().decode("utf-8", "replace")


def check(code):
    if code != QVM_OK:
        raise NativeError(code, last_error())


def device_count():
    n = C.c_int(0)
    code = lib().qvm_device_count(C.byref(n))
    return n.value if code == QVM_OK else 0


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


class ClassifiedOp(object):
    """One classified gate: the ``qvm_op_t`` plus its reduced matrix / diagonal."""
    __slots__ = ("kind", "targets", "ctrl_mask", "mat")

    def __init__(self, kind, targets, ctrl_mask, mat):
        self.kind = kind
        self.targets = targets          # tuple of positions, MSB first
        self.ctrl_mask = ctrl_mask
        self.mat = mat                  # 1-D complex128 (4^k or 2^k entries)

    def remapped(self, mapping):
        """Same op with every position p replaced by mapping[p]."""
        mask, m, p = 0, self.ctrl_mask, 0
        while m:
            if m & 1:
                mask |= 1 << mapping[p]
            m >>= 1
            p += 1
        return ClassifiedOp(self.kind, tuple(mapping[t] for t in self.targets), mask, self.mat)

    @property
    def ctrl_positions(self):
        out, m, p = [], self.ctrl_mask, 0
        while m:
            if m & 1:
                out.append(p)
            m >>= 1
            p += 1
        return out

    def __repr__(self):
        return "ClassifiedOp(%s, targets=%r, ctrl=%r)" % ({0: "NOP", 1: "DENSE", 2: "DIAG"}[self.kind],
                                                         self.targets, self.ctrl_positions)


def classify(matrix, qubits):
    """``qvm_classify``: matrix (2^k x 2^k) on ``qubits`` (first = MSB) -> ClassifiedOp.

    Shape errors follow the reference's apply_gate (unitary_generator.py:284-300): TypeError."""
    m = np.asarray(matrix)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise TypeError("Gate array must be two-dimensional and square matrix.")
    dim = m.shape[0]
    if dim == 0 or dim & (dim - 1):
        raise TypeError("Invalid gate size. Must be power of 2! Received {} size".format(m.shape))
    k = dim.bit_length() - 1
    if k != len(qubits):
        raise TypeError("Invalid gate size: %d-qubit matrix applied to %d qubits" % (k, len(qubits)))
    if k < 1:
        raise TypeError("Invalid gate size. k-qubit gates supported, for k in [1, num_qubits]")
    m = _c128(m)
    q = np.ascontiguousarray(qubits, dtype=np.int32)
    op = Op()
    out = np.empty(dim * dim, dtype=np.complex128)
    code = lib().qvm_classify(k, q.ctypes.data, m.ctypes.data, C.byref(op), out.ctypes.data)
    if code == QVM_ERR_UNSUPPORTED:
        raise NotImplementedError(last_error())
    check(code)
    return ClassifiedOp(op.kind, tuple(op.targets[j] for j in range(op.k)), int(op.ctrl_mask),
                        out[:op.mat_entries].copy())


def marshal_ops(ops):
    """ClassifiedOps -> (qvm_op_t array, matrix pool, entries) as the C ABI takes them."""
    arr = (Op * max(len(ops), 1))()
    total = sum(o.mat.size for o in ops)
    pool = np.empty(max(total, 1), dtype=np.complex128)
    off = 0
    for i, o in enumerate(ops):
        a = arr[i]
        a.kind, a.k, a.ctrl_mask = o.kind, len(o.targets), o.ctrl_mask
        for j in range(MAX_TARGETS):
            a.targets[j] = o.targets[j] if j < len(o.targets) else -1
        a.mat_offset, a.mat_entries = off, o.mat.size
        pool[off:off + o.mat.size] = o.mat
        off += o.mat.size
    return arr, pool, total


def jit_wait():
    """Block until all background compilations of specialised kernels have finished."""
    check(lib().qvm_jit_wait())


def jit_counters():
    """{submitted to NVRTC, loaded from the disk cache, written to it, kernels ready} of this process."""
    out = (C.c_uint64 * 4)()
    check(lib().qvm_jit_counters(C.cast(out, _P)))
    return {"nvrtc_compiles": int(out[0]), "disk_hits": int(out[1]), "disk_stores": int(out[2]), "ready": int(out[3])}


def jit_defer(n):
    """The next ``n`` ahead-of-time submissions are compiled after everything queued the normal way."""
    check(lib().qvm_jit_defer(int(n)))


def trim():
    """Give the parked resources of destroyed handles (``qvm_destroy``) back to the driver."""
    check(lib().qvm_trim())


def jit_share(stride, offset):
    """Of this thread's following ahead-of-time submissions only every ``stride``-th (from ``offset``) is compiled
    here; the others are expected in the shared disk cache from the sibling processes."""
    check(lib().qvm_jit_share(int(stride), int(offset)))


def jit_cache_dir():
    buf = C.create_string_buffer(1024)
    check(lib().qvm_jit_cache_dir(buf, len(buf)))
    return buf.value.decode("utf-8", "replace")


def spec_precompile(n_local, tile_bits, ops, n_global=0, shard_index=0, bring_low=(), n_low=0):
    """``qvm_spec_precompile`` (no device): (status, new_pos); status 0 = cubin in the disk cache, 1 = the group
    stays with the interpreter."""
    ops = [o for o in ops if o.kind != OP_NOP]
    arr, pool, total = marshal_ops(ops)
    tb = np.ascontiguousarray(sorted(tile_bits), dtype=np.int32)
    bl = np.ascontiguousarray(list(bring_low), dtype=np.int32)
    new_pos = np.empty(tb.size, dtype=np.int32)
    code = lib().qvm_spec_precompile(n_local, n_global, shard_index, tb.size, tb.ctypes.data, len(ops), C.cast(arr, _P),
                                     pool.ctypes.data, total, n_low, bl.size, bl.ctypes.data, new_pos.ctypes.data)
    if code < 0:
        raise NativeError(code, last_error())
    return code, new_pos.tolist()


def group_probe(n_local, tile_bits, ops, n_global=0, shard_index=0, bring_low=(), n_low=0):
    """``qvm_group_probe`` (no device): (new_pos, {"rounds", "ops", "dirty_rounds", "specialisable"}) for one group."""
    ops = [o for o in ops if o.kind != OP_NOP]
    arr, pool, total = marshal_ops(ops)
    tb = np.ascontiguousarray(sorted(tile_bits), dtype=np.int32)
    bl = np.ascontiguousarray(list(bring_low), dtype=np.int32)
    new_pos = np.empty(tb.size, dtype=np.int32)
    info = np.zeros(4, dtype=np.int32)
    check(lib().qvm_group_probe(n_local, n_global, shard_index, tb.size, tb.ctypes.data, len(ops), C.cast(arr, _P),
                                pool.ctypes.data, total, n_low, bl.size, bl.ctypes.data, new_pos.ctypes.data,
                                info.ctypes.data))
    return new_pos.tolist(), {"rounds": int(info[0]), "ops": int(info[1]), "dirty_rounds": int(info[2]),
                              "specialisable": bool(info[3])}


def spec_compile_check(n_local, tile_bits, ops, n_global=0, shard_index=0, bring_low=(), n_low=0):
    """``qvm_spec_compile_check``: 0 compiled, 1 left to the interpreter; raises with the NVRTC log."""
    ops = [o for o in ops if o.kind != OP_NOP]
    arr, pool, total = marshal_ops(ops)
    tb = np.ascontiguousarray(sorted(tile_bits), dtype=np.int32)
    log = C.create_string_buffer(1 << 16)
    bl = np.ascontiguousarray(list(bring_low), dtype=np.int32)
    code = lib().qvm_spec_compile_check(n_local, n_global, shard_index, tb.size, tb.ctypes.data, len(ops),
                                        C.cast(arr, _P), pool.ctypes.data, total, log, len(log), n_low, bl.size,
                                        bl.ctypes.data)
    if code < 0:
        raise NativeError(code, log.value.decode("utf-8", "replace") or last_error())
    return code


class Plan(object):
    """Owner of one ``qvm_plan_t``: the recorded launches of a compiled program."""

    def __init__(self, pointer):
        self._p = pointer

    @property
    def launches(self):
        return int(lib().qvm_plan_launches(self._p)) if self._p else 0

    def close(self):
        if self._p:
            lib().qvm_plan_destroy(self._p)
            self._p = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass


class DeviceState(object):
    """Owner of one ``qvm_t``: a shard of 2^n_local complex128 amplitudes in HBM."""

    def __init__(self, n_local, device=0, external_ptr=None):
        self._h = _P()
        self.n_local = int(n_local)
        self.device = int(device)
        if external_ptr is None:
            code = lib().qvm_create(self.n_local, self.device, C.byref(self._h))
        else:
            code = lib().qvm_create_external(self.n_local, self.device, _P(external_ptr), C.byref(self._h))
        if code == QVM_ERR_NOMEM:
            raise MemoryError(last_error())
        check(code)
        self.n_global = 0
        self.shard_index = 0
        self.last_push_fused = False

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if self._h:
            lib().qvm_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass

    @property
    def handle(self):
        if not self._h:
            raise NativeError(QVM_ERR_INVALID, "state was closed")
        return self._h

    def __len__(self):
        return 1 << self.n_local

    def set_shard(self, n_global, shard_index):
        check(lib().qvm_set_shard(self.handle, n_global, shard_index))
        self.n_global, self.shard_index = int(n_global), int(shard_index)

    def set_stream(self, cuda_stream_ptr):
        check(lib().qvm_set_stream(self.handle, _P(cuda_stream_ptr)))

    def set_tile_bits(self, tile_bits, low_bits):
        check(lib().qvm_set_tile_bits(self.handle, tile_bits, low_bits))

    def set_jit(self, policy):
        check(lib().qvm_set_jit(self.handle, int(policy)))

    def set_kernel_path(self, use_regtile):
        check(lib().qvm_set_kernel_path(self.handle, 1 if use_regtile else 0))

    def opcode_hist(self):
        out = np.zeros(32, dtype=np.uint64)
        check(lib().qvm_debug_opcode_hist(self.handle, out.ctypes.data))
        return out

    def device_ptr(self):
        return int(lib().qvm_device_ptr(self.handle) or 0)

    def stream_ptr(self):
        return int(lib().qvm_stream(self.handle) or 0)

    def sync(self):
        check(lib().qvm_sync(self.handle))

    def stats(self):
        s = Stats()
        check(lib().qvm_stats(self.handle, C.byref(s)))
        return s.as_dict()

    def stats_reset(self):
        check(lib().qvm_stats_reset(self.handle))

    # -- state I/O -------------------------------------------------------------------------
    def reset_zero(self, set_amp0=True):
        check(lib().qvm_reset_zero(self.handle, 1 if set_amp0 else 0))

    def set_state(self, amplitudes, offset=0):
        a = _c128(np.asarray(amplitudes).reshape(-1))
        check(lib().qvm_set_state(self.handle, offset, a.size, a.ctypes.data))

    def get_state(self, offset=0, count=None, out=None):
        count = (len(self) - offset) if count is None else int(count)
        if out is None:
            out = np.empty(count, dtype=np.complex128)
        assert out.dtype == np.complex128 and out.flags.c_contiguous and out.size >= count
        check(lib().qvm_get_state(self.handle, offset, count, out.ctypes.data))
        return out

    def get_strided(self, offset, stride, count):
        out = np.empty(int(count), dtype=np.complex128)
        check(lib().qvm_get_strided(self.handle, offset, stride, count, out.ctypes.data))
        return out

    def get_layout(self, offset, stride, count, pos_of):
        """Amplitudes of the logical indices offset + j*stride under the layout ``pos_of``."""
        out = np.empty(int(count), dtype=np.complex128)
        pos = np.ascontiguousarray(pos_of, dtype=np.int32)
        code = lib().qvm_get_state_layout(self.handle, offset, stride, count, pos.size, pos.ctypes.data, out.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return out

    def copy_from(self, other):
        check(lib().qvm_copy_state(self.handle, other.handle))

    def clone(self):
        dup = DeviceState(self.n_local, self.device)
        if self.n_global:
            dup.set_shard(self.n_global, self.shard_index)
        dup.copy_from(self)
        return dup

    def norm2(self):
        v = C.c_double(0.0)
        check(lib().qvm_norm2(self.handle, C.byref(v)))
        return v.value

    # -- gates -----------------------------------------------------------------------------
    def apply(self, matrix, qubits):
        """One gate through ``qvm_apply`` (library picks the tile)."""
        m = _c128(matrix)
        q = np.ascontiguousarray(qubits, dtype=np.int32)
        code = lib().qvm_apply(self.handle, q.size, q.ctypes.data, m.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        if code == QVM_ERR_UNSUPPORTED:
            raise NotImplementedError(last_error())
        check(code)

    def apply_group(self, tile_bits, ops, bring_low=None, n_low=0, push=None):
        """A run of ClassifiedOps in one HBM pass (``qvm_apply_group``).

        With ``bring_low`` (positions of the tile) the pass also relabels qubits inside the tile
        (``qvm_apply_group_relabel``) and the new position of every tile position is returned.
        With ``push`` = (victim positions ascending, own global-bit value, destination pointers) the
        pass's stores are the global<->local exchange (``qvm_apply_group_push``); ``self.last_push_fused``
        says whether the library could fuse it."""
        ops = [o for o in ops if o.kind != OP_NOP]
        if not ops and not bring_low and push is None:
            return None if bring_low is None else sorted(tile_bits)
        arr, pool, total = marshal_ops(ops)
        tb = np.ascontiguousarray(sorted(tile_bits), dtype=np.int32)
        new_pos = None
        if push is not None:
            bl = np.ascontiguousarray(bring_low if bring_low is not None else [], dtype=np.int32)
            new_pos = np.empty(tb.size, dtype=np.int32)
            bits, mine, dst = push
            vb = np.ascontiguousarray(bits, dtype=np.int32)
            ptrs = (_P * len(dst))(*[_P(int(d)) if d else _P() for d in dst])
            fused = C.c_int(0)
            code = lib().qvm_apply_group_push(self.handle, tb.size, tb.ctypes.data, len(ops),
                                              C.cast(arr, _P) if ops else None, pool.ctypes.data, total, n_low, bl.size,
                                              bl.ctypes.data, new_pos.ctypes.data, vb.size, vb.ctypes.data, int(mine),
                                              C.cast(ptrs, _P), C.byref(fused))
            self.last_push_fused = bool(fused.value)
        elif bring_low is None:
            code = lib().qvm_apply_group(self.handle, tb.size, tb.ctypes.data, len(ops), C.cast(arr, _P),
                                         pool.ctypes.data, total)
        else:
            bl = np.ascontiguousarray(bring_low, dtype=np.int32)
            new_pos = np.empty(tb.size, dtype=np.int32)
            code = lib().qvm_apply_group_relabel(self.handle, tb.size, tb.ctypes.data, len(ops), C.cast(arr, _P),
                                                 pool.ctypes.data, total, n_low, bl.size, bl.ctypes.data,
                                                 new_pos.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        if code == QVM_ERR_UNSUPPORTED:
            raise NotImplementedError(last_error())
        check(code)
        return None if new_pos is None else new_pos.tolist()

    def dephase_all(self, n_rho, qubit_mask, factor):
        check(lib().qvm_dephase_all(self.handle, n_rho, qubit_mask, factor))

    # -- measurement / sampling ------------------------------------------------------------
    def measure(self, qubit, r):
        """Single-kernel MEASURE: returns (outcome, p0)."""
        outcome, p0 = C.c_int(0), C.c_double(0.0)
        code = lib().qvm_measure(self.handle, qubit, r, C.byref(outcome), C.byref(p0))
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return outcome.value, p0.value

    def measure_hist(self, high_positions, collapse=None, want_hist=True, apply=True):
        """One pass of a MEASURE run (``qvm_measure_hist``): optionally apply ``collapse`` = (keep_mask,
        keep_value, scale), then return hist[bin, lane] over ``high_positions`` (each >= 5, at most 5)
        and the 5 lowest index bits.  ``apply=False``: the state is left alone and the histogram is taken over
        the amplitudes consistent with ``collapse`` only (probabilities times scale^2) -- only those are read."""
        hp = np.ascontiguousarray(high_positions, dtype=np.int32)
        hist = np.empty((1 << hp.size, 32), dtype=np.float64) if want_hist else None
        mask, value, scale = collapse if collapse is not None else (0, 0, 1.0)
        mode = 0 if collapse is None else (1 if apply else 2)
        code = lib().qvm_measure_hist(self.handle, hp.size, hp.ctypes.data, mode, mask, value,
                                      float(scale), hist.ctypes.data if want_hist else None)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return hist

    def reset_batched(self, n_per_trial):
        """Every 2^n_per_trial-amplitude segment := |0..0> (one segment per trial of a batched run())."""
        check(lib().qvm_reset_batched(self.handle, n_per_trial))

    def gather_segments(self, src, n_per_trial, src_index):
        """Segment j of this state := segment ``src_index[j]`` of ``src`` (regrouping of batched trials)."""
        idx = np.ascontiguousarray(src_index, dtype=np.uint32)
        code = lib().qvm_gather_segments(self.handle, src.handle, n_per_trial, idx.size, idx.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)

    def measure_batched(self, n_per_trial, qubit, uniforms):
        """MEASURE ``qubit`` in every segment with that segment's own uniform: (outcomes, p0) arrays."""
        segs = 1 << (self.n_local - n_per_trial)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        if u.size != segs:
            raise ValueError("%d uniforms for %d trial slots" % (u.size, segs))
        outcomes = np.empty(segs, dtype=np.int32)
        p0 = np.empty(segs, dtype=np.float64)
        code = lib().qvm_measure_batched(self.handle, n_per_trial, qubit, u.ctypes.data, outcomes.ctypes.data,
                                         p0.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return outcomes, p0

    def prob_zero(self, qubit):
        v = C.c_double(0.0)
        check(lib().qvm_prob_zero(self.handle, qubit, C.byref(v)))
        return v.value

    def density_prob_zero(self, n_rho, qubit):
        v = C.c_double(0.0)
        check(lib().qvm_density_prob_zero(self.handle, n_rho, qubit, C.byref(v)))
        return v.value

    def collapse(self, qubit, outcome, scale):
        check(lib().qvm_collapse(self.handle, qubit, outcome, scale))

    def density_offdiag(self, n_rho, pos_of, col_xor):
        """rho[i, i ^ col_xor] for all rows i through the layout (complex, 0 where not owned)."""
        pos = np.ascontiguousarray(pos_of, dtype=np.int32)
        out = np.empty(1 << n_rho, dtype=np.complex128)
        code = lib().qvm_density_offdiag(self.handle, n_rho, pos.ctypes.data, int(col_xor), out.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return out

    def pauli_expectation(self, xmask, zmask):
        """This shard's sum_x conj(psi[x ^ xmask]) (-1)^popcount(x & zmask) psi[x] (physical masks)."""
        out = np.zeros(2, dtype=np.float64)
        code = lib().qvm_pauli_expectation(self.handle, int(xmask), int(zmask), out.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return complex(out[0], out[1])

    def inner_product(self, other):
        """sum_x conj(self[x]) other[x]."""
        out = np.zeros(2, dtype=np.float64)
        check(lib().qvm_inner_product(self.handle, other.handle, out.ctypes.data))
        return complex(out[0], out[1])

    def density_diag(self, n_rho, pos_of):
        """This shard's part of diag(rho) under the layout ``pos_of`` (2^n doubles, 0.0 where not owned)."""
        pos = np.ascontiguousarray(pos_of, dtype=np.int32)
        out = np.empty(1 << n_rho, dtype=np.float64)
        code = lib().qvm_density_diag(self.handle, n_rho, pos.ctypes.data, out.ctypes.data)
        if code == QVM_ERR_INVALID:
            raise ValueError(last_error())
        check(code)
        return out

    def probs_prepare(self, mode=0, n_rho=0):
        v = C.c_double(0.0)
        check(lib().qvm_probs_prepare(self.handle, mode, n_rho, C.byref(v)))
        return v.value

    def sample(self, uniforms):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.empty(u.size, dtype=np.uint64)
        check(lib().qvm_sample(self.handle, u.size, u.ctypes.data, out.ctypes.data))
        return out

    # -- exchange plumbing -------------------------------------------------------------------
    def pack_half(self, local_bit, bit_value, elem_offset, count, device_dst_ptr):
        check(lib().qvm_pack_half(self.handle, local_bit, bit_value, elem_offset, count, _P(device_dst_ptr)))

    def ipc_export(self):
        """64-byte CUDA IPC handle of this shard's allocation (bytes)."""
        buf = C.create_string_buffer(64)
        check(lib().qvm_ipc_export(self.handle, buf))
        return buf.raw

    def ipc_open(self, handle_bytes):
        """Map a peer process's shard; returns its device pointer in this process."""
        ptr = _P()
        check(lib().qvm_ipc_open(self.handle, C.c_char_p(bytes(handle_bytes)), C.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, peer_ptr):
        check(lib().qvm_ipc_close(self.handle, _P(peer_ptr)))

    def exchange_half(self, local_bit, mine_value, peer_ptr, elem_offset, count):
        check(lib().qvm_exchange_half(self.handle, local_bit, mine_value, _P(peer_ptr), elem_offset, count))

    def exchange_block(self, local_bits, mine_value, peer_value, peer_ptr, elem_offset, count):
        bits = np.ascontiguousarray(local_bits, dtype=np.int32)
        check(lib().qvm_exchange_block(self.handle, bits.size, bits.ctypes.data, mine_value, peer_value,
                                       _P(peer_ptr), elem_offset, count))

    # -- compiled programs --------------------------------------------------------------------
    def record_begin(self):
        check(lib().qvm_record_begin(self.handle))

    def record_end(self, keep=True):
        """Close the recording; returns a ``Plan`` (``keep=False``: discards it)."""
        if not keep:
            check(lib().qvm_record_end(self.handle, None))
            return None
        ptr = _P()
        check(lib().qvm_record_end(self.handle, C.byref(ptr)))
        return Plan(ptr)

    def plan_run(self, plan, reset_first=False):
        check(lib().qvm_plan_run(self.handle, plan._p, 1 if reset_first else 0))

    def alt_alloc(self):
        """Second buffer for the out-of-place exchange; False when it does not fit in HBM."""
        code = lib().qvm_alt_alloc(self.handle)
        if code in (QVM_ERR_NOMEM, QVM_ERR_UNSUPPORTED):
            return False
        check(code)
        return True

    def alt_free(self):
        check(lib().qvm_alt_free(self.handle))

    def alt_ptr(self):
        return int(lib().qvm_alt_ptr(self.handle) or 0)

    def ipc_export_alt(self):
        buf = C.create_string_buffer(64)
        check(lib().qvm_ipc_export_alt(self.handle, buf))
        return buf.raw

    def exchange_push(self, local_bits, mine_value, dst_ptrs):
        """Out-of-place exchange pass (``qvm_exchange_push``); current and alternate buffer swap roles."""
        bits = np.ascontiguousarray(local_bits, dtype=np.int32)
        ptrs = (_P * len(dst_ptrs))(*[_P(int(d)) if d else _P() for d in dst_ptrs])
        check(lib().qvm_exchange_push(self.handle, bits.size, bits.ctypes.data, int(mine_value), C.cast(ptrs, _P)))

    def unpack_half(self, local_bit, bit_value, elem_offset, count, device_src_ptr):
        check(lib().qvm_unpack_half(self.handle, local_bit, bit_value, elem_offset, count, _P(device_src_ptr)))
