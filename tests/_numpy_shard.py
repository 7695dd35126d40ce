"""Numpy stand-in for ``_sharded.CudaShard`` (tests only): lets the sharded engine's HOST logic --
layout tracking, victim choice, chunked half-shard exchange, reductions, sampling ownership -- run
under gloo on CPU ranks.  The arithmetic is the oracle's closed form; the CUDA path never sees this
file."""
import numpy as np
import torch

from oracle import qvm_oracle as orc
from referenceqvm import _native as nat
from referenceqvm import _planner


def localise(op, n_local, shard_index):
    """Rewrite a physical-coordinate op for one shard: global controls / diagonal targets are
    resolved from the shard index.  Returns None when the op is inactive on this shard."""
    ctrl_local = 0
    for p in op.ctrl_positions:
        if p >= n_local:
            if not (shard_index >> (p - n_local)) & 1:
                return None
        else:
            ctrl_local |= 1 << p
    if op.kind == nat.OP_DENSE:
        assert all(t < n_local for t in op.targets)
        return nat.ClassifiedOp(op.kind, op.targets, ctrl_local, op.mat)
    k = len(op.targets)
    keep = [j for j, t in enumerate(op.targets) if t < n_local]
    diag = np.empty(1 << len(keep), dtype=np.complex128)
    for idx in range(diag.size):
        full = 0
        for j, t in enumerate(op.targets):
            if t >= n_local:
                bit = (shard_index >> (t - n_local)) & 1
            else:
                bit = (idx >> (len(keep) - 1 - keep.index(j))) & 1
            full |= bit << (k - 1 - j)
        diag[idx] = op.mat[full]
    return nat.ClassifiedOp(op.kind, tuple(op.targets[j] for j in keep), ctrl_local, diag)


def apply_local(psi, op, n_local):
    ctrls = op.ctrl_positions
    k = len(op.targets)
    if op.kind == nat.OP_DENSE:
        block = op.mat.reshape(1 << k, 1 << k)
    else:
        block = np.diag(op.mat)
    qubits = list(ctrls) + list(op.targets)
    if not qubits:
        return psi * block[0, 0]
    dim = 1 << len(qubits)
    full = np.eye(dim, dtype=np.complex128)
    full[dim - (1 << k):, dim - (1 << k):] = block
    return orc.fast_apply_gate(psi, full, qubits, n_local)


class _FakeState(object):
    def __init__(self):
        self.n_global = 0


class NumpyShard(object):
    def __init__(self, n_local, ctx):
        self.n_local = n_local
        self.ctx = ctx
        self.psi = np.zeros(1 << n_local, dtype=np.complex128)
        self.launches = 0
        self.fused_pushes = 0
        self.pushes = 0
        self.state = _FakeState()
        self._cdf = None
        self.alt = None             # second buffer (push exchange model)
        self.peers = None
        self.peer_error = None

    def close(self):
        self.psi = None

    # -- push exchange model (in-process ranks only: "pointers" are (shard, which buffer) pairs) ----------
    def alloc_alt(self):
        self.alt = np.zeros_like(self.psi)
        self._which = 0             # psi is buffer 0, alt buffer 1; swapped at every push
        return True

    def drop_alt(self):
        self.alt = None

    def close_peers(self):
        self.peers = None

    def open_peers(self, comm, with_alt):
        if not getattr(comm, "same_process", False):
            self.peer_error = "the numpy backend maps peers only between in-process ranks"
            return False
        shards = comm.allgather_object(self)
        self.peers = {r: ((shards[r], 0), (shards[r], 1)) for r in range(comm.world) if r != self.ctx.rank}
        return True

    def _buffer(self, which):
        return self.psi if which == self._which else self.alt

    def exchange_push(self, local_bits, mine_value, dst_ptrs):
        """Model of qvm_exchange_push: every amplitude is written where it will live, then the buffers
        swap roles."""
        x = np.arange(self.psi.size, dtype=np.int64)
        L = np.zeros_like(x)
        mask = 0
        mine_bits = 0
        for i, p in enumerate(local_bits):
            L |= ((x >> p) & 1) << i
            mask |= 1 << p
            mine_bits |= ((mine_value >> i) & 1) << p
        y = (x & ~mask) | mine_bits
        for value in range(1 << len(local_bits)):
            sel = L == value
            if dst_ptrs[value]:
                shard, which = dst_ptrs[value]
                target = shard._buffer(which)
                assert target is not shard.psi or shard is not self
            else:
                assert value == mine_value
                target = self.alt
            target[y[sel]] = self.psi[x[sel]]
        self.psi, self.alt = self.alt, self.psi
        self._which ^= 1
        self.pushes += 1
        self.launches += 1

    def alloc(self, n_amps):
        return torch.empty(2 * n_amps, dtype=torch.float64)

    def scalar_tensor(self, values, dtype=None):
        return torch.tensor(np.asarray(values), dtype=dtype or torch.float64)

    def apply_groups(self, ops, tile_bits, low_bits):
        n = 0
        for group in _planner.plan(ops, self.n_local, tile_bits, low_bits):
            for op in group.ops:
                lop = localise(op, self.n_local, self.ctx.rank)
                if lop is not None:
                    self.psi = apply_local(self.psi, lop, self.n_local)
            n += 1
        self.launches += n
        return n

    def apply_group(self, tile_positions, ops, bring_low, n_low, push=None):
        """One pass with in-tile relabelling, modelled: run the ops, then let every requested position
        >= n_low trade places with a low position that was not asked for.  ``push``: the pass's stores are
        the exchange (fused when no victim position lies in the tile, as in the library)."""
        new_pos = self._apply_group(tile_positions, ops, bring_low, n_low)
        if push is not None:
            bits, mine, dst = push
            self.exchange_push(bits, mine, dst)
            if not (set(bits) & set(tile_positions)):
                self.fused_pushes += 1
                self.launches -= 1
                self.pushes -= 1
        return new_pos

    def _apply_group(self, tile_positions, ops, bring_low, n_low):
        for op in ops:
            lop = localise(op, self.n_local, self.ctx.rank)
            if lop is not None:
                self.psi = apply_local(self.psi, lop, self.n_local)
        self.launches += 1
        tile = list(tile_positions)
        new_pos = list(tile)
        incoming = [p for p in bring_low if p >= n_low]
        victims = [p for p in range(n_low) if p not in bring_low and p in tile][:len(incoming)]
        if incoming and victims:
            idx = np.arange(self.psi.size, dtype=np.int64)
            for a, b in zip(incoming, victims):
                new_pos[tile.index(a)], new_pos[tile.index(b)] = b, a
                differ = ((idx >> a) ^ (idx >> b)) & 1
                idx = idx ^ (differ * ((1 << a) | (1 << b)))
            self.psi = self.psi[idx]
        return new_pos

    def measure_hist(self, high_positions, collapse=None, want_hist=True, apply=True):
        """Model of qvm_measure_hist on this shard: optional collapse (applied, or with ``apply=False`` only used
        as a filter: the state is left alone), then |amp|^2 per (high-bin, lane)."""
        idx = np.arange(self.psi.size, dtype=np.int64)
        psi = self.psi
        if collapse is not None:
            mask, value, scale = collapse
            psi = np.where((idx & mask) == value, self.psi * scale, 0.0)
            if apply:
                self.psi = psi
        self.launches += 1
        if not want_hist:
            return None
        bins = np.zeros_like(idx)
        for i, p in enumerate(high_positions):
            assert 5 <= p < self.n_local
            bins |= ((idx >> p) & 1) << i
        hist = np.zeros((1 << len(high_positions), 32))
        np.add.at(hist, (bins, idx & 31), np.abs(psi) ** 2)
        return hist

    def density_diag(self, n_rho, pos_of):
        """Model of qvm_density_diag: the diagonal entries this shard owns, through the layout."""
        i = np.arange(1 << n_rho, dtype=np.int64)
        phys = np.zeros_like(i)
        for q in range(n_rho):
            bit = (i >> q) & 1
            phys |= (bit << pos_of[q]) | (bit << pos_of[q + n_rho])
        mine = (phys >> self.n_local) == self.ctx.rank
        out = np.zeros(1 << n_rho)
        out[mine] = self.psi[phys[mine] & ((1 << self.n_local) - 1)].real
        return out

    def density_offdiag(self, n_rho, pos_of, col_xor):
        """Model of qvm_density_offdiag: rho[i, i ^ col_xor] for the rows this shard owns, through the layout."""
        i = np.arange(1 << n_rho, dtype=np.int64)
        col = i ^ col_xor
        phys = np.zeros_like(i)
        for q in range(n_rho):
            phys |= (((i >> q) & 1) << pos_of[q + n_rho]) | (((col >> q) & 1) << pos_of[q])
        mine = (phys >> self.n_local) == self.ctx.rank
        out = np.zeros(1 << n_rho, dtype=np.complex128)
        out[mine] = self.psi[phys[mine] & ((1 << self.n_local) - 1)]
        return out

    def pauli_expectation(self, xmask, zmask):
        """Model of qvm_pauli_expectation on this shard (physical masks; Z bits may be shard bits)."""
        x = np.arange(self.psi.size, dtype=np.int64)
        full = x | (self.ctx.rank << self.n_local)
        par = np.zeros_like(x)
        m = full & zmask
        while np.any(m):
            par ^= m & 1
            m >>= 1
        sign = np.where(par == 1, -1.0, 1.0)
        return complex(np.sum(np.conj(self.psi[x ^ xmask]) * sign * self.psi))

    def copy_from(self, other):
        self.psi = other.psi.copy()

    def sample_distribution(self, probabilities, uniforms):
        cdf = np.cumsum(probabilities)
        idx = np.minimum(np.searchsorted(cdf, np.asarray(uniforms) * cdf[-1], side="right"), cdf.size - 1)
        return idx.astype(np.uint64), float(cdf[-1])

    def _half_index(self, bit, value, off, cnt):
        j = np.arange(off, off + cnt, dtype=np.int64)
        lo = j & ((1 << bit) - 1)
        return ((j >> bit) << (bit + 1)) | lo | (value << bit)

    def pack_half(self, bit, value, off, cnt, buf):
        buf[:2 * cnt] = torch.from_numpy(self.psi[self._half_index(bit, value, off, cnt)].view(np.float64).copy())

    def unpack_half(self, bit, value, off, cnt, buf):
        self.psi[self._half_index(bit, value, off, cnt)] = buf[:2 * cnt].numpy().view(np.complex128)

    def reset_zero(self, set_amp0):
        self.psi[:] = 0
        if set_amp0:
            self.psi[0] = 1.0

    def set_state(self, amps):
        self.psi = np.array(amps, dtype=np.complex128)

    def get_state(self, offset=0, count=None):
        count = self.psi.size - offset if count is None else count
        return self.psi[offset:offset + count].copy()

    def norm2(self):
        return float(np.sum(np.abs(self.psi) ** 2))

    def prob_zero(self, position):
        if position >= self.n_local:
            return 0.0 if (self.ctx.rank >> (position - self.n_local)) & 1 else self.norm2()
        return orc.fast_prob_zero(self.psi, position, self.n_local)

    def collapse(self, position, outcome, scale):
        if position >= self.n_local:
            mine = (self.ctx.rank >> (position - self.n_local)) & 1
            self.psi = self.psi * (scale if mine == outcome else 0.0)
            return
        keep = ((np.arange(self.psi.size) >> position) & 1) == outcome
        self.psi = np.where(keep, self.psi * scale, 0.0)

    def probs_prepare(self):
        self._cdf = np.cumsum(np.abs(self.psi) ** 2)
        return float(self._cdf[-1])

    def sample(self, uniforms):
        target = np.asarray(uniforms) * self._cdf[-1]
        return np.minimum(np.searchsorted(self._cdf, target, side="right"), self.psi.size - 1).astype(np.uint64)

    def sync(self):
        pass

    def stats(self):
        return {"kernel_launches": self.launches}
