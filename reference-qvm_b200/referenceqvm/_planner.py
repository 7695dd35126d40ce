"""Fusion planner: packs a stream of classified gates into one-HBM-pass tile groups.

The reference applies one full-space operator per gate (qvm_wavefunction.py:158-162) -- one sweep
of the state per gate.  Here a *group* is a set of gates whose dense targets all lie inside one set
of at most ``tile_bits`` index bits; the tile kernel stages those 2^tile_bits amplitudes in shared
memory and applies the whole group in a single read + write of HBM.

Scheduling is commutation aware.  An op touches a qubit either *densely* (it is a dense target) or
*diagonally* (it is a control or a target of a diagonal op).  Two ops commute when every qubit they
share is touched diagonally by both, so e.g. RZ / CPHASE / controls never pin a tile bit and can be
hoisted into any group once the dense ops before them on the same qubits have run.
"""
import ctypes
import os

import numpy as np

from referenceqvm import _native as nat
from referenceqvm._native import OP_DENSE, OP_DIAG, OP_NOP

MAX_GROUP_OPS = 1024        # descriptor budget of one launch (shared-memory copy of the op list)
LOOKAHEAD = 8192            # ops scanned past the first unscheduled one when filling a group
#: how hard ``run_relabelled`` looks for a fuller group (``qvm_plan_choose_group``): 0 = first fit; d = also try, d
#: levels deep, banning one more of the qubits a candidate admitted.  Depth 2 takes 10-15 % of the passes off the
#: benchmark circuits (33 qubits: 28 -> 25, 30 qubits: 26 -> 23) for ~0.3 ms of host time per group; depth 4 another
#: 3-5 % (33 qubits: 24) for ~3 ms per group -- worth it only when a pass costs tens of milliseconds.  None (the
#: default, ``QVM_B200_PLAN_SEARCH`` unset) = by shard size: 4 from 2^32 amplitudes (~50 ms per pass), else 2.
SEARCH_DEPTH = None if os.environ.get("QVM_B200_PLAN_SEARCH") is None else \
    max(0, min(4, int(os.environ["QVM_B200_PLAN_SEARCH"])))
DEEP_SEARCH_FROM = 32


def search_depth_for(n_local):
    if SEARCH_DEPTH is not None:
        return SEARCH_DEPTH
    return 4 if n_local >= DEEP_SEARCH_FROM else 2


class Group(object):
    __slots__ = ("tile", "ops")

    def __init__(self, tile, ops):
        self.tile = tile        # sorted tuple of local positions
        self.ops = ops          # ClassifiedOps in execution order

    def __repr__(self):
        return "Group(tile=%r, %d ops)" % (self.tile, len(self.ops))


def _positions(mask):
    out, p = [], 0
    while mask:
        if mask & 1:
            out.append(p)
        mask >>= 1
        p += 1
    return out


def dependencies(ops):
    """deps[i] = indices of earlier ops that op i does not commute with (transitively reduced
    enough for scheduling: last dense op per qubit + diagonal ops since)."""
    last_dense = {}
    diag_since = {}
    deps = []
    for i, op in enumerate(ops):
        d = set()
        if op.kind == OP_NOP:
            deps.append(d)
            continue
        dense = op.targets if op.kind == OP_DENSE else ()
        diag = _positions(op.ctrl_mask) + (list(op.targets) if op.kind == OP_DIAG else [])
        for q in dense:
            if q in last_dense:
                d.add(last_dense[q])
            d.update(diag_since.get(q, ()))
        for q in diag:
            if q in last_dense:
                d.add(last_dense[q])
        for q in dense:
            last_dense[q] = i
            diag_since[q] = []
        for q in diag:
            diag_since.setdefault(q, []).append(i)
        deps.append(d)
    return deps


def plan(ops, n_local, tile_bits=12, low_bits=5, max_group_ops=MAX_GROUP_OPS):
    """Split ``ops`` (ClassifiedOps on physical positions; every dense target < n_local) into Groups.

    The returned groups, executed in order with each group's ops in order, are equivalent to the
    input sequence (only commuting ops are reordered)."""
    ops = [o for o in ops if o.kind != OP_NOP]
    for o in ops:
        if o.kind == OP_DENSE:
            for t in o.targets:
                if not (0 <= t < n_local):
                    raise ValueError("dense target %d is not a local position (n_local=%d)" % (t, n_local))
    k = min(n_local, max(tile_bits, max([len(o.targets) for o in ops if o.kind == OP_DENSE] or [0])))
    forced = tuple(range(min(low_bits, k)))
    if n_local <= k:
        whole = tuple(range(n_local))
        return [Group(whole, ops[i:i + max_group_ops]) for i in range(0, len(ops), max_group_ops)]

    deps = dependencies(ops)
    n = len(ops)
    done = [False] * n
    first = 0
    groups = []
    while first < n:
        while first < n and done[first]:
            first += 1
        if first >= n:
            break
        tile = set(forced)
        head = ops[first]
        if head.kind == OP_DENSE and len(tile | set(head.targets)) > k:
            # a wide gate leaves no room for all forced low bits: keep as many (lowest first) as fit
            tile = set(head.targets)
            for p in forced:
                if len(tile) < k:
                    tile.add(p)
        chosen = []
        in_group = set()
        stop = min(n, first + LOOKAHEAD)
        for i in range(first, stop):
            if done[i]:
                continue
            if any(not (done[d] or d in in_group) for d in deps[i]):
                continue
            op = ops[i]
            if op.kind == OP_DENSE:
                need = [t for t in op.targets if t not in tile]
                if len(tile) + len(need) > k:
                    continue
                tile.update(need)
            chosen.append(i)
            in_group.add(i)
            if len(chosen) >= max_group_ops:
                break
        if not chosen:       # cannot happen: the first unscheduled op always has its deps done
            raise RuntimeError("planner made no progress at op %d" % first)
        # pad the tile with the lowest unused positions (contiguity helps coalescing; costs nothing)
        p = 0
        while len(tile) < k:
            if p not in tile:
                tile.add(p)
            p += 1
        for i in chosen:
            done[i] = True
        groups.append(Group(tuple(sorted(tile)), [ops[i] for i in chosen]))
    return groups


def summarize(groups):
    n_ops = sum(len(g.ops) for g in groups)
    return {"groups": len(groups), "ops": n_ops, "ops_per_group": (n_ops / float(len(groups))) if groups else 0.0}


# ---------------------------------------------------------------------------------------------
# Planning with in-tile qubit relabelling
#
# Every tile contains the `low` lowest physical positions (256-byte coalescing), so with a fixed layout
# only k - low tile bits are free to follow the circuit and the qubits parked on the low positions run
# out of work.  `qvm_apply_group_relabel` lets a pass hand any qubits of its tile over to the low
# positions on the way out (the tile is transposed in shared memory between rounds anyway), which turns
# the constraint into: a group may use k qubits, of which at most k - low were not in the previous
# group's tile.  On the benchmark circuits that takes a quarter off the number of HBM passes.
#
# The layout (logical qubit -> physical position) therefore changes from pass to pass, and planning is
# online: the next group is chosen (in logical qubits) before the current one is launched, the current
# launch is asked to bring the qubits both groups share to the low positions, and the next group's tile
# follows from the layout the library reports back.

class Choices(object):
    """The group choices of one ``run_relabelled`` call, in the order they were made.  A walk of the planner without
    a device (``precompile.walk_pending``, run by ``Engine.flush`` ahead of a long first flush so that NVRTC can start)
    records them; the real run right behind it replays them instead of searching again -- the choices depend only on
    the ops and on the layouts the launches report back, and those are the same in both walks."""
    __slots__ = ("items", "pos", "recording")

    def __init__(self):
        self.items, self.pos, self.recording = [], 0, True

    def replay(self):
        self.pos, self.recording = 0, False
        return self

    def pick(self, compute, valid):
        if self.recording:
            out = compute()
            self.items.append(out)
            return out
        if self.pos < len(self.items):
            tile, chosen = self.items[self.pos]
            self.pos += 1
            if valid(tile, chosen):
                return set(tile), list(chosen)
            self.pos = len(self.items)   # the two walks diverged (they should not): forget the recording
        return compute()                 # (or the recording walk stopped early: carry on searching)


class _Stream(object):
    """The op stream of one ``run_relabelled`` call in the CSR form ``qvm_plan_choose_group`` reads."""

    def __init__(self, ops, deps):
        n = len(ops)
        self.n = n
        self.dense = np.array([1 if o.kind == OP_DENSE else 0 for o in ops], dtype=np.uint8)
        tgt, tgt_off, dep, dep_off = [], [0], [], [0]
        for o, d in zip(ops, deps):
            if o.kind == OP_DENSE:
                tgt.extend(o.targets)
            tgt_off.append(len(tgt))
            dep.extend(sorted(d))
            dep_off.append(len(dep))
        self.tgt = np.array(tgt, dtype=np.int32)
        self.tgt_off = np.array(tgt_off, dtype=np.int32)
        self.dep = np.array(dep, dtype=np.int32)
        self.dep_off = np.array(dep_off, dtype=np.int32)
        self.done = np.zeros(n, dtype=np.uint8)
        self.chosen = np.empty(MAX_GROUP_OPS, dtype=np.int32)
        if self.tgt.size and int(self.tgt.max()) >= 64:
            raise ValueError("the planner handles at most 64 qubits")

    def choose(self, first, k, max_new, old_qubits, seed_tile, max_group_ops, depth):
        """(tile qubits, chosen op indices): see ``qvm_plan_choose_group`` (include/qvm_b200.h)."""
        old_mask = 0
        for q in old_qubits:
            old_mask |= 1 << q
        seed_mask = 0
        for q in seed_tile:
            seed_mask |= 1 << q
        cap = min(int(max_group_ops), self.chosen.size)
        n_chosen, tile_mask = ctypes.c_int32(0), ctypes.c_uint64(0)
        nat.check(nat.lib().qvm_plan_choose_group(
            self.n, self.dense.ctypes.data, self.tgt_off.ctypes.data, self.tgt.ctypes.data, self.dep_off.ctypes.data,
            self.dep.ctypes.data, self.done.ctypes.data, int(first), LOOKAHEAD, int(k), int(max_new), old_mask, seed_mask,
            cap, int(depth), self.chosen.ctypes.data, ctypes.byref(n_chosen), ctypes.byref(tile_mask)))
        return set(_positions(tile_mask.value)), self.chosen[:n_chosen.value].tolist()


def _choose_group(ops, deps, done, first, k, max_new, is_old, seed_tile, max_group_ops):
    """Greedy group in LOGICAL qubits: ops whose dependencies are met and whose dense targets fit a tile
    of k qubits with at most ``max_new`` qubits q for which ``is_old(q)`` is false.  (The first-fit scan in
    Python: what ``qvm_plan_choose_group`` does at depth 0; kept as the reference the tests compare it with.)"""
    tile = set(seed_tile)
    new_count = sum(1 for q in tile if not is_old(q))
    chosen, in_group = [], set()
    for i in range(first, min(len(ops), first + LOOKAHEAD)):
        if done[i]:
            continue
        if any(not (done[d] or d in in_group) for d in deps[i]):
            continue
        op = ops[i]
        if op.kind == OP_DENSE:
            need = [t for t in op.targets if t not in tile]
            if need:
                nn = sum(1 for t in need if not is_old(t))
                if len(tile) + len(need) > k or new_count + nn > max_new:
                    continue
                tile.update(need)
                new_count += nn
        chosen.append(i)
        in_group.add(i)
        if len(chosen) >= max_group_ops:
            break
    return tile, chosen


def run_relabelled(ops, n_local, pos_of, launch, tile_bits=11, low_bits=4, max_group_ops=MAX_GROUP_OPS,
                   defer_tail_below=0, leftover=None, on_last=None, last_tile_bits=0, search_depth=None, choices=None):
    """Run ``ops`` (ClassifiedOps on LOGICAL qubits; every dense target currently on a local position)
    as fused groups with in-tile relabelling.  ``pos_of`` (logical -> physical, mutated in place) is the
    layout; ``launch(tile_positions, physical_ops, bring_low_positions)`` runs one pass and returns the
    new position of every tile position (same order as the sorted tile).  Returns the number of passes.

    ``defer_tail_below``: when the LAST group would hold fewer ops than this, it is not launched and the
    indices (into ``ops``) of its members are appended to the list ``leftover`` instead -- the sharded
    engine merges such a tail with the ops that become runnable after the next exchange rather than
    paying a whole pass for it.

    ``on_last(leftover_indices, tile_positions)``: called right before the LAST launch of the call (the
    thin tail, if any, already set aside); what it returns is handed to that launch as a fourth argument
    -- the sharded engine's fused exchange: the last pass stores straight into the peers' shards.

    ``choices``: a ``Choices`` object -- recording: every group choice is appended to it; replaying
    (``Choices.replay()``): the recorded choices are used instead of searching.

    ``last_tile_bits`` (with ``on_last``): a pass that carries an exchange is bound by NVLink, not by HBM or the
    FP64 pipe, so it may as well do more gate work: as soon as everything that is left fits ONE tile of
    ``last_tile_bits`` (> tile_bits) qubits, it becomes the last pass."""
    origin = [i for i, o in enumerate(ops) if o.kind != OP_NOP]
    ops = [ops[i] for i in origin]
    n = len(ops)
    if not n:
        return 0
    for o in ops:
        if o.kind == OP_DENSE:
            for t in o.targets:
                if not (0 <= pos_of[t] < n_local):
                    raise ValueError("dense target %d is not on a local position" % t)
    k = min(n_local, max(tile_bits, max([len(o.targets) for o in ops if o.kind == OP_DENSE] or [0])))
    low = min(low_bits, k)
    if n_local <= k:
        whole = list(range(n_local))
        for i in range(0, n, max_group_ops):
            launch(whole, [o.remapped(pos_of) for o in ops[i:i + max_group_ops]], None)
        return (n + max_group_ops - 1) // max_group_ops

    deps = dependencies(ops)
    stream = _Stream(ops, deps)
    done = stream.done
    depth = search_depth_for(n_local) if search_depth is None else search_depth

    def legal(tile, chosen, k_, max_new, old, seed):
        """A replayed choice must be something the search could have returned here."""
        if not set(seed) <= tile or len(tile) > k_ or sum(1 for q in tile if q not in old) > max_new:
            return False
        inside = set()
        for i in chosen:
            if done[i] or any(not (done[d] or d in inside) for d in deps[i]):
                return False
            if ops[i].kind == OP_DENSE and not set(ops[i].targets) <= tile:
                return False
            inside.add(i)
        return True

    def pick(compute, k_, max_new, old, seed):
        if choices is None:
            return compute()
        return choices.pick(compute, lambda tile, chosen: legal(tile, chosen, k_, max_new, old, seed))
    qubit_at = {p: q for q, p in enumerate(pos_of)}
    first = 0
    passes = 0
    n_done = 0

    def low_qubits():
        return [qubit_at[p] for p in range(low)]

    def fixed_choice():
        """The layout as it is: the qubits on the low positions are in, k - low more may join."""
        lq = set(low_qubits())
        return pick(lambda: stream.choose(first, k, k - low, lq, lq, max_group_ops, depth), k, k - low, lq, lq)

    k_normal = k
    big = min(int(last_tile_bits or 0), n_local - 1) if on_last is not None else 0
    cur_tile, cur_ops = fixed_choice()
    while True:
        if not cur_ops:
            raise RuntimeError("planner made no progress at op %d" % first)
        k = k_normal
        if big > k_normal:
            lq = set(low_qubits())
            big_tile, big_ops = pick(lambda: stream.choose(first, big, big - low, lq, lq, max_group_ops, 0), big, big - low, lq, lq)
            if len(big_ops) == n - n_done and len(big_ops) > len(cur_ops):
                cur_tile, cur_ops, k = big_tile, big_ops, big       # the rest of the batch in one (larger) tile
        # physical tile: the low positions, the group's qubits, then the lowest unused positions
        positions = set(range(low)) | set(pos_of[q] for q in cur_tile)
        if len(positions) > k:
            # a gate wider than k - low targets: keep as many low positions as fit (lowest first)
            positions = set(pos_of[q] for q in cur_tile)
            for p in range(low):
                if len(positions) < k:
                    positions.add(p)
        p = 0
        while len(positions) < k:
            positions.add(p)
            p += 1
        tile_positions = sorted(positions)
        tile_qubits = set(qubit_at[p] for p in tile_positions)
        for i in cur_ops:
            done[i] = 1
        while first < n and done[first]:
            first += 1
        if first >= n and leftover is not None and len(cur_ops) < defer_tail_below:
            leftover.extend(origin[i] for i in cur_ops)
            return passes
        # the next group, free to take any `low` of this tile's qubits along
        nxt_tile, nxt_ops = (set(), [])
        bring = []
        contiguous = all(tile_positions[j] == j for j in range(low))
        n_done += len(cur_ops)
        if first < n and contiguous:
            nxt_tile, nxt_ops = pick(lambda: stream.choose(first, k, k - low, tile_qubits, (), max_group_ops, depth),
                                     k, k - low, tile_qubits, ())
            if (on_last is not None and leftover is not None and len(nxt_ops) == n - n_done
                    and len(nxt_ops) < defer_tail_below):
                # everything that is left is a thin tail: it will be deferred, so THIS launch is the last one
                tail = [origin[i] for i in nxt_ops]
                push = on_last(tail, tile_positions)
                phys_ops = [ops[i].remapped(pos_of) for i in cur_ops]
                new_pos = launch(tile_positions, phys_ops, None, push) if push is not None else \
                    launch(tile_positions, phys_ops, None)
                passes += 1
                if new_pos is not None:
                    for q, np_ in [(qubit_at[p], np_) for p, np_ in zip(tile_positions, new_pos) if p != np_]:
                        pos_of[q] = np_
                        qubit_at[np_] = q
                leftover.extend(tail)
                return passes
            shared = [q for q in nxt_tile if q in tile_qubits]
            if len(shared) > low:
                # keep the qubits with the most dense work ahead on the low positions: they are in every tile
                work = {}
                for i in range(first, min(n, first + 4 * max(len(nxt_ops), 64))):
                    if not done[i] and ops[i].kind == OP_DENSE:
                        for t in ops[i].targets:
                            work[t] = work.get(t, 0) + 1
                shared.sort(key=lambda q: (-work.get(q, 0), q))
                shared = shared[:low]
            bring = [pos_of[q] for q in shared]
        phys_ops = [ops[i].remapped(pos_of) for i in cur_ops]
        push = on_last([], tile_positions) if (on_last is not None and first >= n) else None
        if push is not None:
            new_pos = launch(tile_positions, phys_ops, bring if bring else None, push)
        else:
            new_pos = launch(tile_positions, phys_ops, bring if bring else None)
        passes += 1
        if new_pos is not None:
            moved = [(qubit_at[p], np_) for p, np_ in zip(tile_positions, new_pos) if p != np_]
            for q, np_ in moved:
                pos_of[q] = np_
                qubit_at[np_] = q
        if first >= n:
            return passes
        # the planned next group stands if it fits the layout that came back
        ok = bool(nxt_ops)
        if ok:
            need = set(range(low)) | set(pos_of[q] for q in nxt_tile)
            ok = len(need) <= k
        if ok:
            cur_tile, cur_ops = nxt_tile, nxt_ops
        else:
            cur_tile, cur_ops = fixed_choice()
