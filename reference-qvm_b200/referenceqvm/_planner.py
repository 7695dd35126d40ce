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
from referenceqvm._native import OP_DENSE, OP_DIAG, OP_NOP

MAX_GROUP_OPS = 1024        # descriptor budget of one launch (shared-memory copy of the op list)
LOOKAHEAD = 8192            # ops scanned past the first unscheduled one when filling a group


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
