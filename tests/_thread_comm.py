"""In-process communicator (tests only): the ranks of a sharded run as THREADS of one process, all on one
CUDA device.  It lets the real ``ShardedEngine`` -- scheduler, victim choice, push / in-place exchange
kernels over "peer" pointers, fused pass + exchange, all-gathered MEASURE runs, sharded sampling -- run on
the single-GPU box the driver's ``pytest -m gpu`` uses.  Collectives are host-side rendezvous; the
stream-ordered barrier becomes "synchronise my stream, then meet" (no kernel ever waits for another
rank's kernel, so nothing here depends on the ranks' kernels being co-resident)."""
import threading

import numpy as np


class ThreadGroup(object):
    def __init__(self, world, timeout=300.0):
        self.world = world
        self.timeout = timeout
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def comm(self, rank):
        return ThreadComm(self, rank)

    def abort(self):
        self.barrier.abort()


class ThreadComm(object):
    same_process = True

    def __init__(self, group, rank):
        self.tg = group
        self.rank, self.world = rank, group.world
        self.group = None

    def _exchange(self, value):
        tg = self.tg
        tg.slots[self.rank] = value
        tg.barrier.wait(tg.timeout)
        out = list(tg.slots)
        tg.barrier.wait(tg.timeout)          # nobody overwrites a slot before everyone has read it
        return out

    def allreduce(self, values, op="sum"):
        parts = self._exchange(np.array(values, dtype=np.float64))
        out = parts[0].copy()
        for p in parts[1:]:                  # rank order on every rank: bit-identical results
            out = np.minimum(out, p) if op == "min" else out + p
        return out

    def allreduce_int64(self, values):
        parts = self._exchange(np.array(values, dtype=np.int64))
        return np.sum(parts, axis=0)

    def allgather(self, values):
        return np.concatenate(self._exchange(np.array(values, dtype=np.float64).reshape(-1)))

    def allgather_object(self, obj):
        return self._exchange(obj)

    def allgather_bytes(self, raw):
        return self._exchange(bytes(raw))

    def broadcast(self, values, src=0):
        return self._exchange(np.array(values, dtype=np.float64))[src].copy()

    def stream_barrier(self, shard=None):
        if shard is not None:
            shard.sync()
        self.tg.barrier.wait(self.tg.timeout)


def run_ranks(world, target, timeout=600.0):
    """Run ``target(comm)`` on ``world`` threads; re-raises the first failure (the others are released
    from their barriers)."""
    tg = ThreadGroup(world)
    errors = [None] * world

    def body(rank):
        try:
            target(tg.comm(rank))
        except threading.BrokenBarrierError as exc:       # a peer failed first
            errors[rank] = errors[rank] or exc
        except BaseException as exc:                      # noqa: BLE001 - reported below
            errors[rank] = exc
            tg.abort()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    alive = [t for t in threads if t.is_alive()]
    if alive:
        tg.abort()
        raise TimeoutError("%d rank threads still running" % len(alive))
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    if any(errors):
        raise [e for e in errors if e is not None][0]
