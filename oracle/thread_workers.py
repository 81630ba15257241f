"""CPU-baseline harness: N oracle circuit replicas on N host threads with an
in-process exchange — the reference's own execution model (Runtime::run spawns
one worker thread per replica, crates/dbsp/src/circuit/runtime.rs:137-180;
exchange through shared mailboxes, operator/communication/exchange.rs:45-64).

TEST/BASELINE INFRASTRUCTURE (lives under oracle/): used only by bench.py's
cpu_baseline / --impl reference legs and by tests.  ctypes releases the GIL
inside the oracle calls, so the workers run in parallel.
"""
from __future__ import annotations

import threading


class ThreadComm:
    """Same interface as dbsp_b200.parallel.Comm for one worker thread."""

    class Shared:
        def __init__(self, n):
            self.n = n
            self.barrier = threading.Barrier(n)
            self.box = [[None] * n for _ in range(n)]   # box[src][dst]
            self.scalars = [0] * n

    def __init__(self, shared: "ThreadComm.Shared", rank: int):
        self.sh, self.rank, self.world_size = shared, rank, shared.n
        self.bytes_sent = 0

    def barrier(self):
        self.sh.barrier.wait()

    def allreduce_max(self, x: int) -> int:
        self.sh.scalars[self.rank] = x
        self.sh.barrier.wait()
        m = max(self.sh.scalars)
        self.sh.barrier.wait()
        return m

    def _exchange(self, parts):
        for p, b in enumerate(parts):
            self.sh.box[self.rank][p] = b
        self.sh.barrier.wait()
        got = [self.sh.box[q][self.rank] for q in range(self.world_size)]
        self.sh.barrier.wait()
        return got

    @staticmethod
    def _merge_all(be, batches):
        while len(batches) > 1:
            nxt = [be.merge(batches[i], batches[i + 1]) for i in range(0, len(batches) - 1, 2)]
            if len(batches) % 2:
                nxt.append(batches[-1])
            batches = nxt
        return batches[0]

    def shard(self, be, b):
        return self._merge_all(be, self._exchange(be.shard_partition(b, self.world_size)))

    def shard_many(self, be, batches):
        # in-process mailboxes have no per-round latency worth fusing: one round per stream (exchange.rs:128-200)
        return [self.shard(be, b) for b in batches]

    def gather(self, be, b, root=0):
        empty = be.batch_empty(b.schema)
        return self._merge_all(be, self._exchange([b if p == root else empty for p in range(self.world_size)]))


def run_workers(n_workers: int, worker_fn):
    """worker_fn(rank, comm) on n_workers threads; returns their results."""
    shared = ThreadComm.Shared(n_workers)
    results, errors = [None] * n_workers, []

    def run(r):
        try:
            results[r] = worker_fn(r, ThreadComm(shared, r))
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(n_workers)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    return results
