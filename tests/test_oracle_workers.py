"""The native C++ worker-thread harness of the CPU arm (oracle/nexmark_workers.cpp) computes, per step, the same
output Z-set as the single-threaded oracle circuit driven through the Python operator layer — for 1, 3 and 4
workers (hash-shard + in-process exchange, shard.rs:264-307: union over workers == unsharded)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import native_workers as nw  # noqa: E402
from dbsp_b200.nexmark import NexmarkGenerator  # noqa: E402
from parity_util import build_query, feed  # noqa: E402


@pytest.mark.parametrize("query,rate,step", [("q3", 0, 60_000), ("q4", 0, 40_000), ("q7", 10_000, 80_000)])
def test_native_workers_match_single_thread(oracle, query, rate, step):
    gen = NexmarkGenerator(first_event_rate=rate)
    steps = [gen.tables(s * step, step) for s in range(6)]
    c, h, out = build_query(oracle, query)
    want = []
    for t in steps:
        feed(h, t)
        c.step()
        cols, w = oracle.flat(out.value)
        want.append((len(w), nw.fingerprint(cols, w)))
    assert sum(n for n, _ in want) > 0
    for threads in (1, 3, 4):
        secs, rows, fps, _ = nw.run(query, threads, steps)
        assert list(zip(rows, fps)) == want, (query, threads)
        assert all(s > 0 for s in secs)
