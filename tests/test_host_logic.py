"""Host-side logic that needs no GPU: the pipelined-ingest plumbing of the
Stream layer and the CPU-baseline thread harness, driven on the oracle backend."""
import os
import sys

import numpy as np

from dbsp_b200 import RootCircuit
from dbsp_b200.nexmark import NexmarkGenerator
from dbsp_b200.nexmark import queries as nq
from parity_util import assert_batches_equal, build_query, feed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_upload_plumbing_matches_plain_feed(oracle):
    gen = NexmarkGenerator()
    c1, h1, o1 = build_query(oracle, "q4")
    c2 = RootCircuit(oracle)
    inp, h2 = nq.add_nexmark_input(c2)
    o2 = nq.q4(inp).output()
    tabs = {"person": inp.person, "auction": inp.auction, "bid": inp.bid}
    for s0 in range(0, 120_000, 40_000):
        t = gen.tables(s0, 40_000)
        feed(h1, t)
        c1.step()
        for k, ts in tabs.items():
            mask = ts.table_mask()
            if mask:
                h2[k].set_upload(oracle.upload_begin(t[k], mask))
            else:
                h2[k].set([np.empty(0, np.uint64)] * 5)
        c2.step()
        assert_batches_equal(o1.value, o2.value, f"q4 upload path step@{s0}")


def test_thread_workers_equal_single(oracle):
    """oracle/thread_workers.py (the --impl reference harness): N worker replicas == 1."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle_backend import OracleBackend
    from thread_workers import run_workers

    gen = NexmarkGenerator()
    tables = [gen.tables(s * 50_000, 50_000) for s in range(3)]
    T = 3

    def worker(rank, comm):
        be = OracleBackend()
        c = RootCircuit(be, comm)
        inp, handles = nq.add_nexmark_input(c)
        out = nq.q4(inp).gather(0).output()
        res = []
        for t in tables:
            feed(handles, {k: [col[rank::T].copy() for col in v] for k, v in t.items()})
            c.step()
            res.append(out.value.rows() if rank == 0 else None)
        return res

    got = run_workers(T, worker)[0]
    c, h, o = build_query(oracle, "q4")
    for i, t in enumerate(tables):
        feed(h, t)
        c.step()
        assert o.value.rows() == got[i]
