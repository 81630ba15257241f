"""Helpers shared by the GPU parity tests, smoke() and bench.py's checker."""
import numpy as np

from dbsp_b200 import RootCircuit
from dbsp_b200.nexmark import NexmarkGenerator
from dbsp_b200.nexmark import queries as nq


def assert_batches_equal(a, b, what=""):
    """Bit-exact comparison of the canonical keys/offs/vals/diffs vectors."""
    assert a.schema == b.schema, f"{what}: schema {a.schema} != {b.schema}"
    da, db = a.download(), b.download()
    assert len(da["diffs"]) == len(db["diffs"]), f"{what}: {len(da['diffs'])} vs {len(db['diffs'])} tuples"
    for i, (x, y) in enumerate(zip(da["keys"], db["keys"])):
        assert np.array_equal(x, y), f"{what}: key lane {i} differs"
    for i, (x, y) in enumerate(zip(da["vals"], db["vals"])):
        assert np.array_equal(x, y), f"{what}: val lane {i} differs"
    if da["offs"] is not None:
        assert np.array_equal(da["offs"], db["offs"]), f"{what}: offs differ"
    assert np.array_equal(da["diffs"], db["diffs"]), f"{what}: diffs differ"


def build_query(be, query, comm=None):
    c = RootCircuit(be, comm)
    inp, handles = nq.add_nexmark_input(c)
    out = nq.QUERIES[query](inp).output()
    return c, handles, out


def feed(handles, tables):
    for k in ("person", "auction", "bid"):
        handles[k].set(tables[k])


def run_nexmark_pair(be_a, be_b, query, n_events, step, seed=0x7FC359184519C0AA, first=0, rate=0):
    """Run `query` on two backends over the same seeded events; every step's
    output Z-set must be identical.  Returns the number of output tuples.
    `rate` = first_event_rate (0 = the reference default, 10 k events per ms of event time: q7's 10 s
    windows then need > 100 M events to close; tests pass a lower rate so that windows open and close)."""
    gen = NexmarkGenerator(seed, first_event_rate=rate)
    ca, ha, oa = build_query(be_a, query)
    cb, hb, ob = build_query(be_b, query)
    total = 0
    for s0 in range(first, first + n_events, step):
        n = min(step, first + n_events - s0)
        t = gen.tables(s0, n)
        feed(ha, t)
        feed(hb, t)
        ca.step()
        cb.step()
        assert_batches_equal(oa.value, ob.value, f"{query} step@{s0}")
        total += len(oa.value)
    return total
