"""The reference's saved proptest regressions (tests/golden/proptest_regressions.json.gz,
made by tests/golden/make_regressions.py from crates/dbsp/proptest-regressions/)
run against any backend.  Expected values are the models the reference's
property tests compare with:

* consolidation (`trace/consolidation/tests/proptests.rs:16-80`): consolidate == BTreeMap sum without zeros;
* quicksort pairs (`consolidation/tests/proptests.rs`): sorted order == `sort()`;
* merge batcher (`trace/ord/merge_batcher/tests.rs:298-346`): push*/seal == aggregated map;
* distinct (`operator/distinct.rs:931-1081`): `distinct()` == `integrate().stream_distinct().differentiate()`
  at every step (the saved case is from the nested variant; each round is replayed as one flat run).
"""
import gzip
import json
import os

import numpy as np

from dbsp_b200 import RootCircuit, Schema

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proptest_regressions.json.gz")
_M = (1 << 64) - 1


def load():
    with gzip.open(_PATH, "rt") as f:
        return json.load(f)


def _wrap_i64(x):
    x &= _M
    return x - (1 << 64) if x >> 63 else x


def model(rows_lists):
    """BTreeMap<(k, v), weight> sum over all tuples, zero weights dropped (wrapping i64)."""
    acc = {}
    for rows in rows_lists:
        for k, v, w in rows:
            acc[(k, v)] = _wrap_i64(acc.get((k, v), 0) + w)
    return [(k, v, w) for (k, v), w in sorted(acc.items()) if w != 0]


def got_rows(b):
    """rows() of a (u64, u64) batch with lanes read back as unsigned."""
    return [(int(r[0]) & _M, int(r[1]) & _M, int(r[2])) for r in b.rows()]


def _cols(rows):
    a = np.array([[r[0], r[1]] for r in rows], dtype=np.uint64).reshape(-1, 2)
    return [np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])], np.array([r[2] for r in rows], dtype=np.int64)


def run_consolidation(be, case):
    rows = case["batch"]
    cols, w = _cols(rows)
    want = model([rows])
    for s in (Schema("uu"), Schema("u", "u")):
        assert got_rows(be.batch_from_columns(s, cols, w)) == want, case["source"]


def run_pairs(be, case):
    data = case["data"]
    cols, w = _cols([(a, b, 1) for a, b in data])
    assert got_rows(be.batch_from_columns(Schema("uu"), cols, w)) == model([[(a, b, 1) for a, b in data]]), case["source"]


def run_batcher_batches(be, case):
    b = be.batcher(Schema("uu"))
    for rows in case["batches"]:
        if rows:
            cols, w = _cols(rows)
            b.push_batch(cols, w)
    assert b.tuples() <= sum(len(r) for r in case["batches"])
    assert got_rows(b.seal()) == model(case["batches"]), case["source"]


def run_batcher_state(be, case):
    """A MergeSorter mid-flight: every queue entry is a list of sorted chunks
    forming one sorted run; then one more unsorted batch is pushed."""
    b = be.batcher(Schema("uu"))
    everything = []
    for lst in case["queue"]:
        run = [t for chunk in lst for t in chunk]
        everything.append(run)
        if not run:
            continue
        cols, w = _cols(run)
        keys = [(t[0], t[1]) for t in run]
        if all(x < y for x, y in zip(keys, keys[1:])) and all(t[2] != 0 for t in run):
            b.push_consolidated_batch(cols, w)
        else:
            b.push_batch(cols, w)
    if case["batch"]:
        cols, w = _cols(case["batch"])
        b.push_batch(cols, w)
        everything.append(case["batch"])
    assert got_rows(b.seal()) == model(everything), case["source"]


def run_distinct(be, case):
    s = Schema("uu")
    for rnd in case["rounds"]:
        deltas = []
        for z in rnd:
            rows = []
            for ki, k in enumerate(z["keys"]):
                for j in range(z["offs"][ki], z["offs"][ki + 1]):
                    rows.append((k, z["vals"][j], z["diffs"][j]))
            deltas.append(rows)
        c = RootCircuit(be)
        it = iter(deltas)
        inp = c.add_source(lambda: be.batch_from_rows(s, next(it)), s).index(1)
        inc = inp.distinct().output()
        noninc = inp.integrate().stream_distinct().differentiate().output()
        for step in range(len(deltas)):
            c.step()
            assert inc.value.rows() == noninc.value.rows(), (case["source"], step)


RUNNERS = {"consolidation": run_consolidation, "pairs": run_pairs, "batcher_batches": run_batcher_batches,
           "batcher_state": run_batcher_state, "distinct": run_distinct}


def all_cases():
    data = load()
    return [(kind, i) for kind in RUNNERS for i in range(len(data[kind]))]


def run(be, kind, i):
    RUNNERS[kind](be, load()[kind][i])
