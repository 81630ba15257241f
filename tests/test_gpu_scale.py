"""Parity at bench scale and on deep traces (CUDA vs the CPU oracle, bit-exact).

* Nexmark q3 / q4 / q7 at the bench's step size (5 M events per step, >= 10 steps): every step's output Z-set is
  compared, so the spines grow as deep as they do under bench.py.
* Traces of more than 16 batches (the per-launch limit of the probe kernels): join, aggregate (Max/Min fast path,
  general gather path) and distinct must chunk / fall back correctly (CursorList semantics, cursor_list.rs:57-127).
"""
import os

import numpy as np
import pytest

from dbsp_b200 import Proj, Schema, Spine, capi, key, lval, rval
from parity_util import assert_batches_equal, run_nexmark_pair

pytestmark = pytest.mark.gpu

SCALE_STEPS = int(os.environ.get("DBSP_SCALE_STEPS", "10"))
SCALE_EVENTS = int(os.environ.get("DBSP_SCALE_EVENTS", "5000000"))


@pytest.mark.parametrize("query,rate", [("q3", 0), ("q4", 0), ("q7", 1_000_000)])
def test_nexmark_bench_scale_parity(cuda, oracle, query, rate):
    # q7 at the bench's rate: 1 M events/s of event time, so a 10 s tumbling window closes every two steps
    total = run_nexmark_pair(cuda, oracle, query, SCALE_STEPS * SCALE_EVENTS, SCALE_EVENTS, rate=rate)
    assert total > 0


def deep_spines(cuda, oracle, schema, n_batches=19):
    """Geometrically shrinking batches (2^18, 2^17, ... rows): the 2x rule never merges them."""
    rng = np.random.default_rng(17)
    sc, so = Spine(cuda, schema), Spine(oracle, schema)
    for i in range(n_batches):
        n = 1 << (n_batches - 1 - i)
        j = np.arange(n, dtype=np.uint64)
        cols = [j % np.uint64(4096), np.full(n, i, np.uint64), j // np.uint64(4096)][: schema.nl]
        if schema.nl == 2:
            cols = [j % np.uint64(4096), np.uint64(i) * np.uint64(1 << 20) + j // np.uint64(4096)]
        w = rng.choice(np.array([-2, -1, 1, 2, 3], dtype=np.int64), n)
        sc.insert(cuda.batch_from_columns(schema, cols, w))
        so.insert(oracle.batch_from_columns(schema, cols, w))
    assert sc.stats()[1] > 16, sc.stats()
    return sc, so


def test_join_over_deep_trace(cuda, oracle):
    s = Schema("u", "uu")
    sc, so = deep_spines(cuda, oracle, s)
    rng = np.random.default_rng(5)
    n = 3000
    dcols = [rng.integers(0, 5000, n).astype(np.uint64), rng.integers(0, 4, n).astype(np.uint64)]
    dw = rng.integers(-2, 3, n)
    ds = Schema("u", "u")
    for left in (True, False):
        if left:   # delta rows feed lval (one value lane), trace rows rval (two)
            proj = Proj(Schema("uu", "uu"), [key(0), lval(0), rval(0), rval(1)], where=[rval(1).ge(lval(0))])
        else:
            proj = Proj(Schema("uu", "uu"), [key(0), rval(0), lval(0), lval(1)], where=[lval(1).ge(rval(0))])
        a = cuda.join_delta_trace(cuda.batch_from_columns(ds, dcols, dw), sc, proj, delta_is_left=left)
        b = oracle.join_delta_trace(oracle.batch_from_columns(ds, dcols, dw), so, proj, delta_is_left=left)
        assert len(a) > 0
        assert_batches_equal(a, b, f"join over {sc.stats()[1]} batches (delta_is_left={left})")


@pytest.mark.parametrize("kind", [capi.AGG_MAX, capi.AGG_MIN, capi.AGG_FOLD_COUNT, capi.AGG_FOLD_SUM])
def test_aggregate_over_deep_trace(cuda, oracle, kind):
    s = Schema("u", "u")
    sc, so = deep_spines(cuda, oracle, s)
    rng = np.random.default_rng(6)
    out_schema = s if kind in (capi.AGG_MAX, capi.AGG_MIN) else Schema("u", "u")
    oc, oo = Spine(cuda, out_schema), Spine(oracle, out_schema)
    for step in range(3):
        n = 2000
        dcols = [rng.integers(0, 5000, n).astype(np.uint64), rng.integers(0, 1 << 30, n).astype(np.uint64)]
        dw = rng.integers(-2, 3, n)
        dc, do = cuda.batch_from_columns(s, dcols, dw), oracle.batch_from_columns(s, dcols, dw)
        # in_trace already contains the delta (aggregate/mod.rs:600-684); keep the trace deep: insert only tiny deltas
        sc.insert(dc)
        so.insert(do)
        a = cuda.aggregate_delta(dc, sc, oc, kind)
        b = oracle.aggregate_delta(do, so, oo, kind)
        assert_batches_equal(a, b, f"aggregate kind {kind} step {step} over {sc.stats()[1]} batches")
        oc.insert(a)
        oo.insert(b)


def test_distinct_over_deep_trace(cuda, oracle):
    s = Schema("u", "u")
    sc, so = deep_spines(cuda, oracle, s)
    rng = np.random.default_rng(7)
    n = 4000
    # half the delta rows hit rows that exist in some batch of the trace
    i = rng.integers(0, 19, n)
    j = rng.integers(0, 1 << 10, n)
    dcols = [(j % 4096).astype(np.uint64), (i.astype(np.uint64) << np.uint64(20)) + (j // 4096).astype(np.uint64)]
    dw = rng.integers(-4, 5, n)
    a = cuda.distinct_delta(cuda.batch_from_columns(s, dcols, dw), sc)
    b = oracle.distinct_delta(oracle.batch_from_columns(s, dcols, dw), so)
    assert_batches_equal(a, b, "distinct over a deep trace")
