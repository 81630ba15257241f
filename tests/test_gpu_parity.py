"""CUDA path vs the CPU oracle on seeded inputs, bit-exact, through the C ABI."""
import numpy as np
import pytest

from dbsp_b200 import Max, Min, Proj, RootCircuit, Schema, Spine, capi, key, lval, rval
from parity_util import assert_batches_equal, run_nexmark_pair

pytestmark = pytest.mark.gpu


def rand_cols(rng, schema, n, domain):
    cols = []
    for t in schema.lanes:
        if t == "i":
            cols.append(rng.integers(-domain, domain, n).astype(np.int64))
        else:
            cols.append(rng.integers(0, domain, n).astype(np.uint64))
    return cols


SCHEMAS = [Schema("u"), Schema("u", "u"), Schema("i", "u"), Schema("uu", "i"), Schema("u", "uuu"), Schema("uu", "uuu"),
           Schema("uiu", "uiuu"), Schema("uuuu", "uuuu")]


@pytest.mark.parametrize("schema", SCHEMAS, ids=lambda s: f"{s.key}_{s.val}")
@pytest.mark.parametrize("n,domain", [(0, 10), (1, 10), (7, 3), (1000, 10), (5000, 1 << 20), (200_000, 1 << 12)])
def test_from_tuples(cuda, oracle, schema, n, domain):
    rng = np.random.default_rng(n * 31 + domain)
    cols = rand_cols(rng, schema, n, domain)
    w = rng.integers(-2, 3, n)
    a = cuda.batch_from_columns(schema, cols, w)
    b = oracle.batch_from_columns(schema, cols, w)
    assert_batches_equal(a, b, "from_tuples")


def test_from_tuples_wide_lanes(cuda, oracle):
    """Full 64-bit lanes: the packed key needs several radix words."""
    rng = np.random.default_rng(5)
    schema = Schema("ui", "u")
    n = 50_000
    cols = [rng.integers(0, 1 << 63, n, dtype=np.uint64) * 2 + 1, rng.integers(-(1 << 62), 1 << 62, n), rng.integers(0, 1 << 40, n).astype(np.uint64)]
    # duplicates
    for c in cols:
        c[n // 2:] = c[: n - n // 2]
    w = rng.integers(-1, 2, n)
    assert_batches_equal(cuda.batch_from_columns(schema, cols, w), oracle.batch_from_columns(schema, cols, w), "wide")


@pytest.mark.parametrize("schema", SCHEMAS, ids=lambda s: f"{s.key}_{s.val}")
@pytest.mark.parametrize("na,nb,domain", [(0, 5, 10), (5, 0, 10), (1, 1, 2), (3000, 2900, 50), (100_000, 1000, 1 << 14), (150_000, 170_000, 1 << 9)])
def test_merge(cuda, oracle, schema, na, nb, domain):
    rng = np.random.default_rng(na + 7 * nb)
    ca, cb = rand_cols(rng, schema, na, domain), rand_cols(rng, schema, nb, domain)
    wa, wb = rng.integers(-2, 3, na), rng.integers(-2, 3, nb)
    ga = cuda.merge(cuda.batch_from_columns(schema, ca, wa), cuda.batch_from_columns(schema, cb, wb))
    oa = oracle.merge(oracle.batch_from_columns(schema, ca, wa), oracle.batch_from_columns(schema, cb, wb))
    assert_batches_equal(ga, oa, "merge")


def test_merge_cancels_everything(cuda, oracle):
    rng = np.random.default_rng(3)
    s = Schema("u", "u")
    cols = rand_cols(rng, s, 20_000, 1 << 30)
    a = cuda.batch_from_columns(s, cols, np.ones(20_000, np.int64))
    m = cuda.merge(a, cuda.neg(a))
    assert len(m) == 0
    # linearity: (a + a) has doubled weights
    d = cuda.merge(a, a).download()
    assert np.all(d["diffs"] == 2)


def test_spine_union(cuda, oracle):
    rng = np.random.default_rng(11)
    s = Schema("u", "uu")
    sc, so = Spine(cuda, s), Spine(oracle, s)
    for i in range(12):
        n = int(rng.integers(1, 20_000))
        cols = rand_cols(rng, s, n, 300)
        w = rng.integers(-1, 2, n)
        sc.insert(cuda.batch_from_columns(s, cols, w))
        so.insert(oracle.batch_from_columns(s, cols, w))
    assert_batches_equal(sc.consolidate(), so.consolidate(), "spine")


def test_join_aggregate_distinct_random(cuda, oracle):
    """Random delta streams through join / aggregate(Max, Min) / distinct."""
    res = {}
    for be in (cuda, oracle):
        rng = np.random.default_rng(99)
        c = RootCircuit(be)
        a, ha = c.add_input_indexed_zset(Schema("u", "u"))
        b, hb = c.add_input_indexed_zset(Schema("u", "uu"))
        outs = []
        a.join(b, Proj(Schema("uu", "uu"), [key(0), lval(0), rval(0), rval(1)], where=[rval(1).ge(lval(0))])).inspect(outs.append)
        a.aggregate(Max).inspect(outs.append)
        a.aggregate(Min).inspect(outs.append)
        b.distinct().inspect(outs.append)
        a.average(lval(0)).inspect(outs.append)
        got = []
        for step in range(6):
            na, nb = int(rng.integers(0, 3000)), int(rng.integers(0, 3000))
            ha.append(zip(rng.integers(0, 200, na).tolist(), rng.integers(0, 50, na).tolist(), rng.integers(-2, 3, na).tolist()))
            hb.append(zip(rng.integers(0, 200, nb).tolist(), rng.integers(0, 10, nb).tolist(), rng.integers(0, 60, nb).tolist(), rng.integers(-2, 3, nb).tolist()))
            outs.clear()
            c.step()
            got.append(list(outs))
        res[be.name] = got
    for sa, sb in zip(res["cuda"], res["oracle"]):
        for x, y in zip(sa, sb):
            assert_batches_equal(x, y, "random circuit")


@pytest.mark.parametrize("query,n_events,step,rate", [("q3", 200_000, 40_000, 0), ("q4", 120_000, 40_000, 0), ("q7", 400_000, 100_000, 0),
                                                      ("q7", 1_200_000, 100_000, 10_000), ("q7", 900_000, 30_000, 5_000), ("q0", 50_000, 25_000, 0)])
def test_nexmark_parity(cuda, oracle, query, n_events, step, rate):
    total = run_nexmark_pair(cuda, oracle, query, n_events, step, rate=rate)
    if rate:   # event time advances fast enough for q7's tumbling windows to close: non-empty outputs were compared
        assert total > 0


def test_shard_partition_union(cuda, oracle):
    rng = np.random.default_rng(4)
    s = Schema("u", "u")
    cols = rand_cols(rng, s, 50_000, 1 << 16)
    b = cuda.batch_from_columns(s, cols, np.ones(50_000, np.int64))
    parts = cuda.shard_partition(b, 4)
    oparts = oracle.shard_partition(oracle.batch_from_columns(s, cols, np.ones(50_000, np.int64)), 4)
    acc = cuda.batch_empty(s)
    for p, q in zip(parts, oparts):
        assert_batches_equal(p, q, "shard part")
        acc = cuda.merge(acc, p)
    assert_batches_equal(acc, b, "union of shards")


def test_pipelined_upload_equals_from_table(cuda, oracle):
    """dbsp_upload_begin + dbsp_batch_from_upload == dbsp_batch_from_table (and the oracle)."""
    from dbsp_b200 import col
    from dbsp_b200.nexmark import NexmarkGenerator

    t = NexmarkGenerator().tables(0, 300_000)
    proj = Proj(Schema("u", "uu"), [col(0), col(2), col(3)])
    ups = [cuda.upload_begin(t["bid"], cuda.proj_table_mask(proj)) for _ in range(3)]   # several in flight
    ref = oracle.batch_from_table(t["bid"], proj)
    direct = cuda.batch_from_table(t["bid"], proj)
    assert_batches_equal(direct, ref, "from_table")
    for u in ups:
        assert_batches_equal(cuda.batch_from_upload(u, proj), ref, "from_upload")
    filt = Proj(Schema("u", "u"), [col(1), col(0)], where=[col(2).eq(10)])
    u = cuda.upload_begin(t["auction"], cuda.proj_table_mask(filt))
    assert_batches_equal(cuda.batch_from_upload(u, filt), oracle.batch_from_table(t["auction"], filt), "filtered upload")


def test_large_scale_properties(cuda):
    """Size-independent properties at a size the oracle is too slow for
    (2 x 8 M rows): sortedness, weight conservation, linearity, cancellation,
    idempotence of consolidation."""
    rng = np.random.default_rng(2024)
    s = Schema("u", "u")
    n = 8_000_000
    def mk():
        k = rng.zipf(1.3, n).astype(np.uint64) % np.uint64(n // 4)
        v = rng.integers(0, 1 << 40, n).astype(np.uint64)
        w = rng.choice(np.array([-2, -1, 1, 2], dtype=np.int64), n)
        return k, v, w
    ka, va, wa = mk()
    kb, vb, wb = mk()
    a = cuda.batch_from_columns(s, [ka, va], wa)
    b = cuda.batch_from_columns(s, [kb, vb], wb)
    m = cuda.merge(a, b)
    d = m.download()
    keys = np.repeat(d["keys"][0], np.diff(d["offs"].astype(np.int64)))
    vals = d["vals"][0]
    # strictly increasing (key, val) rows, no zero weights
    dk = np.diff(keys.astype(np.int64))
    assert np.all(dk >= 0)
    assert np.all((dk > 0) | (np.diff(vals.astype(np.int64)) > 0))
    assert np.all(d["diffs"] != 0)
    # weight conservation (wrapping i64 sums)
    assert int(d["diffs"].sum()) == int(wa.sum() + wb.sum())
    # linearity and cancellation
    da = a.download()
    dd = cuda.merge(a, a).download()
    assert np.array_equal(dd["diffs"], 2 * da["diffs"]) and np.array_equal(dd["vals"][0], da["vals"][0])
    assert len(cuda.merge(m, cuda.neg(m))) == 0
    # consolidating a consolidated batch is the identity
    flat_k = np.repeat(da["keys"][0], np.diff(da["offs"].astype(np.int64)))
    again = cuda.batch_from_columns(s, [flat_k, da["vals"][0]], da["diffs"])
    assert_batches_equal(again, a, "idempotence")
