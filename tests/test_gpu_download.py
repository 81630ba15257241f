"""Asynchronous output reads (dbsp_batch_download_begin / dbsp_download_finish) return the same rows as the
canonical download, also when later work is queued behind them; filtered / unfiltered table projections and
key-segment probes of sizes around the tile boundaries against the oracle."""
import numpy as np
import pytest

from dbsp_b200 import Schema, Spine
from dbsp_b200.zset import Proj, col, key, lval, rval
from parity_util import assert_batches_equal

pytestmark = pytest.mark.gpu


def flat(batch):
    d = batch.download()
    s = batch.schema
    if s.nv == 0:
        return [np.asarray(k).view(np.uint64) for k in d["keys"]], np.asarray(d["diffs"])
    offs = np.asarray(d["offs"]).astype(np.int64)
    reps = np.diff(offs)
    keys = [np.repeat(np.asarray(k).view(np.uint64), reps) for k in d["keys"]]
    return keys + [np.asarray(v).view(np.uint64) for v in d["vals"]], np.asarray(d["diffs"])


@pytest.mark.parametrize("schema", [Schema("u"), Schema("u", "u"), Schema("ui", "iu")], ids=lambda s: f"{s.key}_{s.val}")
@pytest.mark.parametrize("n", [0, 1, 1000, 300_000])
def test_download_begin_matches_download(cuda, schema, n):
    rng = np.random.default_rng(n + schema.nk)
    nl = schema.nk + schema.nv
    cols = [rng.integers(0, 1 << 12, n).astype(np.uint64) for _ in range(nl)]
    b = cuda.batch_from_columns(schema, cols, rng.integers(-2, 3, n))
    m = len(b)
    dst = [np.full(m + 3, 0xDEAD, np.uint64) for _ in range(nl)]
    dw = np.full(m + 3, 77, np.int64)
    dl = cuda.download_begin(b, dst, dw)
    # queue more work behind the copy: it must not disturb it
    other = cuda.merge(b, b)
    got_cols, got_w = dl.finish()
    want_cols, want_w = flat(b)
    assert len(got_w) == m
    for g, w in zip(got_cols, want_cols):
        np.testing.assert_array_equal(g, w)
    np.testing.assert_array_equal(got_w, want_w)
    assert all(int(x[m]) == 0xDEAD for x in dst) and int(dw[m]) == 77   # nothing written past the rows
    assert len(other) <= m
    st = cuda.stats()
    assert st["host_waits"] >= 0 and st["host_wait_ms"] >= 0.0


@pytest.mark.parametrize("n", [1, 255, 1024, 1025, 4096, 70_001])
@pytest.mark.parametrize("filtered", [False, True])
def test_table_projection_sizes(cuda, oracle, n, filtered):
    """flat_map_index over a raw table: one-pass filter + projection, tile boundaries."""
    rng = np.random.default_rng(n * 2 + filtered)
    cols = [np.arange(n, dtype=np.uint64), rng.integers(0, 50, n).astype(np.uint64), rng.integers(0, 1 << 20, n).astype(np.uint64)]
    s = Schema("u", "u")
    preds = [col(1).lt(17)] if filtered else []
    proj = Proj(s, [col(0), col(2) + col(1)], preds)
    got = cuda.batch_from_table(cols, proj)
    want = oracle.batch_from_table(cols, proj)
    assert_batches_equal(got, want, f"table projection n={n} filtered={filtered}")


@pytest.mark.parametrize("nd", [1, 1023, 1024, 1025, 5000, 120_000])
def test_join_segment_sizes(cuda, oracle, nd):
    """delta x trace probe with delta sizes around the key-segment tile; duplicate keys span tiles."""
    s = Schema("u", "u")
    pj = Proj(Schema("u", "uu"), [key(0), lval(0), rval(0)], [])
    res = []
    for be in (cuda, oracle):
        rng = np.random.default_rng(nd)
        tr = Spine(be, s)
        for _ in range(3):
            m = 40_000
            tr.insert(be.batch_from_columns(s, [rng.integers(0, 3000, m).astype(np.uint64), rng.integers(0, 1 << 16, m).astype(np.uint64)], rng.integers(-1, 3, m)))
        dk = rng.integers(0, 3000, nd).astype(np.uint64) // np.uint64(1 if nd < 2000 else 7)
        d = be.batch_from_columns(s, [dk, rng.integers(0, 64, nd).astype(np.uint64)], rng.integers(1, 3, nd))
        res.append(be.join_delta_trace(d, tr, pj, True))
    assert len(res[1]) > 0
    assert_batches_equal(res[0], res[1], f"join nd={nd}")
