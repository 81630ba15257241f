"""Differential tests of the CPU oracle against a pure-Python, NON-incremental
model — the shape of the reference's own differential tests
(`operator/aggregate/mod.rs:706-876`: incremental == non-incremental == linear;
`operator/distinct.rs:759-823`; `operator/join.rs:887-1017` incremental vs
stateless join): at every step the operator's output delta must equal
f(integral of the inputs so far) - f(integral before this step), where f is the
textbook definition of the operator on whole Z-sets.  The model shares no code
with the oracle (dicts of Python ints, wrapping nowhere needed at these sizes)."""
import numpy as np
import pytest

from dbsp_b200 import FoldCount, FoldSum, Max, Min, Proj, RootCircuit, Schema, key, lval, rval


def zadd(acc, rows):
    for r in rows:
        t, w = tuple(r[:-1]), r[-1]
        acc[t] = acc.get(t, 0) + w
        if acc[t] == 0:
            del acc[t]


def zdiff(new, old):
    out = {}
    for t, w in new.items():
        out[t] = w
    for t, w in old.items():
        out[t] = out.get(t, 0) - w
    return sorted((t + (w,)) for t, w in out.items() if w != 0)


def groups(z, nk):
    g = {}
    for t, w in z.items():
        g.setdefault(t[:nk], []).append((t[nk:], w))
    return g


def f_join(za, zb, nk, fn, pred):
    out = {}
    gb = groups(zb, nk)
    for ta, wa in za.items():
        k, va = ta[:nk], ta[nk:]
        for vb, wb in gb.get(k, ()):
            if pred(k, va, vb):
                o = fn(k, va, vb)
                out[o] = out.get(o, 0) + wa * wb
    return {t: w for t, w in out.items() if w != 0}


def f_aggregate(z, nk, agg):
    out = {}
    for k, vals in groups(z, nk).items():
        vs = sorted(v for v, w in vals if w != 0)
        if not vs:
            continue
        if agg == "max":
            r = vs[-1]
        elif agg == "min":
            r = vs[0]
        elif agg == "count":
            r = (len(vs),)
        elif agg == "sum":
            r = (sum(v[0] for v in vs),)
        out[k + r] = 1
    return out


def f_linear(z, nk):
    out = {}
    for k, vals in groups(z, nk).items():
        s = sum(v[0] * w for v, w in vals)
        if s != 0:
            out[k + (s,)] = 1
    return out


def f_average(z, nk):
    out = {}
    for k, vals in groups(z, nk).items():
        s, c = sum(v[0] * w for v, w in vals), sum(w for v, w in vals)
        if c != 0:
            q = abs(s) // abs(c)            # Rust `/`: truncation toward zero
            out[k + (q if (s >= 0) == (c >= 0) else -q,)] = 1
    return out


def f_distinct(z):
    return {t: 1 for t, w in z.items() if w > 0}


def rows_of(b):
    return [tuple(int(x) for x in r) for r in b.rows()]


@pytest.mark.parametrize("seed", range(6))
def test_incremental_equals_model(oracle, seed):
    be = oracle
    rng = np.random.default_rng(1000 + seed)
    sa, sb = Schema("u", "i"), Schema("u", "uu")
    c = RootCircuit(be)
    a, ha = c.add_input_indexed_zset(sa)
    b, hb = c.add_input_indexed_zset(sb)
    outs = {}
    proj = Proj(Schema("ui", "uu"), [key(0), lval(0), rval(0), rval(1)], where=[rval(1).ge(rval(0))])
    a.join(b, proj).inspect(lambda x: outs.__setitem__("join", rows_of(x)))
    a.stream_join(b, proj).inspect(lambda x: outs.__setitem__("stream_join", rows_of(x)))
    a.aggregate(Max).inspect(lambda x: outs.__setitem__("max", rows_of(x)))
    a.aggregate(Min).inspect(lambda x: outs.__setitem__("min", rows_of(x)))
    a.aggregate(FoldCount).inspect(lambda x: outs.__setitem__("count", rows_of(x)))
    b.map_index(Proj(Schema("u", "u"), [key(0), lval(1)])).aggregate(FoldSum).inspect(lambda x: outs.__setitem__("sum", rows_of(x)))
    a.aggregate_linear(lval(0)).inspect(lambda x: outs.__setitem__("linear", rows_of(x)))
    b.distinct().inspect(lambda x: outs.__setitem__("distinct", rows_of(x)))
    b.stream_distinct().inspect(lambda x: outs.__setitem__("stream_distinct", rows_of(x)))
    a.antijoin(b).inspect(lambda x: outs.__setitem__("antijoin", rows_of(x)))
    za, zb = {}, {}
    prev = {}
    for step in range(8):
        na, nb = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        ra = list(zip(rng.integers(0, 40, na).tolist(), rng.integers(-20, 20, na).tolist(), rng.integers(-2, 3, na).tolist()))
        rb = list(zip(rng.integers(0, 40, nb).tolist(), rng.integers(0, 6, nb).tolist(), rng.integers(0, 8, nb).tolist(),
                      rng.integers(-2, 3, nb).tolist()))
        ha.append(ra)
        hb.append(rb)
        c.step()
        da, db = {}, {}
        zadd(da, ra)
        zadd(db, rb)
        zadd(za, ra)
        zadd(zb, rb)
        zb_col = {}
        for t, w in zb.items():
            zb_col[(t[0], t[2])] = zb_col.get((t[0], t[2]), 0) + w
        fn = lambda k, va, vb: k + va + vb
        pred = lambda k, va, vb: vb[1] >= vb[0]
        # antijoin (join.rs:294-320) = self - self |x| distinct(other) with closure (k, v1): every distinct
        # (k, v2) row of `other` contributes once, so a key with m distinct values scales the weight by 1 - m
        nvals_b = {}
        for t in f_distinct(zb):
            nvals_b[t[0]] = nvals_b.get(t[0], 0) + 1
        cur = {
            "join": f_join(za, zb, 1, fn, pred),
            "max": f_aggregate(za, 1, "max"),
            "min": f_aggregate(za, 1, "min"),
            "count": f_aggregate(za, 1, "count"),
            "sum": f_aggregate({t: w for t, w in zb_col.items() if w != 0}, 1, "sum"),
            "linear": f_linear(za, 1),
            "distinct": f_distinct(zb),
            "antijoin": {t: w * (1 - nvals_b.get(t[0], 0)) for t, w in za.items() if nvals_b.get(t[0], 0) != 1},
        }
        for name, z in cur.items():
            assert outs[name] == zdiff(z, prev.get(name, {})), (name, step)
        prev = cur
        # stateless operators see only this step's deltas
        assert outs["stream_join"] == zdiff(f_join(da, db, 1, fn, pred), {}), ("stream_join", step)
        assert outs["stream_distinct"] == zdiff(f_distinct(db), {}), ("stream_distinct", step)


@pytest.mark.parametrize("seed", range(3))
def test_average_equals_model(oracle, seed):
    """average (aggregate/average.rs:227-307): truncating sum/count; counts kept positive
    (the reference divides by the count unconditionally)."""
    be = oracle
    rng = np.random.default_rng(2000 + seed)
    s = Schema("u", "i")
    c = RootCircuit(be)
    a, ha = c.add_input_indexed_zset(s)
    out = {}
    a.average(lval(0)).inspect(lambda x: out.__setitem__("avg", rows_of(x)))
    z, prev = {}, {}
    for step in range(8):
        n = int(rng.integers(1, 300))
        rows = list(zip(rng.integers(0, 30, n).tolist(), rng.integers(-50, 50, n).tolist(), rng.integers(1, 4, n).tolist()))
        ha.append(rows)
        c.step()
        zadd(z, rows)
        cur = f_average(z, 1)
        assert out["avg"] == zdiff(cur, prev), step
        prev = cur
