"""Backend-generic differential tests in the shape of the reference's property
tests: the incremental operator must equal the non-incremental one applied to
the integral, differentiated (`operator/aggregate/mod.rs:706-876`,
`operator/distinct.rs:931-1005`, `operator/join.rs:136-155`)."""
import numpy as np

from dbsp_b200 import FoldCount, FoldSum, Max, Min, Proj, RootCircuit, Schema, key, lval, rval


def rows_of(b):
    return [tuple(int(x) for x in r) for r in b.rows()]


def _deltas(rng, steps, max_tuples, nkeys, vlo, vhi):
    out = []
    for _ in range(steps):
        n = int(rng.integers(0, max_tuples))
        out.append(list(zip(rng.integers(0, nkeys, n).tolist(), rng.integers(vlo, vhi, n).tolist(), rng.integers(-1, 2, n).tolist())))
    return out


def run_aggregate_differential(be, seed=0, steps=15, max_tuples=10, nkeys=5, maxval=3):
    """aggregate(A) == integrate().stream_aggregate(A).differentiate() for Max, Min and the two
    Folds; test_zset of aggregate/mod.rs:831-851: keys 0..5, values -3..3, weights -1..=1, <= 10 tuples."""
    rng = np.random.default_rng(seed)
    deltas = _deltas(rng, steps, max_tuples, nkeys, -maxval, maxval)
    s = Schema("u", "i")
    c = RootCircuit(be)
    it = iter(deltas)
    inp = c.add_source(lambda: be.batch_from_rows(s, next(it)), s).index(1)
    pairs = []
    for agg in (Max, Min, FoldCount, FoldSum):
        inc = inp.aggregate(agg).gather(0).output()
        noninc = inp.integrate().stream_aggregate(agg).differentiate().gather(0).output()
        pairs.append((agg.__name__, inc, noninc))
    for step in range(steps):
        c.step()
        for name, inc, noninc in pairs:
            assert rows_of(inc.value) == rows_of(noninc.value), (name, step)


def run_distinct_differential(be, seed=0, steps=15):
    """distinct() == integrate().stream_distinct().differentiate() (distinct.rs:931-1005)."""
    rng = np.random.default_rng(100 + seed)
    deltas = _deltas(rng, steps, 10, 5, 0, 3)
    s = Schema("u", "u")
    c = RootCircuit(be)
    it = iter(deltas)
    inp = c.add_source(lambda: be.batch_from_rows(s, next(it)), s).index(1)
    inc = inp.distinct().gather(0).output()
    noninc = inp.integrate().stream_distinct().differentiate().gather(0).output()
    for step in range(steps):
        c.step()
        assert rows_of(inc.value) == rows_of(noninc.value), step


def run_join_differential(be, seed=0, steps=12):
    """join() == join_incremental() == differentiate(stream_join of the integrals) (join.rs:136-155, 180-292)."""
    rng = np.random.default_rng(200 + seed)
    da, db = _deltas(rng, steps, 30, 6, 0, 4), _deltas(rng, steps, 30, 6, 0, 4)
    s = Schema("u", "u")
    c = RootCircuit(be)
    ia, ib = iter(da), iter(db)
    a = c.add_source(lambda: be.batch_from_rows(s, next(ia)), s).index(1)
    b = c.add_source(lambda: be.batch_from_rows(s, next(ib)), s).index(1)
    proj = Proj(Schema("uuu"), [key(0), lval(0), rval(0)])
    j1 = a.join(b, proj).gather(0).output()
    j2 = a.join_incremental(b, proj).gather(0).output()
    j3 = a.integrate().stream_join(b.integrate(), proj).differentiate().gather(0).output()
    for step in range(steps):
        c.step()
        assert rows_of(j1.value) == rows_of(j2.value) == rows_of(j3.value), step


ALL_CASES = {}
for _seed in range(4):
    ALL_CASES[f"aggregate_differential_{_seed}"] = lambda be, s=_seed: run_aggregate_differential(be, seed=s)
    ALL_CASES[f"distinct_differential_{_seed}"] = lambda be, s=_seed: run_distinct_differential(be, seed=s)
    ALL_CASES[f"join_differential_{_seed}"] = lambda be, s=_seed: run_join_differential(be, seed=s)
ALL_CASES["aggregate_differential_big"] = lambda be: run_aggregate_differential(be, seed=9, steps=10, max_tuples=3000, nkeys=200, maxval=50)
