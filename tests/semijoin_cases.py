"""Backend-generic tests of SemiJoinStream::eval (operator/semijoin.rs:100-142):
for every key present in both inputs, every (key, val) pair of `pairs` comes out
with weight w_pair * w_key; the output is an OrdZSet<(K, V)> (semijoin.rs:47).
The model is a dict of Python ints and shares no code with either library."""
import numpy as np

from dbsp_b200 import RootCircuit, Schema


def model(pairs, keys, nk):
    kw = {}
    for r in keys:
        kw[tuple(r[:nk])] = kw.get(tuple(r[:nk]), 0) + r[-1]
    acc = {}
    for r in pairs:
        acc[tuple(r[:-1])] = acc.get(tuple(r[:-1]), 0) + r[-1]
    out = []
    for t, w in sorted(acc.items()):
        k = kw.get(t[:nk], 0)
        if w != 0 and k != 0:
            out.append(t + (w * k,))
    return out


def rows_of(b):
    return [tuple(int(x) for x in r) for r in b.rows()]


def run_semijoin_literals(be):
    """weights != 1 (also negative), keys absent on either side, empty sides."""
    sp, sk = Schema("u", "u"), Schema("u")
    P = [(1, 10, 1), (1, 11, 2), (2, 20, -3), (4, 40, 5), (7, 70, 1), (7, 71, -1)]
    K = [(1, 2), (2, -1), (3, 9), (7, 3)]
    out = be.semijoin(be.batch_from_rows(sp, P), be.batch_from_rows(sk, K))
    assert out.schema == Schema("uu"), out.schema
    assert rows_of(out) == [(1, 10, 2), (1, 11, 4), (2, 20, 3), (7, 70, 3), (7, 71, -3)]
    assert rows_of(out) == model(P, K, 1)
    # empty sides
    assert len(be.semijoin(be.batch_empty(sp), be.batch_from_rows(sk, K))) == 0
    assert len(be.semijoin(be.batch_from_rows(sp, P), be.batch_empty(sk))) == 0
    # disjoint key sets
    assert len(be.semijoin(be.batch_from_rows(sp, P), be.batch_from_rows(sk, [(0, 1), (5, 1), (9, 1)]))) == 0
    # an OrdZSet<K> on the pairs side (no value lanes)
    o2 = be.semijoin(be.batch_from_rows(sk, [(1, 3), (3, 2), (8, 1)]), be.batch_from_rows(sk, K))
    assert rows_of(o2) == [(1, 6), (3, 18)]


def run_semijoin_random(be, seed=0, n_pairs=2000, n_keys=300, domain=500, schema_p=Schema("ui", "uu")):
    rng = np.random.default_rng(300 + seed)
    nk = schema_p.nk
    sk = Schema(schema_p.key)

    def lane(t, n):
        return rng.integers(-domain, domain, n) if t == "i" else rng.integers(0, domain, n)

    P = list(zip(*[lane(t, n_pairs).tolist() for t in schema_p.lanes], rng.integers(-3, 4, n_pairs).tolist()))
    K = list(zip(*[lane(t, n_keys).tolist() for t in sk.lanes], rng.integers(-3, 4, n_keys).tolist()))
    out = be.semijoin(be.batch_from_rows(schema_p, P), be.batch_from_rows(sk, K))
    assert out.schema == Schema(schema_p.lanes)
    got = rows_of(out)
    want = model(P, K, nk)
    assert got == want, (len(got), len(want))


def run_semijoin_stream(be, steps=6):
    """semijoin_stream through the circuit API (stateless, per-step batches)."""
    rng = np.random.default_rng(77)
    sp, sk = Schema("u", "u"), Schema("u")
    c = RootCircuit(be)
    a, ha = c.add_input_indexed_zset(sp)
    k, hk = c.add_input_zset(sk)
    out = a.semijoin_stream(k).output()
    for _ in range(steps):
        na, nb = int(rng.integers(0, 400)), int(rng.integers(0, 60))
        P = list(zip(rng.integers(0, 80, na).tolist(), rng.integers(0, 9, na).tolist(), rng.integers(-2, 3, na).tolist()))
        K = list(zip(rng.integers(0, 80, nb).tolist(), rng.integers(-2, 3, nb).tolist()))
        ha.append(P)
        hk.append(K)
        c.step()
        assert rows_of(out.value) == model(P, K, 1)


ALL_CASES = {"literals": run_semijoin_literals, "stream": run_semijoin_stream}
for _s in range(4):
    ALL_CASES[f"random_{_s}"] = lambda be, s=_s: run_semijoin_random(be, seed=s)
ALL_CASES["random_wide"] = lambda be: run_semijoin_random(be, seed=9, n_pairs=60_000, n_keys=9_000, domain=20_000, schema_p=Schema("uu", "iuu"))
ALL_CASES["random_dense"] = lambda be: run_semijoin_random(be, seed=10, n_pairs=50_000, n_keys=40, domain=40, schema_p=Schema("u", "u"))
