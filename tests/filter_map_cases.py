"""The reference's `filter_map_test` (operator/filter_map.rs:737-910): filter / map / flat_map / map_index /
flat_map_index over an OrdZSet<isize> and an OrdIndexedZSet<isize, String>, literal input and outputs.  Closures are
written in the declarative row language of the ABI (`dbsp_proj`); the two `abs` variants have no counterpart there
(no ABS operator) and are left out.  Strings are order-preserving dictionary codes."""
from dbsp_b200 import Proj, Schema, key, val

CODE = {"-1": 0, "-2": 1, "1": 2, "5 foo": 3}          # sorted by the strings' byte order
FOO = [c for s, c in CODE.items() if "foo" in s]
INPUT = [(1, "1"), (-1, "-1"), (5, "5 foo"), (-2, "-2")]

INTS, IDX = Schema("i"), Schema("i", "u")
n, s = key(0), val(0)
pos = n.gt(0).signed()
foo = s.isin(FOO)
S = CODE

# name -> (input, Proj, expected {row lanes: weight})
CASES = {
    "filter_pos": ("ints", Proj(INTS, [n], [pos]), {(1,): 1, (5,): 1}),
    "indexed": ("ints", Proj(Schema("i", "i"), [n, n]), {(1, 1): 1, (-1, -1): 1, (5, 5): 1, (-2, -2): 1}),
    "times2": ("ints", Proj(INTS, [n * 2]), {(2,): 1, (-2,): 1, (10,): 1, (-4,): 1}),
    "times2_pos": ("ints", Proj(INTS, [n * 2], [pos]), {(2,): 1, (10,): 1}),
    "neg": ("ints", Proj(INTS, [-n]), {(-1,): 1, (1,): 1, (-5,): 1, (2,): 1}),
    "neg_pos": ("ints", Proj(INTS, [-n], [pos]), {(-1,): 1, (-5,): 1}),
    "sqr": ("ints", Proj(INTS, [n * n]), {(1,): 2, (25,): 1, (4,): 1}),
    "sqr_pos": ("ints", Proj(INTS, [n * n], [pos]), {(1,): 1, (25,): 1}),
    "sqr_pos_indexed": ("ints", Proj(Schema("i", "i"), [n * n, n], [pos]), {(1, 1): 1, (25, 5): 1}),
    "i_filter_pos": ("indexed", Proj(IDX, [n, s], [pos, foo]), {(5, S["5 foo"]): 1}),
    "i_indexed": ("indexed", Proj(IDX, [n * 2, s]), {(2, S["1"]): 1, (-2, S["-1"]): 1, (10, S["5 foo"]): 1, (-4, S["-2"]): 1}),
    "i_times2": ("indexed", Proj(INTS, [n * 2]), {(2,): 1, (-2,): 1, (10,): 1, (-4,): 1}),
    "i_times2_pos": ("indexed", Proj(INTS, [n * 2], [pos, foo]), {(10,): 1}),
    "i_neg": ("indexed", Proj(INTS, [-n]), {(-1,): 1, (1,): 1, (-5,): 1, (2,): 1}),
    "i_neg_pos": ("indexed", Proj(INTS, [-n], [pos, foo]), {(-5,): 1}),
    "i_sqr": ("indexed", Proj(INTS, [n * n]), {(1,): 2, (25,): 1, (4,): 1}),
    "i_sqr_pos": ("indexed", Proj(INTS, [n * n], [pos, foo]), {(25,): 1}),
    "i_sqr_pos_indexed": ("indexed", Proj(IDX, [n * n, s], [pos]), {(1, S["1"]): 1, (25, S["5 foo"]): 1}),
}


def run_filter_map_case(be, name):
    src, proj, want = CASES[name]
    pairs = be.batch_from_rows(Schema("iu"), [(a, CODE[b], 1) for a, b in INPUT])   # OrdZSet<(isize, String)>
    indexed = be.reindex(pairs, 1)                                                    # input.index()
    ints = be.map_index(indexed, Proj(INTS, [key(0)]))                                # input_indexed.map(|(&x, _)| x)
    out = be.map_index(ints if src == "ints" else indexed, proj)
    got = {tuple(int(x) for x in r[:-1]): int(r[-1]) for r in out.rows()}
    assert got == want, (name, got)
    assert out.schema.nk == proj.schema.nk and out.schema.nv == proj.schema.nv


# ---- index_test / index_with_test (operator/index.rs:240-300): zset! literals with repeated rows (the macro sums
# them), index() / index_with(|&(k, v)| (k, v)), integrate(); outputs are the canonical indexed Z-sets.
INDEX_INPUTS = [
    [(1, "a", 1), (1, "b", 1), (2, "a", 1), (2, "c", 1), (1, "a", 2), (1, "b", -1)],
    [(1, "d", 1), (1, "e", 1), (2, "a", -1), (3, "a", 2)],
]
INDEX_OUTPUTS = [
    {1: {"a": 3}, 2: {"a": 1, "c": 1}},
    {1: {"a": 3, "d": 1, "e": 1}, 2: {"c": 1}, 3: {"a": 2}},
]


def run_index_test(be, with_closure):
    pairs_schema, idx_schema = Schema("uu"), Schema("u", "u")
    acc = be.batch_empty(idx_schema)
    for rows, want in zip(INDEX_INPUTS, INDEX_OUTPUTS):
        z = be.batch_from_rows(pairs_schema, [(k, ord(c), w) for k, c, w in rows])
        if with_closure:   # index_with(|&(k, v)| (k, v)): a projection of the two key lanes into (key, value)
            indexed = be.map_index(z, Proj(idx_schema, [key(0), key(1)]))
        else:              # index(): the zero-copy re-interpretation of the lane split
            indexed = be.reindex(z, 1)
        acc = be.merge(acc, indexed)   # integrate()
        d = acc.download()
        offs = [int(x) for x in d["offs"]]
        got = {int(k): {chr(int(d["vals"][0][i])): int(d["diffs"][i]) for i in range(offs[j], offs[j + 1])}
               for j, k in enumerate(d["keys"][0])}
        assert got == want, got


# ---- zset_sum of operator/neg.rs:85-110 and operator/sum.rs:138-200: 100 steps of a growing Z-set;
# neg().plus(source) == 0, and sum(source3, [source2, source1, source3]) == 0 (source3 supplied twice).
def run_neg_plus_zero(be, steps=100):
    sch = Schema("u")
    s = be.batch_empty(sch)
    step = be.batch_from_rows(sch, [(5, 1), (6, 2)])
    for _ in range(steps):
        assert len(be.merge(be.neg(s), s)) == 0
        s = be.merge(s, step)


def run_sum_zero(be, steps=100):
    sch = Schema("u")
    s1 = s2 = s3 = be.batch_empty(sch)
    d1, d2, d3 = be.batch_from_rows(sch, [(5, 1), (6, 2)]), be.batch_from_rows(sch, [(5, -1)]), be.batch_from_rows(sch, [(6, -1)])
    for _ in range(steps):
        total = s3
        for x in (s2, s1, s3):
            total = be.merge(total, x)
        assert len(total) == 0, total.rows()
        s1, s2, s3 = be.merge(s1, d1), be.merge(s2, d2), be.merge(s3, d3)


# ---- the same test through the circuit API, shaped like the reference's (one circuit, every operator inspected) ----
def run_filter_map_circuit(be):
    from dbsp_b200 import RootCircuit

    c = RootCircuit(be)
    it = iter([be.batch_from_rows(Schema("iu"), [(a, CODE[b], 1) for a, b in INPUT])])
    inp = c.add_source(lambda: next(it), Schema("iu"))
    input_indexed = inp.index(1)
    input_ints = input_indexed.map(Proj(INTS, [key(0)]))
    streams = {
        "filter_pos": input_ints.filter(pos),
        "indexed": input_ints.map_index(Proj(Schema("i", "i"), [n, n])),
        "times2": input_ints.map(Proj(INTS, [n * 2])),
        "times2_pos": input_ints.flat_map(Proj(INTS, [n * 2], [pos])),
        "neg": input_ints.map(Proj(INTS, [-n])),
        "neg_pos": input_ints.flat_map(Proj(INTS, [-n], [pos])),
        "sqr": input_ints.map(Proj(INTS, [n * n])),
        "sqr_pos": input_ints.flat_map(Proj(INTS, [n * n], [pos])),
        "sqr_pos_indexed": input_ints.flat_map_index(Proj(Schema("i", "i"), [n * n, n], [pos])),
        "i_filter_pos": input_indexed.filter(pos, foo),
        "i_indexed": input_indexed.map_index(Proj(IDX, [n * 2, s])),
        "i_times2": input_indexed.map(Proj(INTS, [n * 2])),
        "i_times2_pos": input_indexed.flat_map(Proj(INTS, [n * 2], [pos, foo])),
        "i_neg": input_indexed.map(Proj(INTS, [-n])),
        "i_neg_pos": input_indexed.flat_map(Proj(INTS, [-n], [pos, foo])),
        "i_sqr": input_indexed.map(Proj(INTS, [n * n])),
        "i_sqr_pos": input_indexed.flat_map(Proj(INTS, [n * n], [pos, foo])),
        "i_sqr_pos_indexed": input_indexed.flat_map_index(Proj(IDX, [n * n, s], [pos])),
    }
    assert set(streams) == set(CASES)
    seen = {}
    for name, st in streams.items():
        st.inspect(lambda b, name=name: seen.__setitem__(name, {tuple(int(x) for x in r[:-1]): int(r[-1]) for r in b.rows()}))
    c.step()
    for name, (_, _, want) in CASES.items():
        assert seen[name] == want, (name, seen[name])


def run_sum_circuit(be, steps=20):
    """zset_sum of sum.rs through Stream.sum, source3 supplied twice."""
    from dbsp_b200 import RootCircuit

    sch = Schema("u")
    c = RootCircuit(be)
    state = {"s1": be.batch_empty(sch), "s2": be.batch_empty(sch), "s3": be.batch_empty(sch)}
    deltas = {"s1": [(5, 1), (6, 2)], "s2": [(5, -1)], "s3": [(6, -1)]}

    def gen(k):
        def g():
            res = state[k]
            state[k] = be.merge(res, be.batch_from_rows(sch, deltas[k]))
            return res
        return g

    s1, s2, s3 = (c.add_source(gen(k), sch) for k in ("s1", "s2", "s3"))
    sizes = []
    s3.sum([s2, s1, s3]).inspect(lambda b: sizes.append(len(b)))
    for _ in range(steps):
        c.step()
    assert sizes == [0] * steps


# ---- zset_test_st (operator/input.rs:1058-1100): append / push (with a cancelling pair) / clear_input on the
# add_input_zset handle; the stream sees input_batches(), input_batches() again, then the empty Z-set.
INPUT_BATCHES = [{1: 1, 2: 1, 3: 1}, {5: -1, 10: 2, 11: 11}, {}]


def run_input_zset_test(be):
    from dbsp_b200 import RootCircuit

    c = RootCircuit(be)
    stream, handle = c.add_input_zset(Schema("u"))
    seen = []
    stream.inspect(lambda b: seen.append({int(r[0]): int(r[1]) for r in b.rows()}))
    vecs = [sorted(d.items()) for d in INPUT_BATCHES]
    for v in vecs:
        handle.append(list(v))
        c.step()
    for v in vecs:
        for k, w in v:
            handle.push(k, w)
        handle.push(5, 1)
        handle.push(5, -1)
        c.step()
    for v in vecs:
        handle.append(list(v))
    handle.clear_input()
    c.step()
    assert seen == INPUT_BATCHES + INPUT_BATCHES + [{}], seen
